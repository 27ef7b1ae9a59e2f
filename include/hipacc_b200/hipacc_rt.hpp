// hipacc_rt.hpp -- Hipacc-compatible C++ RUNTIME surface on top of the C ABI (include/hipacc_b200.h).
//
// This header takes the place of runtime/hipacc_cu.hpp + hipacc_cu.tpp + hipacc_cu_standalone.hpp
// (paths relative to the Hipacc tree) for host code that Hipacc's rewriter emits
// (lib/Rewrite/CreateHostStrings.cpp): the same names, argument meaning and error behaviour
// (log and continue, runtime/hipacc_cu.hpp:69-75) for init / memory / accessors / pyramids /
// timing, and descriptor-taking launch calls where the reference launches a generated kernel:
//
//   reference (generated)                                   here
//   hipaccLaunchKernel(kernelFn, grid, block, ep, t, smem, args...)   hipaccLaunchLocalOperator / hipaccLaunchPointOperator /
//                                                                      hipaccLaunchBilateral / hipaccLaunchHarris (in, is, desc, ep, t)
//   hipaccApplyReductionShared<T>(kernelFn, acc, threads, ppt, ep, tex, t)   hipaccApplyReduction<T>(acc, mode, ep, t)
//   hipaccApplyBinningSegmented<T,T2,...>(kernelFn, acc, warps, units, bins, ep, tex, t)   hipaccApplyBinning<T>(acc, bins, index, value, p0, ep, t)
//
// Header-only; link with -lhipacc_b200.  No CUDA headers are needed by the including translation unit.
#ifndef HIPACC_B200_RT_HPP
#define HIPACC_B200_RT_HPP

#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../hipacc_b200.h"

#ifndef HIPACC_B200_NO_TYPEDEFS
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
#endif

// vector pixel types of the DSL (dsl/types.hpp:56-100): 4 interleaved channels
#ifndef HIPACC_B200_NO_VECTOR_TYPES
struct uchar4 { unsigned char x, y, z, w; };
struct int4 { int x, y, z, w; };
struct float4 { float x, y, z, w; };
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
#endif

namespace hipacc_b200 {

template <typename T> struct dtype_of;
template <> struct dtype_of<unsigned char> { static constexpr int value = HB_U8; };
template <> struct dtype_of<signed char> { static constexpr int value = HB_S8; };
template <> struct dtype_of<char> { static constexpr int value = HB_S8; };
template <> struct dtype_of<unsigned short> { static constexpr int value = HB_U16; };
template <> struct dtype_of<short> { static constexpr int value = HB_S16; };
template <> struct dtype_of<int> { static constexpr int value = HB_S32; };
template <> struct dtype_of<unsigned int> { static constexpr int value = HB_U32; };
template <> struct dtype_of<float> { static constexpr int value = HB_F32; };
#ifndef HIPACC_B200_NO_VECTOR_TYPES
template <> struct dtype_of<uchar4> { static constexpr int value = HB_U8X4; };
#endif

// checkErr (runtime/hipacc_cu.hpp:69-75): the library has already logged the error; execution continues
inline void check(int rc, const char *what) {
    if (rc != HB_OK) std::fprintf(stderr, "ERROR: %s (%d): %s\n", what, rc, hb_last_error());
}

inline bool &timing_enabled() {
    static bool on = true;  // the reference brackets every launch with events (hipacc_cu_standalone.hpp:297-326)
    return on;
}

}  // namespace hipacc_b200

// ---------------------------------------------------------------------------------------------------
// init / timing
// ---------------------------------------------------------------------------------------------------
// hipaccInitCUDA (runtime/hipacc_cu_standalone.hpp:113-163).  HIPACC_B200_DEVICE selects the device of this
// process (one process per GPU); the reference always uses device 0.
inline void hipaccInitCUDA() {
    static bool done = false;
    if (done) return;
    const char *e = std::getenv("HIPACC_B200_DEVICE");
    hipacc_b200::check(hb_init(e ? std::atoi(e) : 0), "hipaccInitCUDA()");
    hb_set_timing(hipacc_b200::timing_enabled() ? 1 : 0);
    done = true;
}
// Launches are bracketed by events and synchronised (reference behaviour) unless switched off
inline void hipaccSetTiming(bool on) {
    hipacc_b200::timing_enabled() = on;
    hb_set_timing(on ? 1 : 0);
}
// hipacc_last_kernel_timing (runtime/hipacc_base.hpp:64-66), milliseconds
inline float hipacc_last_kernel_timing() { return hb_last_kernel_ms(); }

// ---------------------------------------------------------------------------------------------------
// execution parameter (runtime/hipacc_cu.hpp:234-245): the stream + pre/post hooks of a launch
// ---------------------------------------------------------------------------------------------------
class HipaccExecutionParameterCudaBase {
    void *stream_{};

  protected:
    void set_stream(void *s) { stream_ = s; }

  public:
    virtual ~HipaccExecutionParameterCudaBase() = default;
    void *get_stream() const { return stream_; }  // cudaStream_t
    virtual void pre_kernel() {}
    virtual void post_kernel() {}
};
using HipaccExecutionParameterCuda = std::shared_ptr<HipaccExecutionParameterCudaBase>;
inline HipaccExecutionParameterCuda hipaccMapExecutionParameter(HipaccExecutionParameterCuda ep) { return ep; }
// convenience: run on an existing cudaStream_t
class HipaccStreamParameter final : public HipaccExecutionParameterCudaBase {
  public:
    explicit HipaccStreamParameter(void *stream) { set_stream(stream); }
};

// ---------------------------------------------------------------------------------------------------
// images (runtime/hipacc_cu.hpp:91-160, hipacc_cu.tpp:32-193)
// ---------------------------------------------------------------------------------------------------
template <typename T> class HipaccImageCudaBase {
  public:
    using pixel_type = T;
    virtual ~HipaccImageCudaBase() = default;
    virtual pixel_type const *get_device_memory() const = 0;
    virtual pixel_type *get_device_memory() = 0;
    virtual pixel_type const *get_host_memory() const = 0;
    virtual pixel_type *get_host_memory() = 0;
    virtual int get_width() const = 0;
    virtual int get_height() const = 0;
    virtual int get_stride() const = 0;
    virtual int get_alignment() const = 0;
    virtual int get_pixel_size() const = 0;
    virtual const hb_view &get_view() const = 0;  // whole image as the C ABI sees it
};
template <typename T> using HipaccImageCuda = std::shared_ptr<HipaccImageCudaBase<T>>;

template <typename T> class HipaccImageCudaRaw final : public HipaccImageCudaBase<T> {
    hb_view view_{};
    int alignment_{};
    bool owns_{};
    std::unique_ptr<T[]> host_;  // the host mirror hipaccReadMemory hands out (hipacc_cu.hpp:129-131)

  public:
    HipaccImageCudaRaw(const hb_view &v, int alignment, bool owns)
        : view_(v), alignment_(alignment), owns_(owns), host_(new T[(size_t)v.img_width * v.img_height]()) {}
    ~HipaccImageCudaRaw() override {
        if (owns_) hipacc_b200::check(hb_image_destroy(&view_), "cudaFree()");
    }
    HipaccImageCudaRaw(const HipaccImageCudaRaw &) = delete;
    HipaccImageCudaRaw &operator=(const HipaccImageCudaRaw &) = delete;
    T const *get_device_memory() const override { return static_cast<T const *>(view_.data); }
    T *get_device_memory() override { return static_cast<T *>(view_.data); }
    T const *get_host_memory() const override { return host_.get(); }
    T *get_host_memory() override { return host_.get(); }
    int get_width() const override { return view_.img_width; }
    int get_height() const override { return view_.img_height; }
    int get_stride() const override { return view_.stride; }
    int get_alignment() const override { return alignment_; }
    int get_pixel_size() const override { return (int)sizeof(T); }
    const hb_view &get_view() const override { return view_; }
};

template <typename T> void hipaccWriteMemory(HipaccImageCuda<T> &img, T *host_mem);

// hipaccCreateMemory<T>(host, w, h[, alignment]) (hipacc_cu.tpp:51-68).  alignment in bytes; 0 = library default
// (rows padded to 256 bytes so that each row is TMA addressable).
template <typename T> HipaccImageCuda<T> hipaccCreateMemory(T *host_mem, size_t width, size_t height, size_t alignment) {
    hipaccInitCUDA();
    hb_view v{};
    hipacc_b200::check(hb_image_create(hipacc_b200::dtype_of<T>::value, (int)width, (int)height, (int)alignment, &v), "hipaccCreateMemory()");
    HipaccImageCuda<T> img = std::make_shared<HipaccImageCudaRaw<T>>(v, (int)alignment, true);
    if (host_mem) hipaccWriteMemory(img, host_mem);
    return img;
}
template <typename T> HipaccImageCuda<T> hipaccCreateMemory(T *host_mem, size_t width, size_t height) {
    return hipaccCreateMemory<T>(host_mem, width, height, 0);
}
// hipaccMapMemory (hipacc_cu.hpp:286): identity on an image ...
template <typename T> HipaccImageCuda<T> hipaccMapMemory(HipaccImageCuda<T> img) { return img; }
// ... and the zero-copy hook for buffers that already live in HBM (stride in pixels)
template <typename T> HipaccImageCuda<T> hipaccMapMemory(T *device_mem, size_t width, size_t height, size_t stride) {
    hipaccInitCUDA();
    hb_view v{};
    hipacc_b200::check(hb_image_wrap(device_mem, hipacc_b200::dtype_of<T>::value, (int)width, (int)height, (int)stride, &v), "hipaccMapMemory()");
    return std::make_shared<HipaccImageCudaRaw<T>>(v, 0, false);
}

// blocking copies (hipacc_cu.tpp:85-116,166-193); host arrays are dense (stride == width)
template <typename T> void hipaccWriteMemory(HipaccImageCuda<T> &img, T *host_mem) {
    if (host_mem == nullptr) return;
    const size_t n = (size_t)img->get_width() * img->get_height();
    if (host_mem != img->get_host_memory()) std::memcpy(img->get_host_memory(), host_mem, n * sizeof(T));
    hipacc_b200::check(hb_image_write(&img->get_view(), host_mem, nullptr), "hipaccWriteMemory()");
}
template <typename T> T *hipaccReadMemory(const HipaccImageCuda<T> &img) {
    hipacc_b200::check(hb_image_read(&img->get_view(), img->get_host_memory(), nullptr), "hipaccReadMemory()");
    return img->get_host_memory();
}
template <typename T> void hipaccCopyMemory(const HipaccImageCuda<T> &src, HipaccImageCuda<T> &dst) {
    hipacc_b200::check(hb_image_copy(&src->get_view(), &dst->get_view(), nullptr), "hipaccCopyMemory()");
}

// ---------------------------------------------------------------------------------------------------
// accessors = image + region, for Accessor and IterationSpace alike (hipacc_cu.hpp:162-186)
// ---------------------------------------------------------------------------------------------------
struct HipaccAccessorBase {
    size_t width, height;
    int32_t offset_x, offset_y;
    HipaccAccessorBase(size_t w, size_t h, int32_t ox, int32_t oy) : width(w), height(h), offset_x(ox), offset_y(oy) {}
};
template <typename T> struct HipaccAccessor : public HipaccAccessorBase {
    HipaccImageCuda<T> img;
    HipaccAccessor(HipaccImageCuda<T> const &img, size_t width, size_t height, int32_t offset_x = 0, int32_t offset_y = 0)
        : HipaccAccessorBase(width, height, offset_x, offset_y), img(img) {}
    HipaccAccessor(HipaccImageCuda<T> const &img) : HipaccAccessorBase(img->get_width(), img->get_height(), 0, 0), img(img) {}
    hb_view view() const {
        hb_view v = img->get_view();
        v.width = (int)width; v.height = (int)height; v.offset_x = offset_x; v.offset_y = offset_y;
        return v;
    }
};
template <typename T> HipaccAccessor<T> hipaccMakeAccessor(HipaccImageCuda<T> const &img) { return HipaccAccessor<T>{img}; }
template <typename T>
HipaccAccessor<T> hipaccMakeAccessor(HipaccImageCuda<T> const &img, size_t width, size_t height, int32_t offset_x = 0, int32_t offset_y = 0) {
    return HipaccAccessor<T>{img, width, height, offset_x, offset_y};
}
template <typename T> void hipaccCopyMemoryRegion(const HipaccAccessor<T> &src, const HipaccAccessor<T> &dst) {
    const hb_view s = src.view(), d = dst.view();
    hipacc_b200::check(hb_image_copy_region(&s, &d, nullptr), "hipaccCopyMemoryRegion()");
}

// ---------------------------------------------------------------------------------------------------
// launches: one call per operator instance (replaces hipaccLaunchKernel + the generated kernel)
// ---------------------------------------------------------------------------------------------------
namespace hipacc_b200 {
struct LaunchScope {  // ep->pre_kernel(), stream, ep->post_kernel() as in hipacc_cu_standalone.hpp:277-329
    const HipaccExecutionParameterCuda &ep;
    bool saved;
    LaunchScope(const HipaccExecutionParameterCuda &ep, bool print_timing) : ep(ep), saved(timing_enabled()) {
        if (print_timing && !saved) hb_set_timing(1);
        if (ep) ep->pre_kernel();
    }
    void *stream() const { return ep ? ep->get_stream() : nullptr; }
    void done(const char *name, bool print_timing) {
        if (ep) ep->post_kernel();
        if (print_timing) {
            std::printf("<HIPACC:> Kernel timing (%s): %f(ms)\n", name, hb_last_kernel_ms());
            if (!saved) hb_set_timing(0);
        }
    }
};
}  // namespace hipacc_b200

// `desc` carries everything but the two views (operator kind, mask, domain, boundary mode, epilogue)
template <typename TI, typename TO>
void hipaccLaunchLocalOperator(const HipaccAccessor<TI> &in, const HipaccAccessor<TO> &is, hb_local_desc desc,
                               HipaccExecutionParameterCuda const &ep = nullptr, bool print_timing = false) {
    hipacc_b200::LaunchScope sc(ep, print_timing);
    desc.in = in.view();
    desc.out = is.view();
    hipacc_b200::check(hb_local_op(&desc, sc.stream()), "hipaccLaunchLocalOperator()");
    sc.done("local operator", print_timing);
}
template <typename T>
void hipaccLaunchBilateral(const HipaccAccessor<T> &in, const HipaccAccessor<T> &is, hb_bilateral_desc desc,
                           HipaccExecutionParameterCuda const &ep = nullptr, bool print_timing = false) {
    hipacc_b200::LaunchScope sc(ep, print_timing);
    desc.in = in.view();
    desc.out = is.view();
    hipacc_b200::check(hb_bilateral(&desc, sc.stream()), "hipaccLaunchBilateral()");
    sc.done("bilateral", print_timing);
}
// inputs[k] / interp[k]: the k-th accessor of the kernel in declaration order (ClassRepresentation.cpp:512-597)
template <typename TI, typename TO>
void hipaccLaunchPointOperator(int op, const std::vector<HipaccAccessor<TI>> &inputs, const std::vector<int> &interp, const HipaccAccessor<TO> &is,
                               double p0 = 0.0, double p1 = 0.0, HipaccExecutionParameterCuda const &ep = nullptr, bool print_timing = false) {
    hipacc_b200::LaunchScope sc(ep, print_timing);
    hb_point_desc d;
    std::memset(&d, 0, sizeof(d));
    d.n_in = (int)inputs.size();
    for (int k = 0; k < d.n_in && k < 3; ++k) {
        d.in[k] = inputs[k].view();
        d.interp[k] = k < (int)interp.size() ? interp[k] : HB_INTERP_NO;
    }
    d.out = is.view();
    d.op = op;
    d.p[0] = p0; d.p[1] = p1;
    hipacc_b200::check(hb_point_op(&d, sc.stream()), "hipaccLaunchPointOperator()");
    sc.done("point operator", print_timing);
}
inline void hipaccLaunchHarris(const HipaccAccessor<uchar> &in, const HipaccAccessor<uchar> &is, float k, float threshold,
                               HipaccExecutionParameterCuda const &ep = nullptr, bool print_timing = false) {
    hipacc_b200::LaunchScope sc(ep, print_timing);
    hb_harris_desc d;
    std::memset(&d, 0, sizeof(d));
    d.in = in.view(); d.out = is.view(); d.k = k; d.threshold = threshold;
    hipacc_b200::check(hb_harris(&d, sc.stream()), "hipaccLaunchHarris()");
    sc.done("harris", print_timing);
}

// hipaccApplyReductionShared (hipacc_cu.tpp:312-408): blocking, returns the scalar by value
template <typename T>
T hipaccApplyReduction(const HipaccAccessor<T> &acc, int reduce_mode, HipaccExecutionParameterCuda const &ep = nullptr, bool print_timing = false) {
    hipacc_b200::LaunchScope sc(ep, print_timing);
    const hb_view v = acc.view();
    T result{};
    hipacc_b200::check(hb_reduce(&v, reduce_mode, &result, sc.stream()), "hipaccApplyReduction()");
    sc.done("reduction", print_timing);
    return result;
}
template <typename T>
T hipaccApplyReduction(const HipaccImageCuda<T> &img, int reduce_mode, HipaccExecutionParameterCuda const &ep = nullptr, bool print_timing = false) {
    return hipaccApplyReduction<T>(HipaccAccessor<T>(img), reduce_mode, ep, print_timing);
}
// hipaccApplyBinningSegmented (hipacc_cu.tpp:410-464): blocking, returns `new T[num_bins]` owned by the caller.
// The binning body is stated as (index kind, value kind, p0) -- hb_bin_index / hb_bin_value -- with reduce = +.
template <typename T, typename BIN = uint>
BIN *hipaccApplyBinning(const HipaccAccessor<T> &acc, unsigned num_bins, int index_kind, int value_kind, double p0,
                        HipaccExecutionParameterCuda const &ep = nullptr, bool print_timing = false) {
    static_assert(sizeof(BIN) == 4, "bins are 32-bit unsigned counters");
    hipacc_b200::LaunchScope sc(ep, print_timing);
    hb_binning_desc d;
    std::memset(&d, 0, sizeof(d));
    d.in = acc.view();
    d.num_bins = (int)num_bins; d.index_kind = index_kind; d.value_kind = value_kind; d.p0 = p0;
    BIN *bins = new BIN[num_bins]();
    hipacc_b200::check(hb_binning(&d, reinterpret_cast<uint32_t *>(bins), sc.stream()), "hipaccApplyBinning()");
    sc.done("binning", print_timing);
    return bins;
}
// -use-graph (runtime/hipacc_cu_standalone.hpp:331-356): record every launch issued on the execution parameter's stream
// between hipaccGraphBegin and hipaccGraphEnd, replay them with hipaccGraphLaunch.  The stream must not be the default one.
class HipaccGraph {
    hb_graph *g_ = nullptr;

  public:
    HipaccGraph() = default;
    HipaccGraph(const HipaccGraph &) = delete;
    HipaccGraph &operator=(const HipaccGraph &) = delete;
    ~HipaccGraph() { hb_graph_destroy(g_); }
    hb_graph **slot() { hb_graph_destroy(g_); g_ = nullptr; return &g_; }
    hb_graph *get() const { return g_; }
};
inline void hipaccGraphBegin(HipaccExecutionParameterCuda const &ep) {
    hipacc_b200::check(hb_graph_begin(ep ? ep->get_stream() : nullptr), "hipaccGraphBegin()");
}
inline void hipaccGraphEnd(HipaccExecutionParameterCuda const &ep, HipaccGraph &graph) {
    hipacc_b200::check(hb_graph_end(ep ? ep->get_stream() : nullptr, graph.slot()), "hipaccGraphEnd()");
}
inline void hipaccGraphLaunch(const HipaccGraph &graph, HipaccExecutionParameterCuda const &ep) {
    hipacc_b200::check(hb_graph_launch(graph.get(), ep ? ep->get_stream() : nullptr), "hipaccGraphLaunch()");
}

// fused min + max + sum in one pass over HBM (float images)
inline void hipaccApplyReductionMinMaxSum(const HipaccAccessor<float> &acc, float &mn, float &mx, float &sum,
                                          HipaccExecutionParameterCuda const &ep = nullptr) {
    const hb_view v = acc.view();
    float r[3] = {0, 0, 0};
    hipacc_b200::check(hb_reduce_minmaxsum_f32(&v, r, ep ? ep->get_stream() : nullptr), "hipaccApplyReductionMinMaxSum()");
    mn = r[0]; mx = r[1]; sum = r[2];
}

// ---------------------------------------------------------------------------------------------------
// pyramids (runtime/hipacc_base_standalone.hpp:85-254, hipacc_cu.hpp:214-231, hipacc_cu.tpp:482-497)
// ---------------------------------------------------------------------------------------------------
class HipaccPyramid {
    const int depth_;
    int level_{0};
    bool bound_{false};

  public:
    explicit HipaccPyramid(int depth) : depth_(depth) {}
    virtual ~HipaccPyramid() = default;
    int depth() const { return depth_; }
    int level() const { return level_; }
    bool is_top_level() const { return level_ == 0; }
    bool is_bottom_level() const { return level_ == depth_ - 1; }
    void levelInc() { ++level_; }
    void levelDec() { --level_; }
    bool bind() { if (bound_) return false; bound_ = true; level_ = 0; return true; }
    void unbind() { bound_ = false; }
};

template <typename T> class HipaccPyramidCuda final : public HipaccPyramid {
    std::vector<HipaccImageCuda<T>> imgs_;

  public:
    explicit HipaccPyramidCuda(int depth) : HipaccPyramid(depth) {}
    void add(const HipaccImageCuda<T> &img) { imgs_.push_back(img); }
    HipaccImageCuda<T> operator()(int relative) {
        assert(level() + relative >= 0 && level() + relative < (int)imgs_.size() && "Accessed pyramid stage is out of bounds.");
        return imgs_.at(level() + relative);
    }
    HipaccImageCuda<T> at(int lvl) { return imgs_.at(lvl); }
    void swap(HipaccPyramidCuda &other) { imgs_.swap(other.imgs_); }
};

// level 0 ALIASES the user image; level l is (w >> l) x (h >> l), truncating (hipacc_cu.tpp:482-497)
template <typename T> HipaccPyramidCuda<T> hipaccCreatePyramid(const HipaccImageCuda<T> &img, size_t depth) {
    HipaccPyramidCuda<T> p((int)depth);
    p.add(img);
    size_t w = img->get_width() / 2, h = img->get_height() / 2;
    for (size_t i = 1; i < depth; ++i) {
        assert(w * h > 0 && "Pyramid stages too deep for image size");
        p.add(hipaccCreateMemory<T>(nullptr, w, h, (size_t)img->get_alignment()));
        w /= 2; h /= 2;
    }
    return p;
}

namespace hipacc_b200 {
struct TraverseState {  // single host thread, like the reference's TU-static globals (dsl/pyramid.hpp:145-146)
    std::vector<const std::function<void()> *> funcs;
    std::vector<std::vector<HipaccPyramid *>> pyramids;
    static TraverseState &get() { static TraverseState s; return s; }
};
}  // namespace hipacc_b200

// hipaccTraverse(pyramids, body): run `body` at level 0; the body recurses with hipaccTraverse(loop, between)
inline void hipaccTraverse(const std::vector<HipaccPyramid *> &pyrs, const std::function<void()> &func) {
    auto &st = hipacc_b200::TraverseState::get();
    for (size_t i = 0; i + 1 < pyrs.size(); ++i) assert(pyrs[i]->depth() == pyrs[i + 1]->depth() && "Pyramid depths do not match.");
    for (auto *p : pyrs) { const bool ok = p->bind(); (void)ok; assert(ok && "Pyramid already bound to another traversal."); }
    st.pyramids.push_back(pyrs);
    st.funcs.push_back(&func);
    func();
    st.funcs.pop_back();
    st.pyramids.pop_back();
    for (auto *p : pyrs) p->unbind();
}
inline void hipaccTraverse(HipaccPyramid &p0, const std::function<void()> &f) { hipaccTraverse(std::vector<HipaccPyramid *>{&p0}, f); }
inline void hipaccTraverse(HipaccPyramid &p0, HipaccPyramid &p1, const std::function<void()> &f) {
    hipaccTraverse(std::vector<HipaccPyramid *>{&p0, &p1}, f);
}
inline void hipaccTraverse(HipaccPyramid &p0, HipaccPyramid &p1, HipaccPyramid &p2, const std::function<void()> &f) {
    hipaccTraverse(std::vector<HipaccPyramid *>{&p0, &p1, &p2}, f);
}
// recursion step: descend one level, run the traversal body `loop` times with `func` in between, ascend
inline void hipaccTraverse(unsigned int loop = 1, const std::function<void()> &func = [] {}) {
    auto &st = hipacc_b200::TraverseState::get();
    assert(!st.pyramids.empty() && "Traverse recursion called outside of traverse.");
    std::vector<HipaccPyramid *> pyrs = st.pyramids.back();
    if (pyrs.at(0)->is_bottom_level()) return;
    for (auto *p : pyrs) p->levelInc();
    for (unsigned int i = 0; i < loop; ++i) {
        (*st.funcs.back())();
        if (i + 1 < loop) func();
    }
    for (auto *p : pyrs) p->levelDec();
}

#endif  // HIPACC_B200_RT_HPP
