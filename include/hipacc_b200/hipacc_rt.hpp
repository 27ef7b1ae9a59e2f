// hipacc_rt.hpp -- Hipacc-compatible C++ RUNTIME surface on top of the C ABI (include/hipacc_b200.h).
//
// This header takes the place of runtime/hipacc_cu.hpp + hipacc_cu.tpp + hipacc_cu_standalone.hpp
// (paths relative to the Hipacc tree) for host code that Hipacc's rewriter emits
// (lib/Rewrite/CreateHostStrings.cpp): the same names, argument meaning and error behaviour
// (log and continue, runtime/hipacc_cu.hpp:69-75) for init / memory / accessors / pyramids /
// timing, and descriptor-taking launch calls where the reference launches a generated kernel:
//
//   reference (generated)                                   here
//   hipaccLaunchKernel(kernelFn, grid, block, ep, t, smem, args...)   the SAME call: `kernelFn` is a hipacc_b200::OperatorKernel (operator +
//                                                                      meaning of each positional argument) instead of a generated __global__;
//                                                                      or directly hipaccLaunchLocalOperator / hipaccLaunchPointOperator /
//                                                                      hipaccLaunchBilateral / hipaccLaunchHarris (in, is, desc, ep, t)
//   hipacc_launch_info, hipaccPrepareKernelLaunch, hipaccCalcGridFromBlock   same names and arithmetic (the pre-built kernels pick their own tiling)
//   hipaccWriteSymbol / hipaccReadSymbol / hipaccWriteDomainFromMask          same names: the "symbol" is the host table the OperatorKernel points at
//   hipaccApplyReductionShared<T>(kernelFn, acc, threads, ppt, ep, tex, t)   same call with a hipacc_b200::ReductionKernel; or hipaccApplyReduction<T>(acc, mode, ep, t)
//   hipaccApplyBinningSegmented<T,T2,...>(kernelFn, acc, warps, units, bins, ep, tex, t)   same call with a hipacc_b200::BinningKernel; or hipaccApplyBinning<T>(...)
//   HipaccPyramidTraversor                                                    same class (runtime/hipacc_base.hpp:159-191)
//
// Header-only; link with -lhipacc_b200.  No CUDA headers are needed by the including translation unit.
#ifndef HIPACC_B200_RT_HPP
#define HIPACC_B200_RT_HPP

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../hipacc_b200.h"

// scalar typedefs + the DSL's vector pixel types with their operator set (dsl/types.hpp:56-516)
#include "hipacc_types.hpp"

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
template <> struct dtype_of<uchar4> { static constexpr int value = HB_U8X4; };
template <> struct dtype_of<char4> { static constexpr int value = HB_S8X4; };
template <> struct dtype_of<ushort4> { static constexpr int value = HB_U16X4; };
template <> struct dtype_of<short4> { static constexpr int value = HB_S16X4; };
template <> struct dtype_of<int4> { static constexpr int value = HB_S32X4; };
template <> struct dtype_of<uint4> { static constexpr int value = HB_U32X4; };
template <> struct dtype_of<float4> { static constexpr int value = HB_F32X4; };

// checkErr (runtime/hipacc_cu.hpp:69-75): the library has already logged the error; execution continues
inline void check(int rc, const char *what) {
    if (rc != HB_OK) std::fprintf(stderr, "ERROR: %s (%d): %s\n", what, rc, hb_last_error());
}

inline bool &timing_enabled() {
    static bool on = true;  // the reference brackets every launch with events (hipacc_cu_standalone.hpp:297-326)
    return on;
}
// timing of kernels compiled from a kernel() body in the user's translation unit (hipacc.hpp under nvcc): they are
// launched by the header, not by the library, so their elapsed time is kept here
inline float &compiled_body_ms() { static float ms = 0.0f; return ms; }
inline bool &compiled_body_was_last() { static bool f = false; return f; }

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
inline float hipacc_last_kernel_timing() { return hipacc_b200::compiled_body_was_last() ? hipacc_b200::compiled_body_ms() : hb_last_kernel_ms(); }

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
        compiled_body_was_last() = false;
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

// Traversor for image pyramid applications (runtime/hipacc_base.hpp:159-191, hipacc_base_standalone.hpp:63-254): the
// stack of traversal bodies and of the pyramids bound to them; hipaccTraverse(pyramids, body) runs `body` at level 0, the
// body recurses with hipaccTraverse(loop, between).
class HipaccPyramidTraversor {
    std::vector<const std::function<void()> *> hipaccTraverseFunc;
    std::vector<std::vector<HipaccPyramid *>> hipaccPyramids;

  public:
    void pushFunc(const std::function<void()> *f) { hipaccTraverseFunc.push_back(f); }
    void popFunc() { hipaccTraverseFunc.pop_back(); }
    const std::function<void()> getLastFunc() { return *hipaccTraverseFunc.back(); }
    void pushPyramids(std::vector<HipaccPyramid *> &pyrs) { hipaccPyramids.push_back(pyrs); }
    void popPyramids() { hipaccPyramids.pop_back(); }
    std::vector<HipaccPyramid *> getLastPyramids() { return hipaccPyramids.back(); }
    bool hasPyramids() { return !hipaccPyramids.empty(); }

    void hipaccTraverse(std::vector<HipaccPyramid *> pyrs, const std::function<void()> &func) {
        for (size_t i = 0; i + 1 < pyrs.size(); ++i) assert(pyrs[i]->depth() == pyrs[i + 1]->depth() && "Pyramid depths do not match.");
        for (auto *p : pyrs) { const bool ok = p->bind(); (void)ok; assert(ok && "Pyramid already bound to another traversal."); }
        pushPyramids(pyrs);
        pushFunc(&func);
        func();
        popFunc();
        popPyramids();
        for (auto *p : pyrs) p->unbind();
    }
    void hipaccTraverse(HipaccPyramid &p0, const std::function<void()> &f) { hipaccTraverse(std::vector<HipaccPyramid *>{&p0}, f); }
    void hipaccTraverse(HipaccPyramid &p0, HipaccPyramid &p1, const std::function<void()> &f) { hipaccTraverse(std::vector<HipaccPyramid *>{&p0, &p1}, f); }
    void hipaccTraverse(HipaccPyramid &p0, HipaccPyramid &p1, HipaccPyramid &p2, const std::function<void()> &f) {
        hipaccTraverse(std::vector<HipaccPyramid *>{&p0, &p1, &p2}, f);
    }
    void hipaccTraverse(HipaccPyramid &p0, HipaccPyramid &p1, HipaccPyramid &p2, HipaccPyramid &p3, const std::function<void()> &f) {
        hipaccTraverse(std::vector<HipaccPyramid *>{&p0, &p1, &p2, &p3}, f);
    }
    void hipaccTraverse(HipaccPyramid &p0, HipaccPyramid &p1, HipaccPyramid &p2, HipaccPyramid &p3, HipaccPyramid &p4, const std::function<void()> &f) {
        hipaccTraverse(std::vector<HipaccPyramid *>{&p0, &p1, &p2, &p3, &p4}, f);
    }
    // recursion step: descend one level, run the traversal body `loop` times with `func` in between, ascend
    void hipaccTraverse(unsigned int loop = 1, const std::function<void()> &func = [] {}) {
        assert(hasPyramids() && "Traverse recursion called outside of traverse.");
        std::vector<HipaccPyramid *> pyrs = getLastPyramids();
        if (pyrs.at(0)->is_bottom_level()) return;
        for (auto *p : pyrs) p->levelInc();
        for (unsigned int i = 0; i < loop; ++i) {
            (*hipaccTraverseFunc.back())();
            if (i + 1 < loop) func();
        }
        for (auto *p : pyrs) p->levelDec();
    }
};

namespace hipacc_b200 {
// the traversor behind the free hipaccTraverse functions the rewritten host code calls (single host thread, like the
// reference's translation-unit static, dsl/pyramid.hpp:145-146)
inline HipaccPyramidTraversor &traversor() { static HipaccPyramidTraversor t; return t; }
}  // namespace hipacc_b200

inline void hipaccTraverse(const std::vector<HipaccPyramid *> &pyrs, const std::function<void()> &func) { hipacc_b200::traversor().hipaccTraverse(pyrs, func); }
inline void hipaccTraverse(HipaccPyramid &p0, const std::function<void()> &f) { hipacc_b200::traversor().hipaccTraverse(p0, f); }
inline void hipaccTraverse(HipaccPyramid &p0, HipaccPyramid &p1, const std::function<void()> &f) { hipacc_b200::traversor().hipaccTraverse(p0, p1, f); }
inline void hipaccTraverse(HipaccPyramid &p0, HipaccPyramid &p1, HipaccPyramid &p2, const std::function<void()> &f) {
    hipacc_b200::traversor().hipaccTraverse(p0, p1, p2, f);
}
inline void hipaccTraverse(unsigned int loop = 1, const std::function<void()> &func = [] {}) { hipacc_b200::traversor().hipaccTraverse(loop, func); }

// ---------------------------------------------------------------------------------------------------
// The launch-side names of the reference runtime (runtime/hipacc_cu.hpp:247-258,295-305,336-360, hipacc_base.hpp:87-101,
// hipacc_cu_standalone.hpp:66-110): host code emitted by lib/Rewrite/CreateHostStrings.cpp:650-1190 links against these
// unchanged.  What changes is the `kernel` it passes: not the address of a generated __global__ function but a
// hipacc_b200::OperatorKernel value -- the operator plus the meaning of each positional argument.
// ---------------------------------------------------------------------------------------------------
#ifndef __CUDACC__
#ifndef HIPACC_B200_NO_DIM3
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x = 1, unsigned int y = 1, unsigned int z = 1) : x(x), y(y), z(z) {}
};
#endif
#endif

struct hipacc_launch_info {   // runtime/hipacc_base.hpp:87-101
    hipacc_launch_info(int size_x, int size_y, int is_width, int is_height, int offset_x, int offset_y, int pixels_per_thread, int simd_width)
        : size_x(size_x), size_y(size_y), is_width(is_width), is_height(is_height), offset_x(offset_x), offset_y(offset_y),
          pixels_per_thread(pixels_per_thread), simd_width(simd_width), bh_start_left(0), bh_start_right(0), bh_start_top(0), bh_start_bottom(0),
          bh_fall_back(0) {}
    hipacc_launch_info(int size_x, int size_y, const HipaccAccessorBase &Acc, int pixels_per_thread, int simd_width)
        : hipacc_launch_info(size_x, size_y, (int)Acc.width, (int)Acc.height, Acc.offset_x, Acc.offset_y, pixels_per_thread, simd_width) {}
    int size_x, size_y;
    int is_width, is_height;
    int offset_x, offset_y;
    int pixels_per_thread, simd_width;
    // calculated by hipaccPrepareKernelLaunch
    int bh_start_left, bh_start_right;
    int bh_start_top, bh_start_bottom;
    int bh_fall_back;
};

// First block (per direction) that needs no border handling / first one that needs it again, for a grid of `block`-sized
// CTAs over the iteration space: the reference's generated kernels branch on these (hipacc_cu_standalone.hpp:66-103).
// The pre-built kernels here classify their tiles themselves (interior / border decided by the tile loader), so the
// values are computed for the caller's benefit and otherwise unused.
inline void hipaccPrepareKernelLaunch(hipacc_launch_info &info, dim3 const &block) {
    const float bw = (float)(block.x * info.simd_width), bh = (float)(block.y * info.pixels_per_thread);
    info.bh_start_left = info.size_x > 0 ? (int)std::ceil((float)(info.offset_x + info.size_x) / bw) : 0;
    info.bh_start_right = (int)std::floor((float)(info.offset_x + info.is_width - (info.size_x > 0 ? info.size_x : 0)) / bw);
    if (info.size_y > 0) {
        const int p_add = (int)std::ceil(2 * info.size_y / (float)block.y);   // blocks staged additionally for shared memory
        info.bh_start_top = (int)std::ceil((float)info.size_y / bh);
        info.bh_start_bottom = (int)std::floor((float)(info.is_height - p_add * (int)block.y) / bh);
    } else {
        info.bh_start_top = 0;
        info.bh_start_bottom = (int)std::floor((float)info.is_height / bh);
    }
    info.bh_fall_back = ((info.bh_start_right - info.bh_start_left) > 1 && (info.bh_start_bottom - info.bh_start_top) > 1) ? 0 : 1;
}
inline dim3 hipaccCalcGridFromBlock(hipacc_launch_info const &info, dim3 const &block) {   // hipacc_cu_standalone.hpp:105-110
    return dim3((unsigned)std::ceil((float)(info.is_width + info.offset_x) / (float)(block.x * info.simd_width)),
                (unsigned)std::ceil((float)info.is_height / (float)(block.y * info.pixels_per_thread)));
}

// hipaccWriteSymbol / hipaccReadSymbol / hipaccWriteDomainFromMask (runtime/hipacc_cu.tpp:262-308): the reference copies a
// Mask into the __constant__ array of the generated kernel.  Here the "symbol" is the host table the OperatorKernel's
// descriptor points at; the library copies it into the kernel parameter (constant) bank at every launch.
template <typename T> void hipaccWriteSymbol(const void *symbol, T *host_mem, size_t width, size_t height) {
    std::memcpy(const_cast<void *>(symbol), host_mem, sizeof(T) * width * height);
}
template <typename T> void hipaccReadSymbol(T *host_mem, const void *symbol, std::string /*symbol_name*/, size_t width, size_t height) {
    std::memcpy(host_mem, symbol, sizeof(T) * width * height);
}
template <typename T> void hipaccWriteDomainFromMask(const void *symbol, T *host_mem, size_t width, size_t height) {
    unsigned char *dom = static_cast<unsigned char *>(const_cast<void *>(symbol));
    for (size_t i = 0; i < width * height; ++i) dom[i] = host_mem[i] == T(0) ? 0 : 1;
}

namespace hipacc_b200 {
// meaning of one positional argument of hipaccLaunchKernel(kernel, grid, block, ep, timing, smem, args...): the rewriter
// emits them in the order of HipaccKernel::createArgInfo (lib/DSL/ClassRepresentation.cpp:512-597): per image member the
// pointer, _width, _height, [_stride], [_offset_x, _offset_y]; scalars; then bh_start_* / bh_fall_back
enum ArgRole { ARG_PTR, ARG_WIDTH, ARG_HEIGHT, ARG_STRIDE, ARG_OFFSET_X, ARG_OFFSET_Y, ARG_SCALAR, ARG_IGNORED };
struct KernelArg {
    ArgRole role;
    int index;   // ARG_PTR .. ARG_OFFSET_Y: image (0 = the iteration space / output, 1.. = the inputs in declaration order);
                 // ARG_SCALAR: parameter slot (point operators: p[0], p[1]; Harris: 0 = k, 1 = threshold; bilateral: 0 = sigma_r)
};
inline KernelArg arg_ptr(int img) { return {ARG_PTR, img}; }
inline KernelArg arg_width(int img) { return {ARG_WIDTH, img}; }
inline KernelArg arg_height(int img) { return {ARG_HEIGHT, img}; }
inline KernelArg arg_stride(int img) { return {ARG_STRIDE, img}; }
inline KernelArg arg_offset_x(int img) { return {ARG_OFFSET_X, img}; }
inline KernelArg arg_offset_y(int img) { return {ARG_OFFSET_Y, img}; }
inline KernelArg arg_scalar(int slot) { return {ARG_SCALAR, slot}; }
inline KernelArg arg_ignored() { return {ARG_IGNORED, 0}; }   // bh_start_left / right / top / bottom, bh_fall_back

// What rewritten host code names where the reference names a generated __global__ function
struct OperatorKernel {
    enum Kind { LOCAL, POINT, BILATERAL, HARRIS } kind = LOCAL;
    hb_local_desc local{};          // views are filled from the arguments at launch; coef_* / domain point at the "symbols"
    hb_bilateral_desc bilateral{};
    int point_op = 0;               // hb_point_kind
    int interp[3] = {0, 0, 0};
    int dtype[4] = {0, 0, 0, 0};    // pixel type of image 0 (output) and of the inputs
    std::vector<KernelArg> signature;
};
struct ReductionKernel { int mode; };                               // hb_reduce_mode: stands for hipacc_shared_reduction<T, reduceFn>
struct BinningKernel { int index_kind, value_kind; double p0; };    // stands for hipacc_binning_reduction<...>

struct ArgValue { const void *ptr = nullptr; long long i = 0; double f = 0.0; };
template <typename T> ArgValue to_arg(T *p) { ArgValue a; a.ptr = p; return a; }
template <typename T, typename std::enable_if<std::is_integral<T>::value || std::is_enum<T>::value, int>::type = 0> ArgValue to_arg(T v) {
    ArgValue a; a.i = (long long)v; a.f = (double)v; return a;
}
template <typename T, typename std::enable_if<std::is_floating_point<T>::value, int>::type = 0> ArgValue to_arg(T v) {
    ArgValue a; a.f = (double)v; a.i = (long long)v; return a;
}
}  // namespace hipacc_b200

// hipaccLaunchKernel (runtime/hipacc_cu.hpp:255-256, hipacc_cu_standalone.hpp:277-329): grid / block / shared_memory are
// accepted for source compatibility; the pre-built kernels choose their own launch shape.
template <typename... KernelParameters>
void hipaccLaunchKernel(hipacc_b200::OperatorKernel const &kernel_function, dim3 const & /*gridDim*/, dim3 const & /*blockDim*/,
                        HipaccExecutionParameterCuda const &ep, bool print_timing, size_t /*shared_memory*/, KernelParameters &&...parameters) {
    using namespace hipacc_b200;
    const ArgValue vals[] = {to_arg(parameters)...};
    const size_t n = sizeof...(parameters);
    if (n != kernel_function.signature.size()) {
        std::fprintf(stderr, "ERROR: hipaccLaunchKernel(): %zu arguments for a kernel signature of %zu\n", n, kernel_function.signature.size());
        return;
    }
    hb_view img[4];
    bool have[4] = {false, false, false, false}, have_stride[4] = {false, false, false, false};
    double scalar[4] = {0, 0, 0, 0};
    std::memset(img, 0, sizeof(img));
    for (size_t k = 0; k < n; ++k) {
        const KernelArg &a = kernel_function.signature[k];
        if (a.role == ARG_IGNORED) continue;
        if (a.index < 0 || a.index > 3) { std::fprintf(stderr, "ERROR: hipaccLaunchKernel(): bad signature entry %zu\n", k); return; }
        hb_view &v = img[a.index];
        switch (a.role) {
        case ARG_PTR: v.data = const_cast<void *>(vals[k].ptr); have[a.index] = true; break;
        case ARG_WIDTH: v.width = (int)vals[k].i; break;
        case ARG_HEIGHT: v.height = (int)vals[k].i; break;
        case ARG_STRIDE: v.stride = (int)vals[k].i; have_stride[a.index] = true; break;
        case ARG_OFFSET_X: v.offset_x = (int)vals[k].i; break;
        case ARG_OFFSET_Y: v.offset_y = (int)vals[k].i; break;
        case ARG_SCALAR: scalar[a.index] = vals[k].f; break;
        default: break;
        }
    }
    int n_in = 0;
    for (int i = 0; i < 4; ++i) {
        if (!have[i]) continue;
        hb_view &v = img[i];
        v.dtype = kernel_function.dtype[i];
        // the reference passes the accessor's region, not the allocation: the smallest allocation that holds it
        v.img_width = have_stride[i] ? v.stride : v.offset_x + v.width;
        if (!have_stride[i]) v.stride = v.img_width;
        if (v.img_width < v.offset_x + v.width) v.img_width = v.offset_x + v.width;
        v.img_height = v.offset_y + v.height;
        if (i > 0) n_in = i;
    }
    LaunchScope sc(ep, print_timing);
    switch (kernel_function.kind) {
    case OperatorKernel::LOCAL: {
        hb_local_desc d = kernel_function.local;
        d.in = img[1]; d.out = img[0];
        check(hb_local_op(&d, sc.stream()), "hipaccLaunchKernel() [local operator]");
        break;
    }
    case OperatorKernel::BILATERAL: {
        hb_bilateral_desc d = kernel_function.bilateral;
        d.in = img[1]; d.out = img[0];
        if (scalar[0] != 0.0) d.sigma_r = (int)scalar[0];
        check(hb_bilateral(&d, sc.stream()), "hipaccLaunchKernel() [bilateral]");
        break;
    }
    case OperatorKernel::POINT: {
        hb_point_desc d;
        std::memset(&d, 0, sizeof(d));
        d.n_in = n_in;
        for (int i = 0; i < n_in; ++i) { d.in[i] = img[i + 1]; d.interp[i] = kernel_function.interp[i]; }
        d.out = img[0]; d.op = kernel_function.point_op; d.p[0] = scalar[0]; d.p[1] = scalar[1];
        check(hb_point_op(&d, sc.stream()), "hipaccLaunchKernel() [point operator]");
        break;
    }
    case OperatorKernel::HARRIS: {
        hb_harris_desc d;
        std::memset(&d, 0, sizeof(d));
        d.in = img[1]; d.out = img[0]; d.k = (float)scalar[0]; d.threshold = (float)scalar[1];
        check(hb_harris(&d, sc.stream()), "hipaccLaunchKernel() [harris]");
        break;
    }
    }
    sc.done("hipaccLaunchKernel", print_timing);
}

// hipaccApplyReductionShared (runtime/hipacc_cu.hpp:336-347, hipacc_cu.tpp:312-408): max_threads / pixels_per_thread / tex
// are accepted for source compatibility (the reduction kernels size their own grid; no texture path)
template <typename T>
T hipaccApplyReductionShared(const hipacc_b200::ReductionKernel &kernel2D, const HipaccAccessor<T> &acc, unsigned int /*max_threads*/,
                             unsigned int /*pixels_per_thread*/, HipaccExecutionParameterCuda const &ep, const void * /*tex*/ = nullptr,
                             bool print_timing = false) {
    return hipaccApplyReduction<T>(acc, kernel2D.mode, ep, print_timing);
}
template <typename T>
T hipaccApplyReductionShared(const hipacc_b200::ReductionKernel &kernel2D, const HipaccImageCuda<T> &img, unsigned int max_threads,
                             unsigned int pixels_per_thread, HipaccExecutionParameterCuda const &ep, const void *tex = nullptr, bool print_timing = false) {
    return hipaccApplyReductionShared<T>(kernel2D, HipaccAccessor<T>(img), max_threads, pixels_per_thread, ep, tex, print_timing);
}
// hipaccApplyBinningSegmented (runtime/hipacc_cu.hpp:353-360, hipacc_cu.tpp:410-464): T = bin type, T2 = pixel type;
// returns `new T[num_bins]` owned by the caller
template <typename T, typename T2>
T *hipaccApplyBinningSegmented(hipacc_b200::BinningKernel const &kernel2D, const HipaccAccessor<T2> &acc, unsigned int /*num_warps*/,
                               unsigned int /*num_units*/, unsigned int num_bins, HipaccExecutionParameterCuda const &ep, const void * /*tex*/ = nullptr,
                               bool print_timing = false) {
    return hipaccApplyBinning<T2, T>(acc, num_bins, kernel2D.index_kind, kernel2D.value_kind, kernel2D.p0, ep, print_timing);
}

#endif  // HIPACC_B200_RT_HPP
