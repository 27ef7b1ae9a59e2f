// hipacc.hpp -- Hipacc-compatible DSL surface (namespace hipacc) executing on a B200 through the C ABI.
//
// Mirrors the class API of the reference's dsl/ headers (dsl/image.hpp, iterationspace.hpp, mask.hpp,
// kernel.hpp, pyramid.hpp): Image, BoundaryCondition, Accessor, IterationSpace, Mask, Domain,
// Kernel<T>::kernel()/execute()/reduced_data(), Pyramid, traverse -- same constructors, same lifetime rules
// (accessors hold references, execute() is idempotent per Kernel object), same enum values.
//
// What differs, and why: in the reference a Kernel's `kernel()` body is compiled -- by the host compiler in
// DSL mode, by Hipacc's Clang-based rewriter into a CUDA kernel otherwise (lib/AST/ASTTranslate.cpp).  This
// front ships pre-built sm_100a kernels instead of a compiler, so a Kernel subclass states its operator as a
// value: it overrides `lower()` and returns one of the b200:: descriptions below -- exactly the facts
// Hipacc's KernelStatistics / Convolution passes extract from the body (lib/Analysis/KernelStatistics.cpp,
// lib/AST/Convolution.cpp).  `kernel()` stays in the class as the definition of the semantics (it compiles
// against this header, it is never run on the host: there is no CPU fallback).  A Kernel without a lowering,
// or with one the library has no device kernel for, fails loudly.
#ifndef HIPACC_B200_DSL_HPP
#define HIPACC_B200_DSL_HPP

#include <algorithm>
#include <cmath>
#include <initializer_list>

#include "hipacc_rt.hpp"

#ifndef HIPACC_CODEGEN
#define HIPACC_CODEGEN
#endif

namespace hipacc {

enum class Boundary : uint8_t { UNDEFINED = 0, CLAMP, REPEAT, MIRROR, CONSTANT };  // dsl/image.hpp:46-52
enum class Interpolate : uint8_t { NO = 0, NN, LF, B5, CF, L3 };                   // dsl/image.hpp:54-61
enum class Reduce : uint8_t { SUM = 0, MIN, MAX, PROD, MEDIAN };                   // dsl/kernel.hpp:48-54

namespace math {
template <typename T> inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> inline T max(T a, T b) { return b > a ? b : a; }
using std::abs; using std::exp; using std::sqrt;
using ::expf; using ::sqrtf; using ::fabsf;  // the C library's float functions, as in dsl/math_functions.hpp
}  // namespace math

namespace b200 {
[[noreturn]] inline void host_body_called() {
    std::fprintf(stderr, "hipacc_b200: a kernel() body was executed on the host; this front runs operators on the device only "
                         "(override lower(), see include/hipacc_b200/hipacc.hpp)\n");
    std::abort();
}
}  // namespace b200

// ---------------------------------------------------------------------------------------------------
// Image (dsl/image.hpp:64-214): pixels live in HBM; data() reads them back into the internal host mirror
// ---------------------------------------------------------------------------------------------------
template <typename data_t> class Image {
    HipaccImageCuda<data_t> mem_;

  public:
    using pixel_type = data_t;
    Image(const int width, const int height, data_t *init = nullptr, bool /*deep_copy*/ = true)
        : mem_(hipaccCreateMemory<data_t>(init, (size_t)width, (size_t)height)) {}
    explicit Image(const HipaccImageCuda<data_t> &mem) : mem_(mem) {}  // alias (pyramid levels, mapped device memory)
    int width() const { return mem_->get_width(); }
    int height() const { return mem_->get_height(); }
    Image &operator=(data_t *other) { hipaccWriteMemory(mem_, other); return *this; }
    Image &operator=(const Image &other) {
        if (mem_ && other.mem_ && mem_ != other.mem_) hipaccCopyMemory(other.mem_, mem_);
        else mem_ = other.mem_;
        return *this;
    }
    Image(const Image &) = default;
    data_t *data() { return hipaccReadMemory(mem_); }
    const HipaccImageCuda<data_t> &mem() const { return mem_; }
};

// ---------------------------------------------------------------------------------------------------
// Mask / Domain (dsl/mask.hpp): compile-time sized coefficient tables and 0/1 footprints
// ---------------------------------------------------------------------------------------------------
class MaskBase {
  protected:
    int size_x_, size_y_;
    std::vector<uchar> domain_;  // row-major, 1 = visited

  public:
    MaskBase(int size_x, int size_y) : size_x_(size_x), size_y_(size_y), domain_((size_t)size_x * size_y, 1) {
        assert(size_x > 0 && size_y > 0 && "Size for Domain must be positive!");
    }
    virtual ~MaskBase() = default;
    int size_x() const { return size_x_; }
    int size_y() const { return size_y_; }
    const std::vector<uchar> &domain_bits() const { return domain_; }
    int x() const { b200::host_body_called(); }
    int y() const { b200::host_body_called(); }
};

class Domain : public MaskBase {
  public:
    class Setter {
        uchar &ref_;
      public:
        explicit Setter(uchar &r) : ref_(r) {}
        Setter &operator=(const uchar val) { ref_ = val ? 1 : 0; return *this; }
    };
    Domain(const int size_x, const int size_y) : MaskBase(size_x, size_y) {}
    template <int size_y, int size_x> explicit Domain(const uchar (&domain)[size_y][size_x]) : MaskBase(size_x, size_y) {
        for (int y = 0; y < size_y; ++y)
            for (int x = 0; x < size_x; ++x) domain_[(size_t)y * size_x + x] = domain[y][x] ? 1 : 0;
    }
    explicit Domain(const MaskBase &mask) : MaskBase(mask) {}
    // dom(xf, yf) = 0 punches a hole; offsets are relative to the centre (dsl/mask.hpp:186-190)
    Setter operator()(const int xf, const int yf) {
        return Setter(domain_.at((size_t)(yf + size_y_ / 2) * size_x_ + (xf + size_x_ / 2)));
    }
    Domain &operator=(const uchar *other) {
        for (size_t i = 0; i < domain_.size(); ++i) domain_[i] = other[i] ? 1 : 0;
        return *this;
    }
};

template <typename data_t> class Mask : public MaskBase {
    std::vector<data_t> coef_;

    void sync_domain() {  // zero coefficients are Domain holes (Mask ctor, dsl/mask.hpp:238-250)
        for (size_t i = 0; i < coef_.size(); ++i) domain_[i] = coef_[i] != data_t(0) ? 1 : 0;
    }

  public:
    template <int size_y, int size_x> explicit Mask(const data_t (&mask)[size_y][size_x]) : MaskBase(size_x, size_y), coef_((size_t)size_x * size_y) {
        for (int y = 0; y < size_y; ++y)
            for (int x = 0; x < size_x; ++x) coef_[(size_t)y * size_x + x] = mask[y][x];
        sync_domain();
    }
    Mask(int size_x, int size_y) : MaskBase(size_x, size_y), coef_((size_t)size_x * size_y) {}
    Mask &operator=(const data_t *other) {
        std::copy(other, other + coef_.size(), coef_.begin());
        sync_domain();
        return *this;
    }
    const std::vector<data_t> &coefficients() const { return coef_; }
    // kernel()-body forms: mask(), mask(dom), mask(x, y)
    data_t operator()() const { b200::host_body_called(); }
    data_t operator()(const Domain &) const { b200::host_body_called(); }
    data_t operator()(int, int) const { b200::host_body_called(); }
};

// ---------------------------------------------------------------------------------------------------
// BoundaryCondition / Accessor / IterationSpace (dsl/image.hpp:216-720, dsl/iterationspace.hpp)
// ---------------------------------------------------------------------------------------------------
template <typename data_t> class BoundaryCondition {
  public:
    Image<data_t> &img;
    const int size_x, size_y;
    const Boundary mode;
    const data_t const_val;
    BoundaryCondition(Image<data_t> &Img, const int size_x, const int size_y, const Boundary bmode)
        : img(Img), size_x(size_x), size_y(size_y), mode(bmode), const_val() {
        assert(bmode != Boundary::CONSTANT && "Boundary handling set to Constant, but no Constant specified.");
    }
    BoundaryCondition(Image<data_t> &Img, const int size, const Boundary bmode) : BoundaryCondition(Img, size, size, bmode) {}
    BoundaryCondition(Image<data_t> &Img, MaskBase &Mask, const Boundary bmode) : BoundaryCondition(Img, Mask.size_x(), Mask.size_y(), bmode) {}
    BoundaryCondition(Image<data_t> &Img, const int size_x, const int size_y, const Boundary bmode, const data_t val)
        : img(Img), size_x(size_x), size_y(size_y), mode(bmode), const_val(val) {
        assert(bmode == Boundary::CONSTANT && "Constant for boundary handling specified, but boundary mode is different.");
    }
    BoundaryCondition(Image<data_t> &Img, const int size, const Boundary bmode, const data_t val) : BoundaryCondition(Img, size, size, bmode, val) {}
    BoundaryCondition(Image<data_t> &Img, MaskBase &Mask, const Boundary bmode, const data_t val)
        : BoundaryCondition(Img, Mask.size_x(), Mask.size_y(), bmode, val) {}
};

class AccessorBase {
  public:
    virtual ~AccessorBase() = default;
};

template <typename data_t> class Accessor : public AccessorBase {
  public:
    Image<data_t> &img;
    const int width_, height_, offset_x_, offset_y_;
    const Boundary bmode;
    const data_t const_val;
    const Interpolate imode;
    const bool has_bc;

    Accessor(Image<data_t> &Img, const Interpolate imode = Interpolate::NO)
        : img(Img), width_(Img.width()), height_(Img.height()), offset_x_(0), offset_y_(0), bmode(Boundary::CLAMP), const_val(), imode(imode),
          has_bc(false) {}
    Accessor(Image<data_t> &Img, const int width, const int height, const int xf, const int yf, const Interpolate imode = Interpolate::NO)
        : img(Img), width_(width), height_(height), offset_x_(xf), offset_y_(yf), bmode(Boundary::CLAMP), const_val(), imode(imode), has_bc(false) {}
    Accessor(const BoundaryCondition<data_t> &BC, const Interpolate imode = Interpolate::NO)
        : img(BC.img), width_(BC.img.width()), height_(BC.img.height()), offset_x_(0), offset_y_(0), bmode(BC.mode), const_val(BC.const_val),
          imode(imode), has_bc(true) {}
    Accessor(const BoundaryCondition<data_t> &BC, const int width, const int height, const int xf, const int yf,
             const Interpolate imode = Interpolate::NO)
        : img(BC.img), width_(width), height_(height), offset_x_(xf), offset_y_(yf), bmode(BC.mode), const_val(BC.const_val), imode(imode),
          has_bc(true) {}

    int width() const { return width_; }
    int height() const { return height_; }
    HipaccAccessor<data_t> rt() const { return HipaccAccessor<data_t>(img.mem(), (size_t)width_, (size_t)height_, offset_x_, offset_y_); }

    // kernel()-body forms
    data_t &operator()() { b200::host_body_called(); }
    data_t &operator()(const int, const int) { b200::host_body_called(); }
    data_t &operator()(const MaskBase &) { b200::host_body_called(); }
    int x() const { b200::host_body_called(); }
    int y() const { b200::host_body_called(); }
};

template <typename data_t> class IterationSpace {
  public:
    Image<data_t> &img;
    const int width_, height_, offset_x_, offset_y_;
    explicit IterationSpace(Image<data_t> &img) : img(img), width_(img.width()), height_(img.height()), offset_x_(0), offset_y_(0) {}
    IterationSpace(Image<data_t> &img, const int width, const int height) : img(img), width_(width), height_(height), offset_x_(0), offset_y_(0) {}
    IterationSpace(Image<data_t> &img, const int width, const int height, const int offset_x, const int offset_y)
        : img(img), width_(width), height_(height), offset_x_(offset_x), offset_y_(offset_y) {}
    int width() const { return width_; }
    int height() const { return height_; }
    int offset_x() const { return offset_x_; }
    int offset_y() const { return offset_y_; }
    HipaccAccessor<data_t> rt() const { return HipaccAccessor<data_t>(img.mem(), (size_t)width_, (size_t)height_, offset_x_, offset_y_); }
};

// ---------------------------------------------------------------------------------------------------
// b200::Lowering -- the operator a kernel() body denotes, as a value
// ---------------------------------------------------------------------------------------------------
namespace b200 {

struct Epilogue {  // output() = epilogue(acc)   (hb_epilogue)
    int kind = HB_EPI_CAST;
    double p[3] = {0, 0, 0};
};
inline Epilogue cast() { return {}; }                                        // (T)acc
inline Epilogue add_cast(double a) { return {HB_EPI_ADD_CAST, {a, 0, 0}}; }  // (T)(acc + a)
inline Epilogue add_clamp_cast(double a, double lo, double hi) { return {HB_EPI_ADD_CLAMP_CAST, {a, lo, hi}}; }
inline Epilogue div_int_cast(int n) { return {HB_EPI_DIVI_CAST, {(double)n, 0, 0}}; }
inline Epilogue div_float_cast(double n) { return {HB_EPI_DIVF_CAST, {n, 0, 0}}; }

// the body of Kernel::binning(x, y, pixel):  bin(INDEX(pixel)) = VALUE(pixel)   (hb_bin_index / hb_bin_value)
struct Binning {
    int index_kind = -1, value_kind = HB_BIN_VALUE_ONE;
    double p0 = 0.0;
};
inline Binning bin_scaled_count(double divisor) { return {HB_BIN_INDEX_SCALE, HB_BIN_VALUE_ONE, divisor}; }  // bin(pixel/divisor*num_bins) = 1
inline Binning bin_pixel_count() { return {HB_BIN_INDEX_PIXEL, HB_BIN_VALUE_ONE, 0.0}; }                     // bin(pixel) = 1
inline Binning bin_pixel_sum() { return {HB_BIN_INDEX_PIXEL, HB_BIN_VALUE_PIXEL, 0.0}; }                     // bin(pixel) = pixel

struct Lowering {
    enum Kind { NONE, LOCAL, BILATERAL, POINT, HARRIS } kind = NONE;
    std::function<void(const hb_view &out, void *stream)> launch;
};

namespace detail {
// boundary constant as the C ABI carries it (a scalar; vector pixels: the x channel, broadcast)
template <typename T> inline double const_of(const T &v) { return (double)v; }
#ifndef HIPACC_B200_NO_VECTOR_TYPES
inline double const_of(const uchar4 &v) { return (double)v.x; }
#endif
template <typename T> hb_view in_view(const Accessor<T> &a) { return a.rt().view(); }
inline int mode_of(Reduce m) { assert(m != Reduce::MEDIAN && "MEDIAN is not implemented"); return (int)m; }

template <typename TI, typename TM>
Lowering local(int kind, const Accessor<TI> &in, const MaskBase &shape, const std::vector<TM> *coef, Reduce mode, int tap, Epilogue epi, int acc_dtype) {
    auto cf = std::make_shared<std::vector<float>>();
    auto ci = std::make_shared<std::vector<int>>();
    auto dom = std::make_shared<std::vector<uchar>>(shape.domain_bits());
    constexpr bool fmask = std::is_floating_point<TM>::value;
    if (coef) {
        if (fmask) cf->assign(coef->begin(), coef->end());
        else ci->assign(coef->begin(), coef->end());
    }
    if (acc_dtype < 0) acc_dtype = (coef && fmask) || std::is_floating_point<TI>::value ? HB_F32 : HB_S32;
    hb_local_desc d;
    std::memset(&d, 0, sizeof(d));
    d.in = in_view(in);
    d.kind = kind; d.reduce_mode = mode_of(mode); d.tap = tap; d.acc_dtype = acc_dtype;
    d.size_x = shape.size_x(); d.size_y = shape.size_y();
    d.boundary = (int)in.bmode; d.boundary_const = detail::const_of(in.const_val);
    d.epilogue = epi.kind;
    for (int i = 0; i < 3; ++i) d.epi_p[i] = epi.p[i];
    Lowering L;
    L.kind = Lowering::LOCAL;
    L.launch = [d, cf, ci, dom, kind](const hb_view &out, void *stream) mutable {
        d.out = out;
        d.coef_f32 = cf->empty() ? nullptr : cf->data();
        d.coef_s32 = ci->empty() ? nullptr : ci->data();
        d.domain = kind == HB_LOCAL_REDUCE_DOMAIN ? dom->data() : nullptr;
        hipacc_b200::check(hb_local_op(&d, stream), "Kernel::execute() [local operator]");
    };
    return L;
}
}  // namespace detail

// output() = epi(convolve(mask, mode, [&]{ return mask() * in(mask); }))            (dsl/kernel.hpp:241-267)
template <typename TI, typename TM>
Lowering convolve(const Accessor<TI> &in, const Mask<TM> &mask, Reduce mode = Reduce::SUM, Epilogue epi = {}, int acc_dtype = -1) {
    return detail::local<TI, TM>(HB_LOCAL_CONVOLVE, in, mask, &mask.coefficients(), mode, HB_TAP_MUL, epi, acc_dtype);
}
// output() = epi(reduce(dom, mode, [&]{ return mask(dom) * in(dom); }))  -- zero taps are not visited   (dsl/kernel.hpp:270-296)
template <typename TI, typename TM>
Lowering reduce(const Accessor<TI> &in, const Domain &dom, const Mask<TM> &mask, Reduce mode = Reduce::SUM, Epilogue epi = {}, int acc_dtype = -1) {
    return detail::local<TI, TM>(HB_LOCAL_REDUCE_DOMAIN, in, dom, &mask.coefficients(), mode, HB_TAP_MUL, epi, acc_dtype);
}
// output() = epi(reduce(dom, mode, [&]{ return in(dom); }))            (Dilate, Erode, Box_Blur)
template <typename TI> Lowering reduce(const Accessor<TI> &in, const Domain &dom, Reduce mode, Epilogue epi = {}, int acc_dtype = -1) {
    return detail::local<TI, int>(HB_LOCAL_REDUCE_DOMAIN, in, dom, nullptr, mode, HB_TAP_IN, epi, acc_dtype);
}
// the iterate() body of samples-public/3_Preprocessing/Bilateral_Filter/src/main.cpp:63-77
template <typename T> Lowering bilateral(const Accessor<T> &in, const Mask<float> &mask, int sigma_r) {
    auto cf = std::make_shared<std::vector<float>>(mask.coefficients());
    hb_bilateral_desc d;
    std::memset(&d, 0, sizeof(d));
    d.in = detail::in_view(in);
    d.size = mask.size_x(); d.sigma_r = sigma_r; d.boundary = (int)in.bmode; d.boundary_const = detail::const_of(in.const_val);
    Lowering L;
    L.kind = Lowering::BILATERAL;
    L.launch = [d, cf](const hb_view &out, void *stream) mutable {
        d.out = out; d.coef_f32 = cf->data();
        hipacc_b200::check(hb_bilateral(&d, stream), "Kernel::execute() [bilateral]");
    };
    return L;
}
// point operators: output() = f(in0(), in1(), in2())   (hb_point_kind); interpolating accessors allowed
template <typename T> Lowering point(int op, std::initializer_list<const Accessor<T> *> ins, double p0 = 0.0, double p1 = 0.0) {
    hb_point_desc d;
    std::memset(&d, 0, sizeof(d));
    int k = 0;
    for (const Accessor<T> *a : ins) {
        assert(k < 3);
        d.in[k] = detail::in_view(*a);
        d.interp[k] = (int)a->imode;
        ++k;
    }
    d.n_in = k; d.op = op; d.p[0] = p0; d.p[1] = p1;
    Lowering L;
    L.kind = Lowering::POINT;
    L.launch = [d](const hb_view &out, void *stream) mutable {
        d.out = out;
        hipacc_b200::check(hb_point_op(&d, stream), "Kernel::execute() [point operator]");
    };
    return L;
}
// the whole 9-kernel pipeline of samples-public/3_Preprocessing/Harris_Corner/src/main.cpp:230-305 as one operator
inline Lowering harris(const Accessor<uchar> &in, float k, float threshold) {
    hb_harris_desc d;
    std::memset(&d, 0, sizeof(d));
    d.in = detail::in_view(in); d.k = k; d.threshold = threshold;
    Lowering L;
    L.kind = Lowering::HARRIS;
    L.launch = [d](const hb_view &out, void *stream) mutable {
        d.out = out;
        hipacc_b200::check(hb_harris(&d, stream), "Kernel::execute() [harris]");
    };
    return L;
}

}  // namespace b200

// ---------------------------------------------------------------------------------------------------
// Kernel (dsl/kernel.hpp:56-330)
// ---------------------------------------------------------------------------------------------------
template <typename data_t, typename bin_t = data_t> class Kernel {
    IterationSpace<data_t> &iteration_space_;
    std::vector<AccessorBase *> inputs_;
    data_t reduction_result_{};
    bool executed_ = false, reduced_ = false;

    // identify the user's `reduce(left, right)` among {SUM, MIN, MAX, PROD} by evaluating it on probe values;
    // anything else has no device kernel
    int probe_reduce_mode() const {
        const bin_t a1 = (bin_t)2, b1 = (bin_t)3, a2 = (bin_t)5, b2 = (bin_t)4;
        const bin_t r1 = reduce(a1, b1), r2 = reduce(a2, b2), r3 = reduce(b1, a1);
        if (r1 == (bin_t)5 && r2 == (bin_t)9) return HB_REDUCE_SUM;
        if (r1 == (bin_t)2 && r2 == (bin_t)4 && r3 == (bin_t)2) return HB_REDUCE_MIN;
        if (r1 == (bin_t)3 && r2 == (bin_t)5 && r3 == (bin_t)3) return HB_REDUCE_MAX;
        if (r1 == (bin_t)6 && r2 == (bin_t)20) return HB_REDUCE_PROD;
        return -1;
    }

  public:
    explicit Kernel(IterationSpace<data_t> &iteration_space) : iteration_space_(iteration_space) {}
    virtual ~Kernel() = default;
    virtual void kernel() = 0;
    virtual b200::Lowering lower() { return {}; }
    virtual bin_t reduce(bin_t, bin_t) const { assert(false && "No reduce method specified"); return {}; }
    virtual void binning(unsigned int, unsigned int, data_t) { assert(false && "No binning method specified"); }  // dsl/kernel.hpp:91
    virtual b200::Binning lower_binning() { return {}; }
    void add_accessor(AccessorBase *acc) { inputs_.push_back(acc); }

    void execute(const HipaccExecutionParameterCuda &ep = nullptr) {
        if (executed_) return;  // idempotent per Kernel object (dsl/kernel.hpp:95,117)
        b200::Lowering L = lower();
        if (L.kind == b200::Lowering::NONE || !L.launch) {
            std::fprintf(stderr, "ERROR: Kernel::execute(): this Kernel has no lower(); there is no host fallback\n");
            return;
        }
        if (ep) ep->pre_kernel();
        L.launch(iteration_space_.rt().view(), ep ? ep->get_stream() : nullptr);
        if (ep) ep->post_kernel();
        executed_ = true;
    }

    // global reduction of the OUTPUT image over the iteration space (dsl/kernel.hpp:121-161)
    data_t reduced_data() {
        if (!executed_) execute();
        if (!reduced_) {
            const int mode = probe_reduce_mode();
            if (mode < 0) {
                std::fprintf(stderr, "ERROR: Kernel::reduced_data(): reduce() is not SUM/MIN/MAX/PROD; no device kernel, no host fallback\n");
                return reduction_result_;
            }
            reduction_result_ = hipaccApplyReduction<data_t>(iteration_space_.rt(), mode);
            reduced_ = true;
        }
        return reduction_result_;
    }

    // binning over the OUTPUT image (dsl/kernel.hpp:163-199): returns `new bin_t[num_bins]`, owned by the caller
    bin_t *binned_data(const unsigned int num_bins) {
        if (!executed_) execute();
        const b200::Binning b = lower_binning();
        if (b.index_kind < 0 || probe_reduce_mode() != HB_REDUCE_SUM || sizeof(bin_t) != 4) {
            std::fprintf(stderr, "ERROR: Kernel::binned_data(): needs lower_binning() and reduce() == left + right on 32-bit bins; "
                                 "no device kernel otherwise, no host fallback\n");
            return new bin_t[num_bins]();
        }
        num_bins_ = num_bins;
        return hipaccApplyBinning<data_t, bin_t>(iteration_space_.rt(), num_bins, b.index_kind, b.value_kind, b.p0);
    }

  protected:
    unsigned int num_bins_ = 0;
    unsigned int num_bins() const { return num_bins_; }
    bin_t &bin(const unsigned int) { b200::host_body_called(); }
    // kernel()-body vocabulary; compiles, never runs on the host
    data_t &output() { b200::host_body_called(); }
    int x() const { b200::host_body_called(); }
    int y() const { b200::host_body_called(); }
    template <typename M, typename F> auto convolve(M &, Reduce, const F &f) -> decltype(f()) { b200::host_body_called(); }
    template <typename F> auto reduce(Domain &, Reduce, const F &f) -> decltype(f()) { b200::host_body_called(); }
    template <typename F> void iterate(Domain &, const F &) { b200::host_body_called(); }
};

// ---------------------------------------------------------------------------------------------------
// Pyramid / traverse (dsl/pyramid.hpp).  Level 0 aliases the user image like the emitted code
// (runtime/hipacc_cu.tpp:485-486); the DSL's host-mode deep copy is the documented divergence.
// ---------------------------------------------------------------------------------------------------
using PyramidBase = HipaccPyramid;

template <typename data_t> class Pyramid : public PyramidBase {
    std::vector<Image<data_t>> imgs_;

  public:
    Pyramid(Image<data_t> &img, const int depth) : PyramidBase(depth) {
        imgs_.emplace_back(img.mem());
        int w = img.width() / 2, h = img.height() / 2;
        for (int i = 1; i < depth; ++i) {
            assert(w * h > 0 && "Pyramid stages too deep for image size.");
            imgs_.emplace_back(w, h);
            w /= 2; h /= 2;
        }
    }
    Pyramid(Pyramid const &) = delete;
    Pyramid &operator=(Pyramid const &) = delete;
    Image<data_t> &operator()(const int relative) {
        assert(level() + relative >= 0 && level() + relative < (int)imgs_.size() && "Accessed pyramid stage is out of bounds.");
        return imgs_.at(level() + relative);
    }
    void swap(Pyramid<data_t> &other) { imgs_.swap(other.imgs_); }
};

inline void traverse(PyramidBase &p0, const std::function<void()> &f) { hipaccTraverse(p0, f); }
inline void traverse(PyramidBase &p0, PyramidBase &p1, const std::function<void()> &f) { hipaccTraverse(p0, p1, f); }
inline void traverse(PyramidBase &p0, PyramidBase &p1, PyramidBase &p2, const std::function<void()> &f) { hipaccTraverse(p0, p1, p2, f); }
inline void traverse(std::vector<PyramidBase *> const &pyrs, const std::function<void()> &f) { hipaccTraverse(pyrs, f); }
inline void traverse(int loop = 1, const std::function<void()> &f = [] {}) { hipaccTraverse((unsigned)loop, f); }

}  // namespace hipacc

#endif  // HIPACC_B200_DSL_HPP
