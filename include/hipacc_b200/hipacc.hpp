// hipacc.hpp -- Hipacc-compatible DSL surface (namespace hipacc) executing on a B200 through the C ABI.
//
// Mirrors the class API of the reference's dsl/ headers (dsl/image.hpp, iterationspace.hpp, mask.hpp,
// kernel.hpp, pyramid.hpp): Image, BoundaryCondition, Accessor, IterationSpace, Mask, Domain,
// Kernel<T>::kernel()/execute()/reduced_data(), Pyramid, traverse -- same constructors, same lifetime rules
// (accessors hold references, execute() is idempotent per Kernel object), same enum values.
//
// What differs, and why: in the reference a Kernel's `kernel()` body is compiled -- by the host compiler in
// DSL mode, by Hipacc's Clang-based rewriter into a CUDA kernel otherwise (lib/AST/ASTTranslate.cpp).  Here an
// operator reaches the device in one of two ways:
//
//   lower()   (host compiler or nvcc)  the Kernel subclass states its operator as a value -- one of the b200::
//             descriptions below, exactly the facts Hipacc's KernelStatistics / Convolution passes extract from the
//             body (lib/Analysis/KernelStatistics.cpp, lib/AST/Convolution.cpp) -- and execute() launches the
//             library's pre-built, hand-tuned sm_100a kernel for it (the hot path of every BASELINE config);
//   kernel()  (translation unit compiled by nvcc, `-x cu`)  the body itself is compiled for the device: every
//             DSL form it may use (output(), x(), y(), Accessor / Mask / Domain calls, convolve / reduce / iterate,
//             the vector types) is a __host__ __device__ function of this header, and execute() launches
//             dsl_kernel<YourKernel>, one thread per pixel of the iteration space.  Any body runs this way -- no
//             lower() needed -- which is what the reference's ASTTranslate does with a compiler and this header does
//             with templates.  Sample sources compile unmodified: the macro at the end of this header makes the
//             member `kernel()` __host__ __device__ and adds the typed launch hook to the class.
//
// With both present, lower() wins (it is the tuned kernel) and HIPACC_B200_CHECK_LOWERING=1 runs the compiled
// body next to it and aborts on any difference -- a lower() that no longer matches an edited body is caught.
// A body is never run on the HOST: a Kernel with neither a lowering nor a device-compiled body fails loudly.
#ifndef HIPACC_B200_DSL_HPP
#define HIPACC_B200_DSL_HPP

#include <algorithm>
#include <cmath>
#include <initializer_list>
#include <type_traits>

#include "hipacc_rt.hpp"

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define HIPACC_B200_DEVICE_DSL 1
#endif

#ifndef HIPACC_CODEGEN
#define HIPACC_CODEGEN
#endif

namespace hipacc {

enum class Boundary : uint8_t { UNDEFINED = 0, CLAMP, REPEAT, MIRROR, CONSTANT };  // dsl/image.hpp:46-52
enum class Interpolate : uint8_t { NO = 0, NN, LF, B5, CF, L3 };                   // dsl/image.hpp:54-61
enum class Reduce : uint8_t { SUM = 0, MIN, MAX, PROD, MEDIAN };                   // dsl/kernel.hpp:48-54

namespace math {
template <typename T, typename std::enable_if<!hipacc_b200::vec4<T>::is, int>::type = 0> HB_HD T min(T a, T b) { return b < a ? b : a; }
template <typename T, typename std::enable_if<!hipacc_b200::vec4<T>::is, int>::type = 0> HB_HD T max(T a, T b) { return b > a ? b : a; }
using hipacc_b200::vmath::min;   // element-wise on the vector types, vector or scalar bound (dsl/math_functions.hpp)
using hipacc_b200::vmath::max;
using std::abs; using std::exp; using std::sqrt;
using ::expf; using ::sqrtf; using ::fabsf; using ::powf;  // the C library's float functions, as in dsl/math_functions.hpp
#define HB_VEC_FUN1(NAME) HB_HD float4 NAME(float4 v) { return make_float4(::NAME(v.x), ::NAME(v.y), ::NAME(v.z), ::NAME(v.w)); }
HB_VEC_FUN1(sqrtf) HB_VEC_FUN1(expf) HB_VEC_FUN1(fabsf)
#undef HB_VEC_FUN1
}  // namespace math

namespace b200 {
[[noreturn]] inline void host_body_called() {
    std::fprintf(stderr, "hipacc_b200: a kernel() body was executed on the host; this front runs operators on the device only "
                         "(override lower(), or compile the translation unit with nvcc; see include/hipacc_b200/hipacc.hpp)\n");
    std::abort();
}

// ---------------------------------------------------------------------------------------------------
// device side of the compiled-body path (only what device code of this header touches)
// ---------------------------------------------------------------------------------------------------
namespace dev {
constexpr int kSlots = 8;          // Mask / Domain objects one kernel() body may iterate
constexpr int kBlockX = 32, kBlockY = 8;

#ifdef __CUDACC__
// position of this thread in the iteration space
__device__ __forceinline__ int gx() { return (int)(blockIdx.x * kBlockX + threadIdx.x); }
__device__ __forceinline__ int gy() { return (int)(blockIdx.y * kBlockY + threadIdx.y); }
// the current offset of a Mask / Domain iteration (convolve / reduce / iterate set it, mask() / in(mask) read it) is
// per-thread state of an object that all threads share: it lives in shared memory, one cell per (object, thread)
// (two ints per cell, x and y: a 32-bit store read back by a 32-bit load of the same address is forwarded by the compiler
// when the tap loop is unrolled, a packed word read back in halves is not)
__device__ __forceinline__ int2 &iter_state(int slot) {
    __shared__ int2 cells[kSlots][kBlockX * kBlockY];
    return cells[slot][threadIdx.y * kBlockX + threadIdx.x];
}
#endif

#ifdef __CUDACC__
// A device-side field of a snapshot object.  The objects live in the launch's arena in GLOBAL memory and nothing writes
// them during the kernel, but the compiler only sees a generic `this` that might alias the shared-memory cells of
// iter_state: without help it reloads every field after every seek.  __builtin_assume(__isGlobal(p)) tells it the address space: a global load cannot alias a shared-memory store, so the
// fields are loaded once per pixel.
template <typename T> __device__ __forceinline__ T fld(const T &f) {
    __builtin_assume(__isGlobal(&f));   // a plain load the compiler knows to be global: common subexpressions across the taps
    return f;
}
constexpr int kInterpSlots = 4;    // interpolating Accessors one kernel() body may read
// size of the iteration space (the interpolating accessors scale by it), published by every thread of the CTA
__device__ __forceinline__ int *is_dims() {
    __shared__ int dims[2];
    return dims;
}
// an interpolating accessor returns a reference to a computed value (dsl/image.hpp:310-318 keeps it in the Accessor): one
// 16-byte cell per (accessor, thread)
template <typename T> __device__ __forceinline__ T &interp_cell(int slot) {
    static_assert(sizeof(T) <= 16, "pixel type too large");
    __shared__ __align__(16) unsigned char cells[kInterpSlots][kBlockX * kBlockY][16];
    return *reinterpret_cast<T *>(cells[slot][threadIdx.y * kBlockX + threadIdx.x]);
}
#endif
// interpolation weights (dsl/image.hpp:321-383)
HB_HD float w_binomial5(float d) {
    d = d < 0 ? -d : d;
    return d < 0.5f ? 6.0f / 8.0f : d < 1.0f ? 4.0f / 8.0f : d < 1.5f ? 1.0f / 8.0f : 0.0f;
}
HB_HD float w_bicubic(float d) {   // Keys' cubic convolution, a = -0.5
    d = d < 0 ? -d : d;
    const float a = -0.5f;
    if (d < 1.0f) return (a + 2.0f) * d * d * d - (a + 3.0f) * d * d + 1.0f;
    if (d < 2.0f) return a * d * d * d - 5.0f * a * d * d + 8.0f * a * d - 4.0f * a;
    return 0.0f;
}
HB_HD float w_lanczos3(float d) {   // evaluated in double, rounded once
    d = d < 0 ? -d : d;
    const double pi = 3.14159265358979323846;
    if (d == 0.0f) return 1.0f;
    if (d < 3.0f) return (float)(3.0 * (sin(pi * (double)d / 3.0) * sin(pi * (double)d)) / (pi * pi * (double)d * (double)d));
    return 0.0f;
}

// index remapping of a boundary mode on [lo, hi) (dsl/image.hpp:574-612; decisions of SURVEY.md 8c: REPEAT wraps with
// `while`, upper test first) -- the same function the library's tile loaders apply (csrc/hb_common.cuh)
HB_HD int remap(int idx, int lo, int hi, Boundary mode) {
    switch (mode) {
    case Boundary::CLAMP:
        if (idx >= hi) idx = hi - 1;
        if (idx < lo) idx = lo;
        break;
    case Boundary::MIRROR:
        if (idx >= hi) idx = hi - (idx + 1 - hi);
        if (idx < lo) idx = lo + (lo - idx - 1);
        break;
    case Boundary::REPEAT: {
        const int n = hi - lo;
        while (idx >= hi) idx -= n;
        while (idx < lo) idx += n;
        break;
    }
    default: break;
    }
    return idx;
}

// host side: live DSL objects a kernel() body may reference; a launch copies the ones the Kernel object points to
// into device memory and redirects the pointers (see launch_generic)
struct Arena;
struct Obj {
    const void *host;
    size_t size;
    void (*prepare)(const void *host_obj, size_t snapshot_at, Arena &arena);   // fills the device-side fields of the copy at arena.bytes[snapshot_at]
};
inline std::vector<Obj> &registry() { static std::vector<Obj> r; return r; }
inline void enroll(const void *p, size_t n, void (*prep)(const void *, size_t, Arena &)) { registry().push_back(Obj{p, n, prep}); }
inline void retire(const void *p) {
    auto &r = registry();
    for (size_t i = r.size(); i-- > 0;)
        if (r[i].host == p) { r.erase(r.begin() + (long)i); return; }
}
// bytes that go to the device with one launch: object snapshots, coefficient tables, domain bitmaps.  Pointers into the
// arena are kept as offsets (+ 1, so that 0 stays null) in `fixups` until the device address is known.
struct Arena {
    std::vector<unsigned char> bytes;
    std::vector<size_t> fixups;   // offsets (within bytes) of pointer fields holding (arena offset + 1)
    int next_slot = 0, next_interp_slot = 0;
    size_t put(const void *src, size_t n, size_t align = 16) {
        const size_t at = (bytes.size() + align - 1) / align * align;
        bytes.resize(at + n);
        if (src) std::memcpy(bytes.data() + at, src, n);
        return at;
    }
};
struct LaunchCtx {
    void *stream;
    void *out_override;   // non-null: write the result here instead of the iteration space's image (lowering check)
    bool *launched;       // set when a device-compiled body exists and was launched
};
}  // namespace dev
}  // namespace b200

// ---------------------------------------------------------------------------------------------------
// Image (dsl/image.hpp:64-214): pixels live in HBM; data() reads them back into the internal host mirror
// ---------------------------------------------------------------------------------------------------
template <typename data_t> class Accessor;
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
    // img = acc: copy the accessor's region into this image (dsl/image.hpp:150-160; sizes must match)
    Image &operator=(const Accessor<data_t> &other);
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
    std::vector<uchar> domain_;  // row-major, 1 = visited (host)
    // device-side fields, valid in the snapshot a launch makes
    const uchar *d_domain_ = nullptr;
    int d_slot_ = 0;

    static void prepare_base(const MaskBase &m, size_t snap_at, b200::dev::Arena &a) {
        const size_t at = a.put(m.domain_.data(), m.domain_.size(), 16);
        MaskBase &snap = *reinterpret_cast<MaskBase *>(a.bytes.data() + snap_at);   // after put(): it may move the buffer
        snap.d_domain_ = reinterpret_cast<const uchar *>(at + 1);
        a.fixups.push_back(snap_at + (size_t)((const unsigned char *)&snap.d_domain_ - (const unsigned char *)&snap));
        snap.d_slot_ = a.next_slot++;
        if (snap.d_slot_ >= b200::dev::kSlots) {
            std::fprintf(stderr, "hipacc_b200: a kernel() body may reference at most %d Mask / Domain objects\n", b200::dev::kSlots);
            std::abort();
        }
    }
    static void prepare(const void *host_obj, size_t snap_at, b200::dev::Arena &a) {
        prepare_base(*static_cast<const MaskBase *>(host_obj), snap_at, a);
    }

  public:
    MaskBase(int size_x, int size_y) : size_x_(size_x), size_y_(size_y), domain_((size_t)size_x * size_y, 1) {
        assert(size_x > 0 && size_y > 0 && "Size for Domain must be positive!");
    }
    MaskBase(const MaskBase &o) : size_x_(o.size_x_), size_y_(o.size_y_), domain_(o.domain_) {}
    virtual ~MaskBase() = default;
    HB_HD int size_x() const { return size_x_; }
    HB_HD int size_y() const { return size_y_; }
    const std::vector<uchar> &domain_bits() const { return domain_; }
    // offset of the current iteration step relative to the centre (dsl/mask.hpp:100-110)
    HB_HD int x() const {
#ifdef __CUDA_ARCH__
        return b200::dev::iter_state(b200::dev::fld(d_slot_)).x;
#else
        b200::host_body_called();
#endif
    }
    HB_HD int y() const {
#ifdef __CUDA_ARCH__
        return b200::dev::iter_state(b200::dev::fld(d_slot_)).y;
#else
        b200::host_body_called();
#endif
    }
#ifdef __CUDACC__
    __device__ __forceinline__ void dev_seek(int dx, int dy) const { b200::dev::iter_state(b200::dev::fld(d_slot_)) = make_int2(dx, dy); }
    __device__ __forceinline__ bool dev_visited(int k) const { return b200::dev::fld(b200::dev::fld(d_domain_)[k]) != 0; }
    __device__ __forceinline__ int dev_size_x() const { return b200::dev::fld(size_x_); }
    __device__ __forceinline__ int dev_size_y() const { return b200::dev::fld(size_y_); }
#endif
};

class Domain : public MaskBase {
  public:
    class Setter {
        uchar &ref_;
      public:
        explicit Setter(uchar &r) : ref_(r) {}
        Setter &operator=(const uchar val) { ref_ = val ? 1 : 0; return *this; }
    };
    Domain(const int size_x, const int size_y) : MaskBase(size_x, size_y) { b200::dev::enroll(this, sizeof(*this), &MaskBase::prepare); }
    template <int size_y, int size_x> explicit Domain(const uchar (&domain)[size_y][size_x]) : MaskBase(size_x, size_y) {
        for (int y = 0; y < size_y; ++y)
            for (int x = 0; x < size_x; ++x) domain_[(size_t)y * size_x + x] = domain[y][x] ? 1 : 0;
        b200::dev::enroll(this, sizeof(*this), &MaskBase::prepare);
    }
    explicit Domain(const MaskBase &mask) : MaskBase(mask) { b200::dev::enroll(this, sizeof(*this), &MaskBase::prepare); }
    Domain(const Domain &o) : MaskBase(o) { b200::dev::enroll(this, sizeof(*this), &MaskBase::prepare); }
    ~Domain() override { b200::dev::retire(this); }
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
    const data_t *d_coef_ = nullptr;   // device side (snapshot)

    void sync_domain() {  // zero coefficients are Domain holes (Mask ctor, dsl/mask.hpp:238-250)
        for (size_t i = 0; i < coef_.size(); ++i) domain_[i] = coef_[i] != data_t(0) ? 1 : 0;
    }
    static void prepare(const void *host_obj, size_t snap_at, b200::dev::Arena &a) {
        const Mask &m = *static_cast<const Mask *>(host_obj);
        const size_t at = a.put(m.coef_.data(), m.coef_.size() * sizeof(data_t), 16);
        Mask &snap = *reinterpret_cast<Mask *>(a.bytes.data() + snap_at);   // after put(): it may move the buffer
        snap.d_coef_ = reinterpret_cast<const data_t *>(at + 1);
        a.fixups.push_back(snap_at + (size_t)((const unsigned char *)&snap.d_coef_ - (const unsigned char *)&snap));
        MaskBase::prepare_base(m, snap_at, a);
    }

  public:
    template <int size_y, int size_x> explicit Mask(const data_t (&mask)[size_y][size_x]) : MaskBase(size_x, size_y), coef_((size_t)size_x * size_y) {
        for (int y = 0; y < size_y; ++y)
            for (int x = 0; x < size_x; ++x) coef_[(size_t)y * size_x + x] = mask[y][x];
        sync_domain();
        b200::dev::enroll(this, sizeof(*this), &Mask::prepare);
    }
    Mask(int size_x, int size_y) : MaskBase(size_x, size_y), coef_((size_t)size_x * size_y) { b200::dev::enroll(this, sizeof(*this), &Mask::prepare); }
    Mask(const Mask &o) : MaskBase(o), coef_(o.coef_) { b200::dev::enroll(this, sizeof(*this), &Mask::prepare); }
    ~Mask() override { b200::dev::retire(this); }
    // run-time coefficients (dsl/mask.hpp:252-262): the device copy is refreshed at every launch
    Mask &operator=(const data_t *other) {
        std::copy(other, other + coef_.size(), coef_.begin());
        sync_domain();
        return *this;
    }
    const std::vector<data_t> &coefficients() const { return coef_; }
    // kernel()-body forms: mask(), mask(dom), mask(x, y)
    HB_HD data_t operator()() const {
#ifdef __CUDA_ARCH__
        const int sx = dev_size_x(), sy = dev_size_y();
        return b200::dev::fld(b200::dev::fld(d_coef_)[(y() + sy / 2) * sx + x() + sx / 2]);
#else
        b200::host_body_called();
#endif
    }
    HB_HD data_t operator()(const Domain &dom) const {
#ifdef __CUDA_ARCH__
        const int sx = dev_size_x(), sy = dev_size_y();
        return b200::dev::fld(b200::dev::fld(d_coef_)[(dom.y() + sy / 2) * sx + dom.x() + sx / 2]);
#else
        (void)dom; b200::host_body_called();
#endif
    }
    HB_HD data_t operator()(int xf, int yf) const {
#ifdef __CUDA_ARCH__
        const int sx = dev_size_x(), sy = dev_size_y();
        return b200::dev::fld(b200::dev::fld(d_coef_)[(yf + sy / 2) * sx + xf + sx / 2]);
#else
        (void)xf; (void)yf; b200::host_body_called();
#endif
    }
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
    // device-side fields, valid in the snapshot a launch makes
    data_t *d_ptr_ = nullptr;
    int d_stride_ = 0, d_iw_ = 0, d_ih_ = 0, d_islot_ = 0;
    data_t d_const_{};

    static void prepare(const void *host_obj, size_t snap_at, b200::dev::Arena &arena) {
        const Accessor &a = *static_cast<const Accessor *>(host_obj);
        Accessor &snap = *reinterpret_cast<Accessor *>(arena.bytes.data() + snap_at);
        const hb_view &v = a.img.mem()->get_view();
        snap.d_ptr_ = static_cast<data_t *>(v.data);
        snap.d_stride_ = v.stride; snap.d_iw_ = v.img_width; snap.d_ih_ = v.img_height;
        snap.d_const_ = a.const_val;
        if (a.imode != Interpolate::NO) {
            snap.d_islot_ = arena.next_interp_slot++;
            if (snap.d_islot_ >= 4) {
                std::fprintf(stderr, "hipacc_b200: a kernel() body may read at most 4 interpolating Accessors\n");
                std::abort();
            }
        }
    }

  public:
    Image<data_t> &img;
    const int width_, height_, offset_x_, offset_y_;
    const Boundary bmode;
    const data_t const_val;
    const Interpolate imode;
    const bool has_bc;

    Accessor(Image<data_t> &Img, const Interpolate imode = Interpolate::NO)
        : img(Img), width_(Img.width()), height_(Img.height()), offset_x_(0), offset_y_(0), bmode(Boundary::CLAMP), const_val(), imode(imode),
          has_bc(false) { b200::dev::enroll(this, sizeof(*this), &Accessor::prepare); }
    Accessor(Image<data_t> &Img, const int width, const int height, const int xf, const int yf, const Interpolate imode = Interpolate::NO)
        : img(Img), width_(width), height_(height), offset_x_(xf), offset_y_(yf), bmode(Boundary::CLAMP), const_val(), imode(imode), has_bc(false) {
        b200::dev::enroll(this, sizeof(*this), &Accessor::prepare);
    }
    Accessor(const BoundaryCondition<data_t> &BC, const Interpolate imode = Interpolate::NO)
        : img(BC.img), width_(BC.img.width()), height_(BC.img.height()), offset_x_(0), offset_y_(0), bmode(BC.mode), const_val(BC.const_val),
          imode(imode), has_bc(true) { b200::dev::enroll(this, sizeof(*this), &Accessor::prepare); }
    Accessor(const BoundaryCondition<data_t> &BC, const int width, const int height, const int xf, const int yf,
             const Interpolate imode = Interpolate::NO)
        : img(BC.img), width_(width), height_(height), offset_x_(xf), offset_y_(yf), bmode(BC.mode), const_val(BC.const_val), imode(imode),
          has_bc(true) { b200::dev::enroll(this, sizeof(*this), &Accessor::prepare); }
    Accessor(const Accessor &o)
        : img(o.img), width_(o.width_), height_(o.height_), offset_x_(o.offset_x_), offset_y_(o.offset_y_), bmode(o.bmode), const_val(o.const_val),
          imode(o.imode), has_bc(o.has_bc) { b200::dev::enroll(this, sizeof(*this), &Accessor::prepare); }
    ~Accessor() override { b200::dev::retire(this); }

    HB_HD int width() const { return width_; }
    HB_HD int height() const { return height_; }
    HipaccAccessor<data_t> rt() const { return HipaccAccessor<data_t>(img.mem(), (size_t)width_, (size_t)height_, offset_x_, offset_y_); }

    // acc = img / acc = other: region copies on the device (dsl/image.hpp:657-680; sizes must match)
    Accessor &operator=(const Image<data_t> &other) {
        assert(width_ == other.width() && height_ == other.height() && "Size of Accessor and Image have to be the same!");
        hipaccCopyMemoryRegion(HipaccAccessor<data_t>(other.mem()), rt());
        return *this;
    }
    Accessor &operator=(const Accessor &other) {
        assert(width_ == other.width_ && height_ == other.height_ && "Accessor sizes have to be the same!");
        hipaccCopyMemoryRegion(other.rt(), rt());
        return *this;
    }

#ifdef __CUDA_ARCH__
    // image pixel (x, y) through the boundary mode of this accessor's region: pixel_bh of dsl/image.hpp:574-612
    __device__ __forceinline__ data_t &dev_pixel_bh(int x, int y) const {
        using b200::dev::fld;
        const int ox = fld(offset_x_), oy = fld(offset_y_), w = fld(width_), h = fld(height_), iw = fld(d_iw_), ih = fld(d_ih_);
        const Boundary bm = fld(bmode);
        if (bm == Boundary::CONSTANT) {
            if (x < ox || x >= ox + w || y < oy || y >= oy + h) return const_cast<data_t &>(d_const_);
        } else if (bm != Boundary::UNDEFINED) {
            x = b200::dev::remap(x, ox, ox + w, bm);
            y = b200::dev::remap(y, oy, oy + h, bm);
        }
        x = x < 0 ? 0 : x >= iw ? iw - 1 : x;   // memory safety for UNDEFINED / degenerate regions
        y = y < 0 ? 0 : y >= ih ? ih - 1 : y;
        return fld(d_ptr_)[(size_t)y * fld(d_stride_) + x];
    }
    // the value this thread's pixel sees at tap (dx, dy): plain, or through the interpolation mode (dsl/image.hpp:390-528)
    __device__ __forceinline__ data_t &dev_fetch(int dx, int dy) const {
        if (b200::dev::fld(imode) == Interpolate::NO)
            return dev_pixel_bh(b200::dev::fld(offset_x_) + b200::dev::gx() + dx, b200::dev::fld(offset_y_) + b200::dev::gy() + dy);
        return dev_fetch_interp(dx, dy);
    }
    // out of line: the plain path above stays a handful of instructions wherever an accessor call is inlined
    __device__ __noinline__ data_t &dev_fetch_interp(int dx, int dy) const {
        using hipacc_b200::to_float;
        typedef typename hipacc_b200::float_of<data_t>::type F;
        const int *is = b200::dev::is_dims();
        const float stride_x = width_ / (float)is[0], stride_y = height_ / (float)is[1];
        const float x_mapped = offset_x_ + stride_x / 2 + stride_x * (b200::dev::gx() + dx);
        const float y_mapped = offset_y_ + stride_y / 2 + stride_y * (b200::dev::gy() + dy);
        data_t &cell = b200::dev::interp_cell<data_t>(d_islot_);
        if (imode == Interpolate::NN) { cell = dev_pixel_bh((int)x_mapped, (int)y_mapped); return cell; }
        float xb = x_mapped - 0.5f, yb = y_mapped - 0.5f;
        if (xb < 0.0f) xb = 0.0f;
        if (yb < 0.0f) yb = 0.0f;
        const int xi = (int)xb, yi = (int)yb;
        float fx = xb - xi, fy = yb - yi;
        if (imode == Interpolate::LF) {
            const F r = (1.0f - fx) * (1.0f - fy) * to_float(dev_pixel_bh(xi, yi)) + fx * (1.0f - fy) * to_float(dev_pixel_bh(xi + 1, yi)) +
                        (1.0f - fx) * fy * to_float(dev_pixel_bh(xi, yi + 1)) + fx * fy * to_float(dev_pixel_bh(xi + 1, yi + 1));
            cell = hipacc_b200::from_float<data_t>(r);
            return cell;
        }
        // B5 / CF / L3: TAPS x TAPS neighbourhood; L3's rows start at y - 1 and its second row repeats the weight of tap 5
        // (dsl/image.hpp:478,489) -- kept, they decide results
        const bool b5 = imode == Interpolate::B5, cf = imode == Interpolate::CF;
        const int taps = (b5 || cf) ? 4 : 6, x0 = b5 ? 0 : cf ? -1 : -2, y0 = b5 ? 0 : -1;
        if (b5) { fx = (float)((double)fx + 0.5); fy = (float)((double)fy + 0.5); }
        F r{};
        for (int j = 0; j < taps; ++j) {
            F acc{};
            for (int i = 0; i < taps; ++i) {
                const int wi = (!b5 && !cf && j == 1 && i == 4) ? 5 : i;
                const float w = b5 ? b200::dev::w_binomial5(fx - wi) : cf ? b200::dev::w_bicubic(fx - 1 + wi) : b200::dev::w_lanczos3(fx - 2 + wi);
                const F v = to_float(dev_pixel_bh(xi + x0 + i, yi + y0 + j)) * w;
                if (i == 0) acc = v; else acc = acc + v;
            }
            const float wy = b5 ? b200::dev::w_binomial5(fy - j) : cf ? b200::dev::w_bicubic(fy - 1 + j) : b200::dev::w_lanczos3(fy - 2 + j);
            if (j == 0) r = acc * wy; else r = r + acc * wy;
        }
        cell = hipacc_b200::from_float<data_t>(r);
        return cell;
    }
#endif
    // kernel()-body forms
    HB_HD data_t &operator()() {
#ifdef __CUDA_ARCH__
        return dev_fetch(0, 0);
#else
        b200::host_body_called();
#endif
    }
    HB_HD data_t &operator()(const int xf, const int yf) {
#ifdef __CUDA_ARCH__
        return dev_fetch(xf, yf);
#else
        (void)xf; (void)yf; b200::host_body_called();
#endif
    }
    HB_HD data_t &operator()(const MaskBase &m) {
#ifdef __CUDA_ARCH__
        return dev_fetch(m.x(), m.y());
#else
        (void)m; b200::host_body_called();
#endif
    }
    // x and y refer to the area defined by the Accessor (dsl/image.hpp:683-690)
    HB_HD data_t &pixel_at(const int x, const int y) {
#ifdef __CUDA_ARCH__
        return d_ptr_[(size_t)(y + offset_y_) * d_stride_ + x + offset_x_];
#else
        (void)x; (void)y; b200::host_body_called();
#endif
    }
    HB_HD int x() const {   // dsl/image.hpp:692-698
#ifdef __CUDA_ARCH__
        return imode == Interpolate::NO ? b200::dev::gx() : (int)(b200::dev::gx() * width_ / (float)b200::dev::is_dims()[0]);
#else
        b200::host_body_called();
#endif
    }
    HB_HD int y() const {
#ifdef __CUDA_ARCH__
        return imode == Interpolate::NO ? b200::dev::gy() : (int)(b200::dev::gy() * height_ / (float)b200::dev::is_dims()[1]);
#else
        b200::host_body_called();
#endif
    }
};

template <typename data_t> Image<data_t> &Image<data_t>::operator=(const Accessor<data_t> &other) {
    assert(width() == other.width() && height() == other.height() && "Size of Image and Accessor have to be the same!");
    hipaccCopyMemoryRegion(other.rt(), HipaccAccessor<data_t>(mem_));
    return *this;
}

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

// HIPACC_B200_CHECK_LOWERING: integers and bytes must be identical; float pipelines agree within 1e-5 relative (the
// contract of BASELINE.json: the compiled body may contract a * b + c, the library kernels never do)
inline bool pixels_agree(float a, float b) { const float d = a > b ? a - b : b - a, m = (a < 0 ? -a : a) > (b < 0 ? -b : b) ? (a < 0 ? -a : a) : (b < 0 ? -b : b); return a == b || d <= 1e-5f * m + 1e-30f; }
template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0> inline bool pixels_agree(T a, T b) { return a == b; }
template <typename V, hipacc_b200::if_vec<V> = 0> inline bool pixels_agree(const V &a, const V &b) {
    return pixels_agree(a.x, b.x) && pixels_agree(a.y, b.y) && pixels_agree(a.z, b.z) && pixels_agree(a.w, b.w);
}

namespace detail {
// boundary constant as the C ABI carries it (a scalar; vector pixels: the x channel, broadcast)
template <typename T, typename std::enable_if<!hipacc_b200::vec4<T>::is, int>::type = 0> inline double const_of(const T &v) { return (double)v; }
template <typename V, hipacc_b200::if_vec<V> = 0> inline double const_of(const V &v) { return (double)v.x; }
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
template <typename data_t, typename bin_t = data_t> class Kernel;
namespace b200 { namespace dev {
#ifdef HIPACC_B200_DEVICE_DSL
template <size_t N, size_t A> struct alignas(A) Blob { unsigned char b[N]; };
template <class K> __global__ void __launch_bounds__(kBlockX *kBlockY) dsl_kernel(const __grid_constant__ Blob<sizeof(K), alignof(K)> blob, int is_w, int is_h) {
    is_dims()[0] = is_w; is_dims()[1] = is_h;   // every thread writes the same two values
    if (gx() >= is_w || gy() >= is_h) return;
    // the Kernel object as the host built it, references redirected to the device copies of what they point to.  Each
    // thread works on its own copy: the members (above all the redirected references) then live in registers, which lets
    // the compiler see that every tap reads the same Accessor / Mask objects
    alignas(K) unsigned char local[sizeof(K)];
    memcpy(local, blob.b, sizeof(K));
    K &k = *reinterpret_cast<K *>(local);
    k.K::kernel();
}

// Launch dsl_kernel<K> for the host object `k`: every DSL object the Kernel refers to (its Accessor / Mask / Domain
// reference members) is copied into a device arena together with the tables it owns, and the pointers inside the copy
// of `k` are redirected to those copies.  `k` itself travels as the kernel parameter.
template <class K> void launch_generic(K &k, const LaunchCtx &ctx) {
    static_assert(sizeof(K) <= 3072, "Kernel object too large for the kernel parameter block");
    Arena arena;
    Blob<sizeof(K), alignof(K)> blob;
    k.hb_fill_device_fields_(ctx.out_override);
    std::memcpy(blob.b, (const void *)&k, sizeof(K));
    struct Done { const void *host; size_t at; };
    std::vector<Done> done;
    std::vector<size_t> blob_fixups;   // byte offsets in blob.b holding (arena offset + 1)
    const std::vector<Obj> objs = registry();
    for (size_t off = 0; off + sizeof(void *) <= sizeof(K); off += sizeof(void *)) {
        uintptr_t w;
        std::memcpy(&w, blob.b + off, sizeof(w));
        if (!w) continue;
        for (const Obj &o : objs) {
            const uintptr_t lo = (uintptr_t)o.host;
            if (w < lo || w >= lo + o.size) continue;
            size_t at = (size_t)-1;
            for (const Done &d : done) if (d.host == o.host) at = d.at;
            if (at == (size_t)-1) {
                at = arena.put(o.host, o.size, 16);
                done.push_back(Done{o.host, at});
                o.prepare(o.host, at, arena);
            }
            const uintptr_t v = (uintptr_t)(at + (w - lo) + 1);
            std::memcpy(blob.b + off, &v, sizeof(v));
            blob_fixups.push_back(off);
            break;
        }
    }
    cudaStream_t s = (cudaStream_t)ctx.stream;
    unsigned char *d_arena = nullptr;
    auto ck = [](cudaError_t e, const char *what) {
        if (e != cudaSuccess) std::fprintf(stderr, "ERROR: %s: %s\n", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    if (!arena.bytes.empty()) {
        if (!ck(cudaMallocAsync((void **)&d_arena, arena.bytes.size(), s), "Kernel::execute() [compiled body]: cudaMallocAsync")) return;
        for (size_t f : arena.fixups) {
            uintptr_t v;
            std::memcpy(&v, arena.bytes.data() + f, sizeof(v));
            v = (uintptr_t)d_arena + (v - 1);
            std::memcpy(arena.bytes.data() + f, &v, sizeof(v));
        }
        for (size_t f : blob_fixups) {
            uintptr_t v;
            std::memcpy(&v, blob.b + f, sizeof(v));
            v = (uintptr_t)d_arena + (v - 1);
            std::memcpy(blob.b + f, &v, sizeof(v));
        }
        ck(cudaMemcpyAsync(d_arena, arena.bytes.data(), arena.bytes.size(), cudaMemcpyHostToDevice, s), "Kernel::execute() [compiled body]: upload");
    }
    const int is_w = k.hb_is_width_(), is_h = k.hb_is_height_();
    const dim3 block(kBlockX, kBlockY), grid((is_w + kBlockX - 1) / kBlockX, (is_h + kBlockY - 1) / kBlockY);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    const bool timed = hipacc_b200::timing_enabled() && cap == cudaStreamCaptureStatusNone;
    if (timed) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s); }
    dsl_kernel<K><<<grid, block, 0, s>>>(blob, is_w, is_h);
    ck(cudaGetLastError(), "Kernel::execute() [compiled body]: launch");
    if (timed) {
        cudaEventRecord(e1, s);
        ck(cudaEventSynchronize(e1), "Kernel::execute() [compiled body]");
        cudaEventElapsedTime(&hipacc_b200::compiled_body_ms(), e0, e1);
        hipacc_b200::compiled_body_was_last() = true;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    if (d_arena) cudaFreeAsync(d_arena, s);
    if (ctx.launched) *ctx.launched = true;
}

// ---- compiled reduce() / binning() bodies (dsl/kernel.hpp:121-199) over the OUTPUT image of the iteration space -----------
template <class K> struct kernel_types;   // data_t / bin_t of a Kernel subclass
template <class D, class B> D data_type_of(const Kernel<D, B> *);
template <class D, class B> B bin_type_of(const Kernel<D, B> *);

// one partial per CTA: tree fold of the CTA's pixels with K::reduce; the host folds the partials in CTA order
template <class K, class D> __global__ void __launch_bounds__(kBlockX *kBlockY) dsl_reduce_kernel(const __grid_constant__ Blob<sizeof(K), alignof(K)> blob,
                                                                                                 int is_w, int is_h, D *partials) {
    __shared__ __align__(16) unsigned char raw[kBlockX * kBlockY * sizeof(D)];
    __shared__ unsigned char valid[kBlockX * kBlockY];
    D *vals = reinterpret_cast<D *>(raw);
    const K &k = *reinterpret_cast<const K *>(blob.b);
    const int tid = threadIdx.y * kBlockX + threadIdx.x;
    const bool in = gx() < is_w && gy() < is_h;
    valid[tid] = in;
    if (in) vals[tid] = const_cast<K &>(k).hb_output_();
    __syncthreads();
    for (int stride = kBlockX * kBlockY / 2; stride > 0; stride >>= 1) {
        if (tid < stride && valid[tid + stride]) {
            vals[tid] = valid[tid] ? k.K::reduce(vals[tid], vals[tid + stride]) : vals[tid + stride];
            valid[tid] = 1;
        }
        __syncthreads();
    }
    if (tid == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = vals[0];
}

// bins[idx] = reduce(bins[idx], value) for every pixel's `bin(idx) = value`.  reduce() is arbitrary code, not an atomic
// instruction, so a bin is updated by a compare-and-swap loop on its bit pattern (4- and 8-byte bins; no locks, nothing can
// dead-lock).  Each CTA folds into a private copy of the bins in shared memory first (when they fit) and merges the bins it
// touched into the global ones at the end.  As in the reference's GPU runtime (runtime/hipacc_cu_red.hpp:527-641) reduce()
// must be associative and commutative and must accept bin_t() as the start value of every partial fold.
template <class B> struct bin_word { typedef typename std::conditional<sizeof(B) == 8, unsigned long long, unsigned int>::type type; };
template <class K, class B> __device__ __forceinline__ void bin_update(const K &k, B *slot, const B &val) {
    typedef typename bin_word<B>::type W;
    W *w = reinterpret_cast<W *>(slot);
    W old = *reinterpret_cast<volatile W *>(w), assumed;
    do {
        assumed = old;
        B cur;
        memcpy(&cur, &assumed, sizeof(B));
        const B next = k.K::reduce(cur, val);
        W nw;
        memcpy(&nw, &next, sizeof(B));
        old = atomicCAS(w, assumed, nw);
    } while (old != assumed);
}
template <class K, class B> __global__ void __launch_bounds__(kBlockX *kBlockY) dsl_binning_kernel(const __grid_constant__ Blob<sizeof(K), alignof(K)> blob,
                                                                                                  int is_w, int is_h, unsigned int num_bins, B *bins,
                                                                                                  int private_bins) {
    extern __shared__ __align__(16) unsigned char hb_bins_raw[];
    B *sbins = reinterpret_cast<B *>(hb_bins_raw);
    const int tid = threadIdx.y * kBlockX + threadIdx.x;
    is_dims()[0] = is_w; is_dims()[1] = is_h;
    if (private_bins) {
        for (unsigned int b = tid; b < num_bins; b += kBlockX * kBlockY) sbins[b] = B();
        __syncthreads();
    }
    alignas(K) unsigned char local[sizeof(K)];   // binning() stores into the Kernel object: each thread works on its own copy
    memcpy(local, blob.b, sizeof(K));
    K &k = *reinterpret_cast<K *>(local);
    if (gx() < is_w && gy() < is_h) {
        k.K::binning((unsigned int)gx(), (unsigned int)gy(), k.hb_output_());
        const unsigned int idx = k.hb_bin_idx_();
        if (idx < num_bins) bin_update<K, B>(k, (private_bins ? sbins : bins) + idx, k.hb_bin_val_());   // out of range: the reference asserts
    }
    if (private_bins) {
        __syncthreads();
        typedef typename bin_word<B>::type W;
        const B zero = B();
        W zw;
        memcpy(&zw, &zero, sizeof(B));
        for (unsigned int b = tid; b < num_bins; b += kBlockX * kBlockY)
            if (reinterpret_cast<const W *>(sbins)[b] != zw) bin_update<K, B>(k, bins + b, sbins[b]);
    }
}

template <class K> void launch_global(K &k, const LaunchCtx &ctx, int what, unsigned int num_bins, void *result) {
    typedef decltype(data_type_of(&k)) D;
    typedef decltype(bin_type_of(&k)) B;
    Blob<sizeof(K), alignof(K)> blob;
    k.hb_fill_device_fields_(nullptr);
    std::memcpy(blob.b, (const void *)&k, sizeof(K));   // reduce() / binning() see the Kernel's scalar members; no Accessor is read
    const int is_w = k.hb_is_width_(), is_h = k.hb_is_height_();
    const dim3 block(kBlockX, kBlockY), grid((is_w + kBlockX - 1) / kBlockX, (is_h + kBlockY - 1) / kBlockY);
    cudaStream_t s = (cudaStream_t)ctx.stream;
    auto ck = [](cudaError_t e, const char *what_) {
        if (e != cudaSuccess) std::fprintf(stderr, "ERROR: %s: %s\n", what_, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    if (what == 0) {
        if constexpr (std::is_same<D, B>::value) {
            const size_t n = (size_t)grid.x * grid.y;
            D *d_part = nullptr;
            if (!ck(cudaMalloc((void **)&d_part, n * sizeof(D)), "Kernel::reduced_data() [compiled body]")) return;
            dsl_reduce_kernel<K, D><<<grid, block, 0, s>>>(blob, is_w, is_h, d_part);
            std::vector<D> part(n);
            ck(cudaMemcpyAsync(part.data(), d_part, n * sizeof(D), cudaMemcpyDeviceToHost, s), "Kernel::reduced_data() [compiled body]");
            ck(cudaStreamSynchronize(s), "Kernel::reduced_data() [compiled body]");
            cudaFree(d_part);
            D r = part[0];
            for (size_t i = 1; i < n; ++i) r = k.K::reduce(r, part[i]);
            *static_cast<D *>(result) = r;
        } else {
            std::fprintf(stderr, "ERROR: Kernel::reduced_data(): reduce() works on bin_t, which differs from the pixel type\n");
            return;
        }
    } else {
        if constexpr (sizeof(B) == 4 || sizeof(B) == 8) {
            B *d_bins = nullptr;
            if (!ck(cudaMalloc((void **)&d_bins, num_bins * sizeof(B)), "Kernel::binned_data() [compiled body]")) return;
            ck(cudaMemcpyAsync(d_bins, result, num_bins * sizeof(B), cudaMemcpyHostToDevice, s), "Kernel::binned_data() [compiled body]");   // bin_t() each
            const size_t smem = (size_t)num_bins * sizeof(B);
            const int priv = smem <= 32 * 1024;
            dsl_binning_kernel<K, B><<<grid, block, priv ? smem : 0, s>>>(blob, is_w, is_h, num_bins, d_bins, priv);
            ck(cudaGetLastError(), "Kernel::binned_data() [compiled body]: launch");
            ck(cudaMemcpyAsync(result, d_bins, num_bins * sizeof(B), cudaMemcpyDeviceToHost, s), "Kernel::binned_data() [compiled body]");
            ck(cudaStreamSynchronize(s), "Kernel::binned_data() [compiled body]");
            cudaFree(d_bins);
        } else {
            std::fprintf(stderr, "ERROR: Kernel::binned_data(): compiled binning supports 4- and 8-byte bin types\n");
            return;
        }
    }
    if (ctx.launched) *ctx.launched = true;
}
#endif
}}  // namespace b200::dev

template <typename data_t, typename bin_t> class Kernel {
    IterationSpace<data_t> &iteration_space_;
    std::vector<AccessorBase *> inputs_;
    data_t reduction_result_{};
    bool executed_ = false, reduced_ = false;
    // device-side fields of the compiled-body path (filled right before the object is copied into the launch)
    data_t *d_out_ = nullptr;
    int d_out_stride_ = 0, d_ox_ = 0, d_oy_ = 0;
    // what the last `bin(idx) = value` of a binning() body stored (dsl/kernel.hpp:62-64,126-129)
    bin_t bin_val_{};
    unsigned int bin_idx_ = 0;

    // identify the user's `reduce(left, right)` among {SUM, MIN, MAX, PROD} by evaluating it on probe values;
    // anything else has no device kernel
    int probe_reduce_mode() const {
        if constexpr (std::is_arithmetic<bin_t>::value) {
            const bin_t a1 = (bin_t)2, b1 = (bin_t)3, a2 = (bin_t)5, b2 = (bin_t)4;
            const bin_t r1 = reduce(a1, b1), r2 = reduce(a2, b2), r3 = reduce(b1, a1);
            if (r1 == (bin_t)5 && r2 == (bin_t)9) return HB_REDUCE_SUM;
            if (r1 == (bin_t)2 && r2 == (bin_t)4 && r3 == (bin_t)2) return HB_REDUCE_MIN;
            if (r1 == (bin_t)3 && r2 == (bin_t)5 && r3 == (bin_t)3) return HB_REDUCE_MAX;
            if (r1 == (bin_t)6 && r2 == (bin_t)20) return HB_REDUCE_PROD;
        }
        return -1;   // vector bins, or a reduce() that is none of the four
    }

    static bool check_requested() {
        static int on = -1;
        if (on < 0) { const char *e = std::getenv("HIPACC_B200_CHECK_LOWERING"); on = (e && std::atoi(e)) ? 1 : 0; }
        return on == 1;
    }
    // HIPACC_B200_CHECK_LOWERING=1: run the compiled body into a scratch image and compare it with what lower() produced
    void check_lowering(void *stream) {
        const hb_view out = iteration_space_.rt().view();
        hb_view scratch{};
        hipacc_b200::check(hb_image_create(out.dtype, out.img_width, out.img_height, 0, &scratch), "check_lowering: scratch image");
        if (scratch.stride != out.stride) {   // the body indexes with the iteration space's stride
            std::fprintf(stderr, "hipacc_b200: HIPACC_B200_CHECK_LOWERING needs library-allocated images\n");
            hb_image_destroy(&scratch);
            return;
        }
        bool launched = false;
        hb_dispatch_(b200::dev::LaunchCtx{stream, scratch.data, &launched});
        if (!launched) {
            std::fprintf(stderr, "hipacc_b200: HIPACC_B200_CHECK_LOWERING needs the translation unit compiled by nvcc (no device-compiled body here)\n");
            hb_image_destroy(&scratch);
            return;
        }
        std::vector<data_t> a((size_t)out.img_width * out.img_height), b(a.size());
        hipacc_b200::check(hb_image_read(&out, a.data(), stream), "check_lowering: read");
        hipacc_b200::check(hb_image_read(&scratch, b.data(), stream), "check_lowering: read");
        hb_image_destroy(&scratch);
        long bad = 0, first = -1;
        for (int y = 0; y < out.height; ++y)
            for (int x = 0; x < out.width; ++x) {
                const size_t i = (size_t)(y + out.offset_y) * out.img_width + x + out.offset_x;
                if (!b200::pixels_agree(a[i], b[i])) { if (!bad) first = (long)i; ++bad; }
            }
        if (bad) {
            std::fprintf(stderr, "hipacc_b200: lower() and kernel() DISAGREE on %ld of %ld pixels (first at index %ld): the lowering does not "
                                 "describe this body\n", bad, (long)out.width * out.height, first);
            std::abort();
        }
    }

  public:
    explicit Kernel(IterationSpace<data_t> &iteration_space) : iteration_space_(iteration_space) {}
    virtual ~Kernel() = default;
    HB_HD virtual void kernel() = 0;
    virtual b200::Lowering lower() { return {}; }
    HB_HD virtual bin_t reduce(bin_t, bin_t) const { assert(false && "No reduce method specified"); return {}; }
    HB_HD virtual void binning(unsigned int, unsigned int, data_t) { assert(false && "No binning method specified"); }  // dsl/kernel.hpp:91
    virtual b200::Binning lower_binning() { return {}; }
    void add_accessor(AccessorBase *acc) { inputs_.push_back(acc); }

    // the typed launch hook of the compiled-body path: the macro at the end of this header overrides it in every Kernel
    // subclass that an nvcc-compiled translation unit defines.  false = this class has no device-compiled body.
    virtual void hb_dispatch_(const b200::dev::LaunchCtx &c) { if (c.launched) *c.launched = false; }
    // compiled reduce() / binning() bodies (opt-in, -DHIPACC_B200_DEVICE_GLOBAL_OPS): same mechanism, hooked by the macros
    // for `binning(...)`; results through c.result
    virtual void hb_dispatch_global_(const b200::dev::LaunchCtx &c, int /*what*/, unsigned int /*num_bins*/, void * /*result*/) { if (c.launched) *c.launched = false; }
    HB_HD unsigned int hb_bin_idx_() const { return bin_idx_; }
    HB_HD const bin_t &hb_bin_val_() const { return bin_val_; }
    void hb_set_num_bins_(unsigned int n) { num_bins_ = n; }
    void hb_fill_device_fields_(void *out_override) {
        const hb_view v = iteration_space_.rt().view();
        d_out_ = static_cast<data_t *>(out_override ? out_override : v.data);
        d_out_stride_ = v.stride; d_ox_ = v.offset_x; d_oy_ = v.offset_y;
    }
    HB_HD data_t &hb_output_() { return output(); }
    int hb_is_width_() const { return iteration_space_.width(); }
    int hb_is_height_() const { return iteration_space_.height(); }

    template <typename T_EP> void execute(T_EP &&ep) { execute_on(ep); }
    void execute() { execute_on(nullptr); }
    void execute_on(const HipaccExecutionParameterCuda &ep) {
        if (executed_) return;  // idempotent per Kernel object (dsl/kernel.hpp:95,117)
        void *stream = ep ? ep->get_stream() : nullptr;
        b200::Lowering L = lower();
        if (ep) ep->pre_kernel();
        if (L.kind != b200::Lowering::NONE && L.launch) {
            hipacc_b200::compiled_body_was_last() = false;
            L.launch(iteration_space_.rt().view(), stream);
            if (check_requested()) check_lowering(stream);
        } else if (bool launched = false; hb_dispatch_(b200::dev::LaunchCtx{stream, nullptr, &launched}), !launched) {
            std::fprintf(stderr, "ERROR: Kernel::execute(): this Kernel has neither a lower() nor a device-compiled kernel() body "
                                 "(compile the translation unit with nvcc); there is no host fallback\n");
            return;
        }
        if (ep) ep->post_kernel();
        executed_ = true;
    }

    // global reduction of the OUTPUT image over the iteration space (dsl/kernel.hpp:121-161)
    data_t reduced_data() {
        if (!executed_) execute();
        if (!reduced_) {
            const int mode = probe_reduce_mode();
            if (mode < 0) {
                // not one of the library's four folds: the reduce() body itself, compiled for the device
                bool launched = false;
                hb_dispatch_global_(b200::dev::LaunchCtx{nullptr, nullptr, &launched}, 0, 0, &reduction_result_);
                if (!launched) {
                    std::fprintf(stderr, "ERROR: Kernel::reduced_data(): reduce() is not SUM/MIN/MAX/PROD and was not compiled for the device "
                                         "(nvcc, -DHIPACC_B200_DEVICE_GLOBAL_OPS); no host fallback\n");
                    return reduction_result_;
                }
            } else if constexpr (std::is_arithmetic<data_t>::value) {
                reduction_result_ = hipaccApplyReduction<data_t>(iteration_space_.rt(), mode);
            }
            reduced_ = true;
        }
        return reduction_result_;
    }

    // binning over the OUTPUT image (dsl/kernel.hpp:163-199): returns `new bin_t[num_bins]`, owned by the caller
    bin_t *binned_data(const unsigned int num_bins) {
        if (!executed_) execute();
        const b200::Binning b = lower_binning();
        num_bins_ = num_bins;
        if constexpr (sizeof(bin_t) == 4 && std::is_arithmetic<data_t>::value) {
            if (b.index_kind >= 0 && probe_reduce_mode() == HB_REDUCE_SUM)
                return hipaccApplyBinning<data_t, bin_t>(iteration_space_.rt(), num_bins, b.index_kind, b.value_kind, b.p0);
        }
        // any other binning() / reduce() pair: the bodies themselves, compiled for the device
        bin_t *bins = new bin_t[num_bins]();
        bool launched = false;
        hb_dispatch_global_(b200::dev::LaunchCtx{nullptr, nullptr, &launched}, 1, num_bins, bins);
        if (!launched)
            std::fprintf(stderr, "ERROR: Kernel::binned_data(): needs lower_binning() with reduce() == left + right on 32-bit bins, or binning() / "
                                 "reduce() compiled for the device (nvcc, -DHIPACC_B200_DEVICE_GLOBAL_OPS); no host fallback\n");
        return bins;
    }

  protected:
    unsigned int num_bins_ = 0;
    HB_HD unsigned int num_bins() const { return num_bins_; }
    HB_HD bin_t &bin(const unsigned int idx) {
        bin_idx_ = idx;
        return bin_val_;
    }

    // ---- kernel()-body vocabulary: device code under nvcc, never run on the host -------------------------------------
    HB_HD data_t &output() {
#ifdef __CUDA_ARCH__
        return d_out_[(size_t)(d_oy_ + b200::dev::gy()) * d_out_stride_ + d_ox_ + b200::dev::gx()];
#else
        b200::host_body_called();
#endif
    }
    // x and y refer to the area defined by the iteration space (dsl/kernel.hpp:131-133)
    HB_HD data_t &output_at(const int xf, const int yf) {
#ifdef __CUDA_ARCH__
        return d_out_[(size_t)(d_oy_ + yf) * d_out_stride_ + d_ox_ + xf];
#else
        (void)xf; (void)yf; b200::host_body_called();
#endif
    }
    HB_HD int x() const {
#ifdef __CUDA_ARCH__
        return b200::dev::gx();
#else
        b200::host_body_called();
#endif
    }
    HB_HD int y() const {
#ifdef __CUDA_ARCH__
        return b200::dev::gy();
#else
        b200::host_body_called();
#endif
    }
    // dsl/kernel.hpp:241-267: row-major over the mask, the first tap initialises, every later one folds by `mode`.
    // The mask's size is a run-time member, but 3x3 / 5x5 / 7x7 cover almost every program: those sizes take a fully unrolled
    // copy of the loop (a warp-uniform switch), where the tap offsets are constants and the compiler keeps the accessor's and
    // the mask's fields in registers across the taps -- 4x fewer instructions per tap than the generic loop.
    template <typename data_m, typename F> HB_HD auto convolve(Mask<data_m> &mask, Reduce mode, const F &fun) -> decltype(fun()) {
#ifdef __CUDA_ARCH__
        const int sx = mask.dev_size_x(), sy = mask.dev_size_y();
        if (sx == 3 && sy == 3) return convolve_n_<1, 1>(mask, mode, fun);
        if (sx == 5 && sy == 5) return convolve_n_<2, 2>(mask, mode, fun);
        if (sx == 7 && sy == 7) return convolve_n_<3, 3>(mask, mode, fun);
        const int hx = sx / 2, hy = sy / 2;
        mask.dev_seek(-hx, -hy);
        auto result = fun();
        for (int dy = -hy; dy <= hy; ++dy)
            for (int dx = (dy == -hy ? -hx + 1 : -hx); dx <= hx; ++dx) {
                mask.dev_seek(dx, dy);
                fold_(result, fun(), mode);
            }
        return result;
#else
        (void)mask; (void)mode; (void)fun; b200::host_body_called();
#endif
    }
    // dsl/kernel.hpp:270-296 with the Domain's holes skipped (dsl/mask.hpp:112-126)
    template <typename F> HB_HD auto reduce(Domain &dom, Reduce mode, const F &fun) -> decltype(fun()) {
#ifdef __CUDA_ARCH__
        const int sx = dom.dev_size_x(), sy = dom.dev_size_y();
        if (sx == 3 && sy == 3) return reduce_n_<1, 1>(dom, mode, fun);
        if (sx == 5 && sy == 5) return reduce_n_<2, 2>(dom, mode, fun);
        if (sx == 7 && sy == 7) return reduce_n_<3, 3>(dom, mode, fun);
        const int hx = sx / 2, hy = sy / 2;
        decltype(fun()) result{};
        bool first = true;
        int k = 0;
        for (int dy = -hy; dy <= hy; ++dy)
            for (int dx = -hx; dx <= hx; ++dx, ++k) {
                if (!dom.dev_visited(k)) continue;
                dom.dev_seek(dx, dy);
                if (first) { result = fun(); first = false; }
                else fold_(result, fun(), mode);
            }
        return result;
#else
        (void)dom; (void)mode; (void)fun; b200::host_body_called();
#endif
    }
    template <typename F> HB_HD void iterate(Domain &dom, const F &fun) {
#ifdef __CUDA_ARCH__
        const int sx = dom.dev_size_x(), sy = dom.dev_size_y();
        if (sx == 3 && sy == 3) return iterate_n_<1, 1>(dom, fun);
        if (sx == 5 && sy == 5) return iterate_n_<2, 2>(dom, fun);
        if (sx == 7 && sy == 7) return iterate_n_<3, 3>(dom, fun);
        const int hx = sx / 2, hy = sy / 2;
        int k = 0;
        for (int dy = -hy; dy <= hy; ++dy)
            for (int dx = -hx; dx <= hx; ++dx, ++k) {
                if (!dom.dev_visited(k)) continue;
                dom.dev_seek(dx, dy);
                fun();
            }
#else
        (void)dom; (void)fun; b200::host_body_called();
#endif
    }

  private:
#ifdef __CUDACC__
    template <int HX, int HY, typename data_m, typename F> __device__ __forceinline__ auto convolve_n_(Mask<data_m> &mask, Reduce mode, const F &fun) -> decltype(fun()) {
        mask.dev_seek(-HX, -HY);
        auto result = fun();
#pragma unroll
        for (int dy = -HY; dy <= HY; ++dy)
#pragma unroll
            for (int dx = -HX; dx <= HX; ++dx) {
                if (dy == -HY && dx == -HX) continue;
                mask.dev_seek(dx, dy);
                fold_(result, fun(), mode);
            }
        return result;
    }
    template <int HX, int HY, typename F> __device__ __forceinline__ auto reduce_n_(Domain &dom, Reduce mode, const F &fun) -> decltype(fun()) {
        decltype(fun()) result{};
        bool first = true;
#pragma unroll
        for (int dy = -HY; dy <= HY; ++dy)
#pragma unroll
            for (int dx = -HX; dx <= HX; ++dx) {
                if (!dom.dev_visited((dy + HY) * (2 * HX + 1) + dx + HX)) continue;
                dom.dev_seek(dx, dy);
                if (first) { result = fun(); first = false; }
                else fold_(result, fun(), mode);
            }
        return result;
    }
    template <int HX, int HY, typename F> __device__ __forceinline__ void iterate_n_(Domain &dom, const F &fun) {
#pragma unroll
        for (int dy = -HY; dy <= HY; ++dy)
#pragma unroll
            for (int dx = -HX; dx <= HX; ++dx) {
                if (!dom.dev_visited((dy + HY) * (2 * HX + 1) + dx + HX)) continue;
                dom.dev_seek(dx, dy);
                fun();
            }
    }
#endif
    template <typename R> HB_HD static void fold_(R &result, const R &v, Reduce mode) {
        switch (mode) {
        case Reduce::SUM: result += v; break;
        case Reduce::MIN: result = hipacc::math::min(v, result); break;
        case Reduce::MAX: result = hipacc::math::max(v, result); break;
        case Reduce::PROD: result *= v; break;
        default: break;
        }
    }
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

// ---------------------------------------------------------------------------------------------------
// The compiled-body hook.  Under nvcc every `void kernel() { ... }` a Kernel subclass defines after this point becomes
//     void hb_dispatch_(ctx) override { launch dsl_kernel<ThisClass> }        <- typed launch of this very class
//     __host__ __device__ void kernel() { ... }                               <- the body, now device code
// so that sample sources compile unmodified (dsl/kernel.hpp:79 declares `virtual void kernel() = 0`).
// ---------------------------------------------------------------------------------------------------
#ifdef HIPACC_B200_DEVICE_DSL
// `kernel()` with no arguments is the member declaration; `kernel(args)` -- a variable that happens to be called kernel,
// as in samples-public/5_Other/Game_of_Life -- is left alone.
#define kernel(...) HB_KERNEL_##__VA_OPT__(ARGS)(__VA_ARGS__)
#define HB_KERNEL_ARGS(...) kernel(__VA_ARGS__)
#define HB_KERNEL_()                                                                   \
    hb_dispatch_(const ::hipacc::b200::dev::LaunchCtx &hb_ctx_) override {             \
        ::hipacc::b200::dev::launch_generic(*this, hb_ctx_);                           \
    }                                                                                  \
    __host__ __device__ void kernel()

// Opt-in (-DHIPACC_B200_DEVICE_GLOBAL_OPS): the members `void binning(x, y, pixel)` and `bin_t reduce(left, right) const`
// become __host__ __device__ as well, so that binned_data() / reduced_data() can run ANY binning / fold on the device
// (dsl_binning_kernel / dsl_reduce_kernel).  `binning(...)` with three parameters is the declaration (it also carries the
// typed launch hook); `reduce(a, b)` with exactly two arguments is the declaration, every other arity -- the reduce(dom,
// mode, lambda) calls inside kernel() bodies -- passes through untouched.
#ifdef HIPACC_B200_DEVICE_GLOBAL_OPS
#define HB_CAT_(a, b) a##b
#define HB_CAT(a, b) HB_CAT_(a, b)
#define HB_ARITY_(a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15, a16, N, ...) N
#define HB_ARITY(...) HB_ARITY_(__VA_ARGS__, M, M, M, M, M, M, M, M, M, M, M, M, M, M, 2, 1, 0)
#define reduce(...) HB_CAT(HB_REDUCE_, HB_ARITY(__VA_ARGS__))(__VA_ARGS__)
#define HB_REDUCE_2(a, b) __host__ __device__ reduce(a, b)
#define HB_REDUCE_1(a) reduce(a)
#define HB_REDUCE_M(...) reduce(__VA_ARGS__)
#define binning(...)                                                                                                        \
    hb_dispatch_global_(const ::hipacc::b200::dev::LaunchCtx &hb_ctx_, int hb_what_, unsigned int hb_bins_, void *hb_res_) override { \
        ::hipacc::b200::dev::launch_global(*this, hb_ctx_, hb_what_, hb_bins_, hb_res_);                                     \
    }                                                                                                                       \
    __host__ __device__ void binning(__VA_ARGS__)
#endif
#endif

#endif  // HIPACC_B200_DSL_HPP
