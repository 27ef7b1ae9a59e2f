// oracle/emit_cpu_fast.cpp -- TEST / BASELINE INFRASTRUCTURE ONLY: the TIMED CPU leg of bench.py.
//
// oracle/emit_cpu.cpp is the generic checker (one interpreted tap loop for every descriptor).  What Hipacc's
// `-emit-cpu` backend really prints for a Kernel is specialised code: the mask is a compile-time constant whose taps
// are unrolled into the body and whose zero entries vanish (lib/AST/Convolution.cpp:397-435), the body sits in plain
// `for gid_y / for gid_x` loops under `#pragma omp parallel for` (lib/Backend/CPU_x86.cpp:3583-3691), and with
// `-vectorize on` the interior of the iteration space runs a variant WITHOUT boundary handling while only the border
// pixels take the index remap (_CreateBoundaryBranching / *_NoBH, lib/Backend/CPU_x86.cpp:3550-3667).  This file
// restates that shape for the operators bench.py times on the host: constexpr-mask templates, interior / border split,
// unit-stride inner loops the compiler vectorises (g++ -O3 -march=x86-64-v3 -mno-fma -ffp-contract=off: AVX2, no
// FMA contraction, so every product and sum is rounded like the DSL's).
//
// Parity: tests/test_oracle.py::test_fast_cpu_leg_equals_generic_oracle checks every entry point here bit-for-bit
// against emit_cpu.cpp (which is pinned against the reference DSL) on ragged sizes, all boundary modes and ROIs.
// Anything this file has no specialisation for returns HB_ERR_UNSUPPORTED and the caller uses emit_cpu.cpp.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <omp.h>

#include "../include/hipacc_b200.h"

namespace {

typedef unsigned char uchar;

inline void norm_view(hb_view &v) {
    if (v.width <= 0 || v.height <= 0) { v.width = v.img_width; v.height = v.img_height; v.offset_x = 0; v.offset_y = 0; }
}

// lib/AST/BorderHandling.cpp:41-120: upper test first, then lower (order of :339-366)
inline int remap(int idx, int lo, int hi, int mode) {
    switch (mode) {
    case HB_BOUNDARY_CLAMP:
        if (idx >= hi) idx = hi - 1;
        if (idx < lo) idx = lo;
        break;
    case HB_BOUNDARY_REPEAT:
        while (idx >= hi) idx -= (hi - lo);
        while (idx < lo) idx += (hi - lo);
        break;
    case HB_BOUNDARY_MIRROR:
        if (idx >= hi) idx = hi - (idx + 1 - hi);
        if (idx < lo) idx = lo + (lo - idx - 1);
        break;
    default: break;
    }
    return idx;
}

template <typename T> struct Src {
    const T *p; int stride, iw, ih;
    int lo_x, hi_x, lo_y, hi_y, mode;
    T cval;
    inline T at(int x, int y) const {   // the boundary-handling variant of an access (border pixels only)
        if (mode == HB_BOUNDARY_CONSTANT) {
            if (x < lo_x || x >= hi_x || y < lo_y || y >= hi_y) return cval;
            return p[(size_t)y * stride + x];
        }
        x = remap(x, lo_x, hi_x, mode);
        y = remap(y, lo_y, hi_y, mode);
        x = std::min(std::max(x, 0), iw - 1);
        y = std::min(std::max(y, 0), ih - 1);
        return p[(size_t)y * stride + x];
    }
};

// ---- compile-time masks: the tables of the reference's samples (Sobel/src/main.cpp:128-140, Laplace/src/main.cpp:106-124)
struct Sobel3X { static constexpr int SX = 3, SY = 3; static constexpr float c[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1}; };
struct Sobel3Y { static constexpr int SX = 3, SY = 3; static constexpr float c[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1}; };
struct Laplace3D { static constexpr int SX = 3, SY = 3; static constexpr float c[9] = {2, 0, 2, 0, -8, 0, 2, 0, 2}; };
struct Laplace3N { static constexpr int SX = 3, SY = 3; static constexpr float c[9] = {0, 1, 0, 1, -4, 1, 0, 1, 0}; };
constexpr float Sobel3X::c[9];
constexpr float Sobel3Y::c[9];
constexpr float Laplace3D::c[9];
constexpr float Laplace3N::c[9];

// one output pixel, taps unrolled in row-major order, zero taps skipped when `holes` (Domain semantics,
// dsl/mask.hpp:112-126), the first visited tap initialises (dsl/kernel.hpp:250,279)
template <class M, bool HOLES, class F>
inline float fold_const(F &&px) {
    float acc = 0.0f;
    bool first = true;
    for (int k = 0; k < M::SX * M::SY; ++k) {   // fully unrolled: M::c is constexpr
        if (HOLES && M::c[k] == 0.0f) continue;
        const float v = M::c[k] * px(k % M::SX - M::SX / 2, k / M::SX - M::SY / 2);
        acc = first ? v : acc + v;
        first = false;
    }
    return acc;
}

template <class M, bool HOLES>
int local_f32_const(const hb_local_desc &d) {
    hb_view in = d.in, out = d.out;
    norm_view(in); norm_view(out);
    constexpr int hx = M::SX / 2, hy = M::SY / 2;
    Src<float> s{static_cast<const float *>(in.data), in.stride, in.img_width, in.img_height, in.offset_x, in.offset_x + in.width,
                 in.offset_y - in.ghost_top, in.offset_y + in.height + in.ghost_bottom, d.boundary, (float)d.boundary_const};
    float *op = static_cast<float *>(out.data);
    const int W = out.width, H = out.height;
    // interior columns of the iteration space: every tap stays inside the accessor's window
    const int x_lo = std::min(W, std::max(0, s.lo_x + hx - in.offset_x)), x_hi = std::max(x_lo, std::min(W, s.hi_x - hx - in.offset_x));
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < H; ++gy) {
        const int iy = in.offset_y + gy;
        float *__restrict__ orow = op + (size_t)(out.offset_y + gy) * out.stride + out.offset_x;
        const bool row_in = iy - hy >= s.lo_y && iy + hy < s.hi_y;
        auto border = [&](int gx) {
            const int ix = in.offset_x + gx;
            orow[gx] = fold_const<M, HOLES>([&](int dx, int dy) { return s.at(ix + dx, iy + dy); });
        };
        if (!row_in) {
            for (int gx = 0; gx < W; ++gx) border(gx);
            continue;
        }
        for (int gx = 0; gx < x_lo; ++gx) border(gx);
        const float *__restrict__ c0 = s.p + (size_t)iy * s.stride + in.offset_x;   // centre row; rows dy away are dy * stride apart
        const ptrdiff_t st = s.stride;
#pragma omp simd
        for (int gx = x_lo; gx < x_hi; ++gx)   // the *_NoBH variant: raw loads, no index tests
            orow[gx] = fold_const<M, HOLES>([&](int dx, int dy) { return c0[gx + dx + dy * st]; });
        for (int gx = x_hi; gx < W; ++gx) border(gx);
    }
    return HB_OK;
}

// run-time coefficients, compile-time size, every tap visited (convolve(), or a Domain without holes)
template <typename TI, typename TO, int SX, int SY>
int local_sum_full(const hb_local_desc &d, const float *coef) {
    hb_view in = d.in, out = d.out;
    norm_view(in); norm_view(out);
    constexpr int hx = SX / 2, hy = SY / 2;
    Src<TI> s{static_cast<const TI *>(in.data), in.stride, in.img_width, in.img_height, in.offset_x, in.offset_x + in.width,
              in.offset_y - in.ghost_top, in.offset_y + in.height + in.ghost_bottom, d.boundary, (TI)d.boundary_const};
    TO *op = static_cast<TO *>(out.data);
    const int W = out.width, H = out.height;
    float c[SX * SY];
    for (int k = 0; k < SX * SY; ++k) c[k] = coef[k];
    const bool add = d.epilogue == HB_EPI_ADD_CAST;
    const float addend = (float)d.epi_p[0];
    auto finish = [&](float acc) -> TO {
        if (add) acc = acc + addend;
        if (sizeof(TO) == 4) return (TO)acc;   // float out
        return (TO)(int)acc;                   // (uchar)f: through int like g++/x86 (cvttss2si), then wrap
    };
    const int x_lo = std::min(W, std::max(0, s.lo_x + hx - in.offset_x)), x_hi = std::max(x_lo, std::min(W, s.hi_x - hx - in.offset_x));
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < H; ++gy) {
        const int iy = in.offset_y + gy;
        TO *__restrict__ orow = op + (size_t)(out.offset_y + gy) * out.stride + out.offset_x;
        const bool row_in = iy - hy >= s.lo_y && iy + hy < s.hi_y;
        auto border = [&](int gx) {
            const int ix = in.offset_x + gx;
            float acc = 0.0f;
            for (int k = 0; k < SX * SY; ++k) {
                const float v = c[k] * (float)s.at(ix + k % SX - hx, iy + k / SX - hy);
                acc = k == 0 ? v : acc + v;
            }
            orow[gx] = finish(acc);
        };
        if (!row_in) {
            for (int gx = 0; gx < W; ++gx) border(gx);
            continue;
        }
        for (int gx = 0; gx < x_lo; ++gx) border(gx);
        const TI *__restrict__ c0 = s.p + (size_t)iy * s.stride + in.offset_x;
        const ptrdiff_t st = s.stride;
        float cc[SX * SY];   // a copy no lambda captures: it cannot alias the (possibly char-typed) output stores
        for (int k = 0; k < SX * SY; ++k) cc[k] = coef[k];
        const float add_v = addend;
        const bool add_f = add;
#pragma omp simd
        for (int gx = x_lo; gx < x_hi; ++gx) {
            float acc = cc[0] * (float)c0[gx - hx - hy * st];
#pragma GCC unroll 64
            for (int k = 1; k < SX * SY; ++k) acc = acc + cc[k] * (float)c0[gx + (k % SX - hx) + (k / SX - hy) * st];
            if (add_f) acc = acc + add_v;
            orow[gx] = sizeof(TO) == 4 ? (TO)acc : (TO)(int)acc;
        }
        for (int gx = x_hi; gx < W; ++gx) border(gx);
    }
    return HB_OK;
}

template <class M> bool mask_is(const hb_local_desc &d, bool &holes_ok) {
    if (d.size_x != M::SX || d.size_y != M::SY || !d.coef_f32) return false;
    for (int k = 0; k < M::SX * M::SY; ++k)
        if (d.coef_f32[k] != M::c[k]) return false;
    // REDUCE_DOMAIN with the footprint derived from the mask (or an explicit one equal to it) skips the zero taps;
    // CONVOLVE visits them (0 * pixel joins the sum)
    holes_ok = true;
    if (d.kind == HB_LOCAL_REDUCE_DOMAIN && d.domain)
        for (int k = 0; k < M::SX * M::SY; ++k)
            if ((d.domain[k] != 0) != (M::c[k] != 0.0f)) return false;
    return true;
}

template <class M> int try_const(const hb_local_desc &d, bool &taken) {
    bool ok = false;
    if (!mask_is<M>(d, ok)) return HB_OK;
    taken = true;
    return d.kind == HB_LOCAL_REDUCE_DOMAIN ? local_f32_const<M, true>(d) : local_f32_const<M, false>(d);
}

}  // namespace

extern "C" {

// The specialised (-emit-cpu shaped) form of oc_local_op for the operators the bench times on the host:
//   float -> float SUM of coef * in, constexpr Sobel / Laplace 3x3 masks or any full 3x3 / 5x5 / 7x7 mask, plain cast;
//   uchar -> uchar float-mask Gaussians 3x3 / 5x5 / 7x7 with the +0.5f epilogue (Gaussian_Blur/src/main.cpp:62-66).
// HB_ERR_UNSUPPORTED = no specialisation (the caller falls back to the generic oc_local_op).
int ocf_local_op(const hb_local_desc *d) {
    if (!d || d->size_x <= 0 || d->size_y <= 0 || d->size_x != d->size_y) return HB_ERR_UNSUPPORTED;
    if (d->reduce_mode != HB_REDUCE_SUM || d->tap != HB_TAP_MUL || !d->coef_f32 || d->acc_dtype != HB_F32) return HB_ERR_UNSUPPORTED;
    const int n = d->size_x * d->size_y;
    bool full = true;   // every tap visited?
    if (d->kind == HB_LOCAL_REDUCE_DOMAIN)
        for (int k = 0; k < n; ++k) full = full && (d->domain ? d->domain[k] != 0 : d->coef_f32[k] != 0.0f);
    if (d->in.dtype == HB_F32 && d->out.dtype == HB_F32 && d->epilogue == HB_EPI_CAST) {
        if (d->size_x == 3) {
            bool taken = false;
            int rc = try_const<Sobel3X>(*d, taken);
            if (!taken) rc = try_const<Sobel3Y>(*d, taken);
            if (!taken) rc = try_const<Laplace3D>(*d, taken);
            if (!taken) rc = try_const<Laplace3N>(*d, taken);
            if (taken) return rc;
        }
        if (!full) return HB_ERR_UNSUPPORTED;
        if (d->size_x == 3) return local_sum_full<float, float, 3, 3>(*d, d->coef_f32);
        if (d->size_x == 5) return local_sum_full<float, float, 5, 5>(*d, d->coef_f32);
        if (d->size_x == 7) return local_sum_full<float, float, 7, 7>(*d, d->coef_f32);
        return HB_ERR_UNSUPPORTED;
    }
    if (d->in.dtype == HB_U8 && d->out.dtype == HB_U8 && full && (d->epilogue == HB_EPI_ADD_CAST || d->epilogue == HB_EPI_CAST)) {
        if (d->size_x == 3) return local_sum_full<uchar, uchar, 3, 3>(*d, d->coef_f32);
        if (d->size_x == 5) return local_sum_full<uchar, uchar, 5, 5>(*d, d->coef_f32);
        if (d->size_x == 7) return local_sum_full<uchar, uchar, 7, 7>(*d, d->coef_f32);
    }
    return HB_ERR_UNSUPPORTED;
}

// The Harris corner pipeline as -emit-cpu prints it: the sample's NINE kernels (Harris_Corner/src/main.cpp:55-164, 230-305)
// one after the other over eight full-size intermediate images, each a plain row loop under OpenMP with its constant 3x3
// mask unrolled, the interior columns without boundary handling (vectorised by the compiler), CLAMP at the image edge.
//   Sobel (uchar -> short): short sum of Input * {-1,0,1} over the six non-zero taps, / 6
//   Square1 x 2, Square2 (short -> short), Gaussian x 3 (short -> short): int sum of Input * {1,2,1;2,4,2;1,2,1}, / 16
//   HarrisCorner (3 x short -> uchar): float R = (x*y - xy*xy) - (k*(x+y))*(x+y); R > threshold
// in / out: dense or pitched uchar images of w x h pixels.  Returns HB_OK.  Bit-identical to the generic oracle pipeline
// (tests/test_oracle.py::test_fast_cpu_harris_equals_generic_oracle).
int ocf_harris(const unsigned char *in, unsigned char *out, int w, int h, int in_stride, int out_stride, float k, float threshold) {
    if (!in || !out || w <= 0 || h <= 0) return HB_ERR_INVALID;
    const size_t n = (size_t)w * h;
    // the eight intermediate images live across calls like the sample's Image objects (allocated once, outside any timed
    // region; not thread-safe: one caller at a time, which is how the tests and the bench use it)
    static short *buf = nullptr;
    static size_t cap = 0;
    if (cap < 8 * n) {
        free(buf);
        buf = static_cast<short *>(malloc(8 * n * sizeof(short)));
        cap = buf ? 8 * n : 0;
    }
    if (!buf) return HB_ERR_INVALID;
    short *dx = buf, *dy = buf + n, *sx = buf + 2 * n, *sy = buf + 3 * n, *sxy = buf + 4 * n, *gx = buf + 5 * n, *gy = buf + 6 * n, *gxy = buf + 7 * n;
    auto cl = [](int v, int hi) { return v < 0 ? 0 : (v >= hi ? hi - 1 : v); };
    // ---- Sobel dx / dy: two kernels
    for (int which = 0; which < 2; ++which) {
        short *dst = which == 0 ? dx : dy;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < h; ++y) {
            const uchar *r0 = in + (size_t)cl(y - 1, h) * in_stride, *r1 = in + (size_t)y * in_stride, *r2 = in + (size_t)cl(y + 1, h) * in_stride;
            short *o = dst + (size_t)y * w;
            auto px = [&](int x, int xm, int xp) -> short {
                short sum;
                if (which == 0) sum = (short)((short)((short)((short)((short)(-r0[xm] + r0[xp]) - r1[xm]) + r1[xp]) - r2[xm]) + r2[xp]);
                else sum = (short)((short)((short)((short)((short)(-r0[xm] - r0[x]) - r0[xp]) + r2[xm]) + r2[x]) + r2[xp]);
                return (short)(sum / 6);
            };
            o[0] = px(0, 0, cl(1, w));
            if (which == 0) {
                for (int x = 1; x < w - 1; ++x) o[x] = (short)((short)(-r0[x - 1] + r0[x + 1] - r1[x - 1] + r1[x + 1] - r2[x - 1] + r2[x + 1]) / 6);
            } else {
                for (int x = 1; x < w - 1; ++x) o[x] = (short)((short)(-r0[x - 1] - r0[x] - r0[x + 1] + r2[x - 1] + r2[x] + r2[x + 1]) / 6);
            }
            if (w > 1) o[w - 1] = px(w - 1, w - 2, w - 1);
        }
    }
    // ---- Square1 (dx), Square1 (dy), Square2 (dx, dy): three kernels
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) { const short *a = dx + (size_t)y * w; short *o = sx + (size_t)y * w; for (int x = 0; x < w; ++x) o[x] = (short)(a[x] * a[x]); }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) { const short *a = dy + (size_t)y * w; short *o = sy + (size_t)y * w; for (int x = 0; x < w; ++x) o[x] = (short)(a[x] * a[x]); }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) { const short *a = dx + (size_t)y * w, *b = dy + (size_t)y * w; short *o = sxy + (size_t)y * w; for (int x = 0; x < w; ++x) o[x] = (short)(a[x] * b[x]); }
    // ---- Gaussian 3x3 / 16 on the three product images: three kernels
    const short *srcs[3] = {sx, sy, sxy};
    short *dsts[3] = {gx, gy, gxy};
    for (int pl = 0; pl < 3; ++pl) {
        const short *src = srcs[pl];
        short *dst = dsts[pl];
#pragma omp parallel for schedule(static)
        for (int y = 0; y < h; ++y) {
            const short *r0 = src + (size_t)cl(y - 1, h) * w, *r1 = src + (size_t)y * w, *r2 = src + (size_t)cl(y + 1, h) * w;
            short *o = dst + (size_t)y * w;
            auto px = [&](int x, int xm, int xp) -> short {
                const int sum = r0[xm] + 2 * r0[x] + r0[xp] + 2 * r1[xm] + 4 * r1[x] + 2 * r1[xp] + r2[xm] + 2 * r2[x] + r2[xp];
                return (short)(sum / 16);
            };
            o[0] = px(0, 0, cl(1, w));
            for (int x = 1; x < w - 1; ++x) {
                const int sum = r0[x - 1] + 2 * r0[x] + r0[x + 1] + 2 * r1[x - 1] + 4 * r1[x] + 2 * r1[x + 1] + r2[x - 1] + 2 * r2[x] + r2[x + 1];
                o[x] = (short)(sum / 16);
            }
            if (w > 1) o[w - 1] = px(w - 1, w - 2, w - 1);
        }
    }
    // ---- HarrisCorner
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const short *a = gx + (size_t)y * w, *b = gy + (size_t)y * w, *c = gxy + (size_t)y * w;
        uchar *o = out + (size_t)y * out_stride;
        for (int x = 0; x < w; ++x) {
            const int X = a[x], Y = b[x], XY = c[x];
            const float R = (float)((X * Y) - (XY * XY)) - (k * (float)(X + Y)) * (float)(X + Y);
            o[x] = R > threshold ? 1 : 0;
        }
    }
    return HB_OK;
}

int ocf_num_threads(void) { return omp_get_max_threads(); }
void ocf_set_num_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
