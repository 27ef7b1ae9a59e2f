// hb_point.cu -- point operators (pure streaming maps over 1-3 inputs) for sm_100a.
//
// Replaces generated point-operator kernels (bodies that only use in() / output(),
// lib/Analysis/KernelStatistics.cpp:366-388) -- e.g. SobelCombine (Sobel/src/main.cpp:76-98),
// Square1/Square2/HarrisCorner (Harris_Corner/src/main.cpp:78-164), Subsample /
// DifferenceOfGaussian / Restore / Blend (Gaussian_Laplacian_Pyramid/src/main.cpp:52-135).
// Inputs may be interpolating accessors (NN, LF; dsl/image.hpp:390-422) with the DSL's default
// CLAMP boundary (dsl/image.hpp:616-620).
//
// HBM-bound: 4 pixels per thread with vector loads / stores when nothing interpolates and the rows
// are aligned; a scalar path otherwise.  Integer promotions and the truncating store are part of
// the result and are kept exactly (C semantics of the sample bodies).
#include "hb_common.cuh"
#include "hb_internal.h"

#include <cstring>

namespace hb {

struct PointIn {
    const void *p;
    int stride, iw, ih;
    int w, h, ox, oy;  // accessor region
    int interp;
};
struct PointParams {
    PointIn in[3];
    int n_in;
    void *out;
    int out_stride, out_ox, out_oy, is_w, is_h;
    int op;
    float pf[2];
    int pi[2];
};

// ---- B5 / CF / L3 (dsl/image.hpp:321-528): TAPS x TAPS neighbourhood, weights from the fractional position --------------
__device__ __forceinline__ float w_binomial5(float d) {
    d = fabsf(d);
    return d < 0.5f ? 0.75f : d < 1.0f ? 0.5f : d < 1.5f ? 0.125f : 0.0f;
}
__device__ __forceinline__ float w_bicubic(float d) {   // Keys, a = -0.5; the library is built with -fmad=false: every * and + rounds
    d = fabsf(d);
    const float a = -0.5f;
    if (d < 1.0f) return (a + 2.0f) * d * d * d - (a + 3.0f) * d * d + 1.0f;
    if (d < 2.0f) return a * d * d * d - 5.0f * a * d * d + 8.0f * a * d - 4.0f * a;
    return 0.0f;
}
__device__ __forceinline__ float w_lanczos3(float d) {   // the reference evaluates this in double and rounds once
    d = fabsf(d);
    const double pi = 3.14159265358979323846;   // == atan(1.0) * 4 in double
    if (d == 0.0f) return 1.0f;
    if (d < 3.0f) return (float)(3.0 * (sin(pi * (double)d / 3.0) * sin(pi * (double)d)) / (pi * pi * (double)d * (double)d));
    return 0.0f;
}
template <typename T, int MODE>
__device__ __forceinline__ T fetch_interp_wide(const ImgRef<T> &im, const Window &w, int x_int, int y_int, float fx, float fy) {
    constexpr int TAPS = MODE == HB_INTERP_L3 ? 6 : 4;
    constexpr int X0 = MODE == HB_INTERP_B5 ? 0 : MODE == HB_INTERP_CF ? -1 : -2;
    constexpr int Y0 = MODE == HB_INTERP_B5 ? 0 : -1;   // L3's rows start at y_int - 1 (dsl/image.hpp:478), kept
    if (MODE == HB_INTERP_B5) { fx = (float)((double)fx + 0.5); fy = (float)((double)fy + 0.5); }
    float wx[TAPS];
#pragma unroll
    for (int i = 0; i < TAPS; ++i)
        wx[i] = MODE == HB_INTERP_B5 ? w_binomial5(fx - i) : MODE == HB_INTERP_CF ? w_bicubic(fx - 1 + i) : w_lanczos3(fx - 2 + i);
    float r = 0.0f;
#pragma unroll
    for (int j = 0; j < TAPS; ++j) {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < TAPS; ++i) {
            // the reference's second Lanczos row repeats the weight of tap 5 for tap 4 (dsl/image.hpp:489): kept, it decides results
            const float wgt = (MODE == HB_INTERP_L3 && j == 1 && i == 4) ? wx[5] : wx[i];
            const float v = (float)fetch_bh(im, w, x_int + X0 + i, y_int + Y0 + j, T(0)) * wgt;
            acc = i == 0 ? v : acc + v;
        }
        const float wy = MODE == HB_INTERP_B5 ? w_binomial5(fy - j) : MODE == HB_INTERP_CF ? w_bicubic(fy - 1 + j) : w_lanczos3(fy - 2 + j);
        r = j == 0 ? acc * wy : r + acc * wy;
    }
    return cast_out<T, float>(r);
}

// accessor value for output pixel (gx,gy) through the accessor's interpolation mode (dsl/image.hpp:390-528)
template <typename T>
__device__ __forceinline__ T fetch_interp(const PointIn &a, int gx, int gy, int is_w, int is_h) {
    ImgRef<T> im{static_cast<const T *>(a.p), a.stride, a.iw, a.ih};
    const Window w{a.ox, a.ox + a.w, a.oy, a.oy + a.h, HB_BOUNDARY_CLAMP};
    if (a.interp == HB_INTERP_NO) return im.p[(size_t)(a.oy + gy) * a.stride + a.ox + gx];
    const float stride_x = __fdiv_rn((float)a.w, (float)is_w);
    const float stride_y = __fdiv_rn((float)a.h, (float)is_h);
    // offset + stride/2 + stride*(x - is_offset)   (left to right, separately rounded)
    const float x_mapped = __fadd_rn(__fadd_rn((float)a.ox, __fdiv_rn(stride_x, 2.0f)), __fmul_rn(stride_x, (float)gx));
    const float y_mapped = __fadd_rn(__fadd_rn((float)a.oy, __fdiv_rn(stride_y, 2.0f)), __fmul_rn(stride_y, (float)gy));
    if (a.interp == HB_INTERP_NN) return fetch_bh(im, w, __float2int_rz(x_mapped), __float2int_rz(y_mapped), T(0));
    float xb = __fadd_rn(x_mapped, -0.5f), yb = __fadd_rn(y_mapped, -0.5f);
    if (xb < 0.0f) xb = 0.0f;
    if (yb < 0.0f) yb = 0.0f;
    const int x_int = __float2int_rz(xb), y_int = __float2int_rz(yb);
    const float xf = __fadd_rn(xb, -(float)x_int), yf = __fadd_rn(yb, -(float)y_int);
    if (a.interp == HB_INTERP_B5) return fetch_interp_wide<T, HB_INTERP_B5>(im, w, x_int, y_int, xf, yf);
    if (a.interp == HB_INTERP_CF) return fetch_interp_wide<T, HB_INTERP_CF>(im, w, x_int, y_int, xf, yf);
    if (a.interp == HB_INTERP_L3) return fetch_interp_wide<T, HB_INTERP_L3>(im, w, x_int, y_int, xf, yf);
    const float omx = __fadd_rn(1.0f, -xf), omy = __fadd_rn(1.0f, -yf);
    const float p00 = (float)fetch_bh(im, w, x_int, y_int, T(0));
    const float p10 = (float)fetch_bh(im, w, x_int + 1, y_int, T(0));
    const float p01 = (float)fetch_bh(im, w, x_int, y_int + 1, T(0));
    const float p11 = (float)fetch_bh(im, w, x_int + 1, y_int + 1, T(0));
    float r = __fmul_rn(__fmul_rn(omx, omy), p00);
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xf, omy), p10));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(omx, yf), p01));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xf, yf), p11));
    return cast_out<T, float>(r);  // convert<T>(float) = C cast (dsl/types.hpp:115-117)
}

template <typename TI, typename TO>
__device__ __forceinline__ TO point_eval(int op, TI a, TI b, TI c, const PointParams &p) {
    constexpr bool isf = DtypeOf<TI>::v == HB_F32;
    switch (op) {
    case HB_POINT_COPY: return (TO)a;
    case HB_POINT_SQUARE: return isf ? (TO)__fmul_rn((float)a, (float)a) : (TO)((int)a * (int)a);
    case HB_POINT_MUL: return isf ? (TO)__fmul_rn((float)a, (float)b) : (TO)((int)a * (int)b);
    case HB_POINT_SUB: return isf ? (TO)__fadd_rn((float)a, -(float)b) : (TO)((int)a - (int)b);
    case HB_POINT_ADD: return isf ? (TO)__fadd_rn((float)a, (float)b) : (TO)((int)a + (int)b);
    case HB_POINT_BLEND: return isf ? (TO)__fadd_rn((float)a, __fdiv_rn((float)b, 2.0f)) : (TO)((int)a + (int)b / 2);
    case HB_POINT_SOBEL_COMBINE: {
        const int norm = p.pi[0];
        const TI in1 = (TI)((int)a / norm), in2 = (TI)((int)b / norm);
        float r = __fsqrt_rn((float)((int)in1 * (int)in1 + (int)in2 * (int)in2));
        r = r < 255.0f ? r : 255.0f;
        r = r > 0.0f ? r : 0.0f;
        return cast_out<TO, float>(r);
    }
    case HB_POINT_HARRIS: {
        // R = ((x*y) - (xy*xy)) - (k*(x+y)*(x+y)): int determinant converted to float, float trace term,
        // no FMA contraction (it can flip the threshold test)
        const int x = (int)a, y = (int)b, xy = (int)c;
        const float det = (float)(x * y - xy * xy);
        const float s = (float)(x + y);
        const float tr = __fmul_rn(__fmul_rn(p.pf[0], s), s);
        const float R = __fadd_rn(det, -tr);
        return (TO)(R > p.pf[1] ? 1 : 0);
    }
    default: return (TO)0;
    }
}

// Scalar / interpolating path: 4 pixels per thread, every pixel through the accessor (NN / LF, tails, unaligned rows).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) point_kernel(const __grid_constant__ PointParams p) {
    const int vw = (p.is_w + 3) / 4;  // 4-pixel groups per row
    const long long total = (long long)vw * p.is_h;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int gy = (int)(g / vw), gx = (int)(g - (long long)gy * vw) * 4;
        TI v[3][4];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < p.n_in)
#pragma unroll
                for (int i = 0; i < 4; ++i) v[k][i] = gx + i < p.is_w ? fetch_interp<TI>(p.in[k], gx + i, gy, p.is_w, p.is_h) : TI(0);
        TO *dst = static_cast<TO *>(p.out) + (size_t)(p.out_oy + gy) * p.out_stride + p.out_ox + gx;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (gx + i < p.is_w) dst[i] = point_eval<TI, TO>(p.op, v[0][i], p.n_in > 1 ? v[1][i] : TI(0), p.n_in > 2 ? v[2][i] : TI(0), p);
    }
}

// Streaming path (no interpolation, 4-pixel aligned rows): one CTA per row chunk of PT x PU
// 4-pixel vectors; every thread first issues its PU independent vector loads per input (PU x NIN x 16 bytes
// in flight for float), then evaluates and stores.  OP and NIN are compile-time so the body is straight-line.
constexpr int PT = 256, PU = 4;
constexpr int kStreamCtasPerSm = 0;   // one CTA per chunk (measured: 803 vs 707 Gpx/s for the persistent grid, tools/stream_grid_sweep.sh)

template <typename T> __device__ __forceinline__ void load4_stream(const T *p, T (&v)[4]) {
    typedef typename Vec4<T>::type V;
    V t = __ldcs(reinterpret_cast<const V *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <typename T> __device__ __forceinline__ void store4_stream(T *p, const T (&v)[4]) {
    typedef typename Vec4<T>::type V;
    V t;
    t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3];
    __stcs(reinterpret_cast<V *>(p), t);
}

template <typename TI, typename TO, int OP, int NIN>
__global__ void __launch_bounds__(PT) point_stream_kernel(const __grid_constant__ PointParams p) {
    const int nvec = p.is_w / 4;                       // full vectors per row
    const int tail0 = nvec * 4;
    const int cpr = nvec > 0 ? (nvec + PT * PU - 1) / (PT * PU) : 1;
    const long long total = (long long)cpr * p.is_h;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int gy = (int)(u / cpr), c = (int)(u - (long long)gy * cpr);
        const TI *row[NIN];
#pragma unroll
        for (int k = 0; k < NIN; ++k) row[k] = static_cast<const TI *>(p.in[k].p) + (size_t)(p.in[k].oy + gy) * p.in[k].stride + p.in[k].ox;
        TO *orow = static_cast<TO *>(p.out) + (size_t)(p.out_oy + gy) * p.out_stride + p.out_ox;
        const int v0 = c * (PT * PU) + threadIdx.x;
        TI v[PU][3][4];
#pragma unroll
        for (int j = 0; j < PU; ++j) {
            const int i = v0 + j * PT;
            if (i < nvec) {
#pragma unroll
                for (int k = 0; k < NIN; ++k) load4_stream(row[k] + 4 * i, v[j][k]);
            }
        }
#pragma unroll
        for (int j = 0; j < PU; ++j) {
            const int i = v0 + j * PT;
            if (i < nvec) {
                TO o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[q] = point_eval<TI, TO>(OP, v[j][0][q], NIN > 1 ? v[j][NIN > 1 ? 1 : 0][q] : TI(0), NIN > 2 ? v[j][NIN > 2 ? 2 : 0][q] : TI(0), p);
                store4_stream(orow + 4 * i, o);
            }
        }
        if (c == 0 && threadIdx.x < p.is_w - tail0) {  // the row's last (is_w % 4) pixels
            const int x = tail0 + threadIdx.x;
            orow[x] = point_eval<TI, TO>(OP, row[0][x], NIN > 1 ? row[NIN > 1 ? 1 : 0][x] : TI(0), NIN > 2 ? row[NIN > 2 ? 2 : 0][x] : TI(0), p);
        }
    }
}

template <typename TI, typename TO, int OP, int NIN>
static void launch_stream(const PointParams &p, cudaStream_t s) {
    const int nvec = p.is_w / 4;
    const int cpr = nvec > 0 ? (nvec + PT * PU - 1) / (PT * PU) : 1;
    const long long blocks = stream_grid((long long)cpr * p.is_h, kStreamCtasPerSm);
    point_stream_kernel<TI, TO, OP, NIN><<<(unsigned)blocks, PT, 0, s>>>(p);
}

template <typename TI, typename TO>
static int launch_point(const PointParams &p, bool vec, cudaStream_t s) {
    if (vec) {
        switch (p.op) {
        case HB_POINT_COPY: launch_stream<TI, TO, HB_POINT_COPY, 1>(p, s); break;
        case HB_POINT_SQUARE: launch_stream<TI, TO, HB_POINT_SQUARE, 1>(p, s); break;
        case HB_POINT_MUL: launch_stream<TI, TO, HB_POINT_MUL, 2>(p, s); break;
        case HB_POINT_SUB: launch_stream<TI, TO, HB_POINT_SUB, 2>(p, s); break;
        case HB_POINT_ADD: launch_stream<TI, TO, HB_POINT_ADD, 2>(p, s); break;
        case HB_POINT_BLEND: launch_stream<TI, TO, HB_POINT_BLEND, 2>(p, s); break;
        case HB_POINT_SOBEL_COMBINE: launch_stream<TI, TO, HB_POINT_SOBEL_COMBINE, 2>(p, s); break;
        default: launch_stream<TI, TO, HB_POINT_HARRIS, 3>(p, s); break;
        }
    } else {
        const long long total = (long long)((p.is_w + 3) / 4) * p.is_h;
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)sm_count() * 32;  // grid-stride beyond a few waves
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        point_kernel<TI, TO><<<(unsigned)blocks, 256, 0, s>>>(p);
    }
    g_launches++;
    return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" int hb_point_op(const hb_point_desc *d, void *stream) {
    HB_REQUIRE(d && d->n_in >= 1 && d->n_in <= 3, HB_ERR_INVALID, "hb_point_op: bad descriptor");
    hb_view out = norm_view(d->out);
    HB_REQUIRE(view_ok(out), HB_ERR_INVALID, "hb_point_op: malformed output view");
    const bool x4 = is_x4(out.dtype);   // element-wise on the channels (dsl/types.hpp vector operators)
    if (x4) out = as_channels(out);
    PointParams p;
    memset(&p, 0, sizeof(p));
    const int it = channel_dtype(d->in[0].dtype);
    const size_t ies = dtype_size(it), oes = dtype_size(out.dtype);
    bool vec = (out.stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(out.data) + (size_t)out.offset_x * oes) % (4 * oes) == 0);
    for (int k = 0; k < d->n_in; ++k) {
        hb_view v = norm_view(d->in[k]);
        if (x4) {
            HB_REQUIRE(is_x4(v.dtype) && d->interp[k] == HB_INTERP_NO && d->op != HB_POINT_HARRIS, HB_ERR_UNSUPPORTED,
                       "hb_point_op: the operands of a 4-channel operator must all be 4-channel images and not interpolated; no CPU fallback");
            v = as_channels(v);
        }
        HB_REQUIRE(view_ok(v) && v.dtype == it, HB_ERR_INVALID, "hb_point_op: malformed input view %d (all inputs share one pixel type)", k);
        const int ip = d->interp[k];
        HB_REQUIRE(ip >= HB_INTERP_NO && ip <= HB_INTERP_L3, HB_ERR_INVALID, "hb_point_op: unknown interpolation mode %d", ip);
        HB_REQUIRE(ip != HB_INTERP_NO || (v.width >= out.width && v.height >= out.height), HB_ERR_INVALID,
                   "hb_point_op: input %d region smaller than the iteration space", k);
        p.in[k] = PointIn{v.data, v.stride, v.img_width, v.img_height, v.width, v.height, v.offset_x, v.offset_y, ip};
        vec = vec && ip == HB_INTERP_NO && (v.stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(v.data) + (size_t)v.offset_x * ies) % (4 * ies) == 0);
    }
    p.n_in = d->n_in;
    p.out = out.data; p.out_stride = out.stride; p.out_ox = out.offset_x; p.out_oy = out.offset_y; p.is_w = out.width; p.is_h = out.height;
    p.op = d->op;
    for (int i = 0; i < 2; ++i) { p.pf[i] = (float)d->p[i]; p.pi[i] = (int)d->p[i]; }
    HB_REQUIRE(d->op >= HB_POINT_COPY && d->op <= HB_POINT_HARRIS, HB_ERR_INVALID, "hb_point_op: unknown op %d", d->op);
    HB_REQUIRE(!(d->op == HB_POINT_SOBEL_COMBINE && p.pi[0] == 0), HB_ERR_INVALID, "hb_point_op: norm == 0");
    const int need = (d->op == HB_POINT_HARRIS) ? 3 : (d->op == HB_POINT_COPY || d->op == HB_POINT_SQUARE) ? 1 : 2;
    HB_REQUIRE(d->n_in == need, HB_ERR_INVALID, "hb_point_op: op %d takes %d inputs", d->op, need);

    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_point_op");
    int rc = HB_ERR_UNSUPPORTED;
    const int ot = out.dtype;
    if (it == HB_F32 && ot == HB_F32) rc = launch_point<float, float>(p, vec, s);
    else if (it == HB_S16 && ot == HB_S16) rc = launch_point<short, short>(p, vec, s);
    else if (it == HB_S16 && ot == HB_U8) rc = launch_point<short, uchar>(p, vec, s);
    else if (it == HB_S32 && ot == HB_U8) rc = launch_point<int, uchar>(p, vec, s);
    else if (it == HB_S32 && ot == HB_S32) rc = launch_point<int, int>(p, vec, s);
    else if (it == HB_U8 && ot == HB_U8) rc = launch_point<uchar, uchar>(p, vec, s);
    else if (it == HB_S8 && ot == HB_S8) rc = launch_point<signed char, signed char>(p, vec, s);
    HB_REQUIRE(rc == HB_OK, HB_ERR_UNSUPPORTED, "hb_point_op: no device kernel for (in %d, out %d); no CPU fallback", it, ot);
    return scope.finish();
}
