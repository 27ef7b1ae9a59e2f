// hb_common.cuh -- device-side building blocks shared by all operator kernels (sm_100a).
//
// Numeric contract (DESIGN.md "Numerics"): every float multiply / add that exists in the DSL
// program is a separately rounded operation (the library is compiled with -fmad=false, FMAs
// are written explicitly where the contract allows them); float -> integer stores go through
// int with truncation like g++/x86 does for `(uchar)f`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hipacc_b200.h"

namespace hb {

typedef unsigned char uchar;

template <typename T> struct DtypeOf;
template <> struct DtypeOf<uchar> { static constexpr int v = HB_U8; };
template <> struct DtypeOf<signed char> { static constexpr int v = HB_S8; };
template <> struct DtypeOf<unsigned short> { static constexpr int v = HB_U16; };
template <> struct DtypeOf<short> { static constexpr int v = HB_S16; };
template <> struct DtypeOf<int> { static constexpr int v = HB_S32; };
template <> struct DtypeOf<unsigned int> { static constexpr int v = HB_U32; };
template <> struct DtypeOf<float> { static constexpr int v = HB_F32; };

// Boundary window of an accessor in image coordinates (vertical ghost rows folded in) plus the
// mode; mirrors lower/upper of lib/AST/BorderHandling.cpp and dsl/image.hpp:574-580.
struct Window {
    int lo_x, hi_x, lo_y, hi_y;
    int mode;
};

// lib/AST/BorderHandling.cpp:41-120 (upper test first, then lower: order of :339-366)
__device__ __forceinline__ int remap_idx(int idx, int lo, int hi, int mode) {
    if (mode == HB_BOUNDARY_CLAMP) {
        if (idx >= hi) idx = hi - 1;
        if (idx < lo) idx = lo;
    } else if (mode == HB_BOUNDARY_MIRROR) {
        if (idx >= hi) idx = hi - (idx + 1 - hi);
        if (idx < lo) idx = lo + (lo - idx - 1);
    } else if (mode == HB_BOUNDARY_REPEAT) {
        const int n = hi - lo;
        while (idx >= hi) idx -= n;
        while (idx < lo) idx += n;
    }
    return idx;
}

// A 2-D image in HBM as the kernels see it.
template <typename T>
struct ImgRef {
    const T *__restrict__ p;
    int stride;  // pixels
    int iw, ih;  // allocation extent (memory-safety clamp for UNDEFINED / degenerate windows)
};

// neighbour fetch through the boundary mode (dsl/image.hpp:574-612); `cval` for CONSTANT
template <typename T>
__device__ __forceinline__ T fetch_bh(const ImgRef<T> &im, const Window &w, int x, int y, T cval) {
    if (w.mode == HB_BOUNDARY_CONSTANT) {
        if (x < w.lo_x || x >= w.hi_x || y < w.lo_y || y >= w.hi_y) return cval;
    } else {
        x = remap_idx(x, w.lo_x, w.hi_x, w.mode);
        y = remap_idx(y, w.lo_y, w.hi_y, w.mode);
        x = min(max(x, 0), im.iw - 1);
        y = min(max(y, 0), im.ih - 1);
    }
    return im.p[(size_t)y * im.stride + x];
}

// C conversions on store.  float -> integer: truncate through int (cvt.rzi.s32.f32), then wrap.
template <typename TO, typename TA> struct CastOut {
    __device__ __forceinline__ static TO f(TA v) { return (TO)v; }
};
template <typename TO> struct CastOut<TO, float> {
    __device__ __forceinline__ static TO f(float v) { return (TO)__float2int_rz(v); }
};
template <> struct CastOut<float, float> {
    __device__ __forceinline__ static float f(float v) { return v; }
};
template <typename TO, typename TA>
__device__ __forceinline__ TO cast_out(TA v) { return CastOut<TO, TA>::f(v); }

// 4-pixel vectors for loads / stores of each pixel type
template <typename T> struct Vec4;
template <> struct Vec4<uchar> { typedef uchar4 type; };
template <> struct Vec4<signed char> { typedef char4 type; };
template <> struct Vec4<short> { typedef short4 type; };
template <> struct Vec4<unsigned short> { typedef ushort4 type; };
template <> struct Vec4<int> { typedef int4 type; };
template <> struct Vec4<float> { typedef float4 type; };

template <typename T>
__device__ __forceinline__ void load4(const T *p, T (&v)[4]) {
    typedef typename Vec4<T>::type V;
    V t = *reinterpret_cast<const V *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <typename T>
__device__ __forceinline__ void store4(T *p, const T (&v)[4]) {
    typedef typename Vec4<T>::type V;
    V t;
    t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3];
    *reinterpret_cast<V *>(p) = t;
}

__host__ __device__ __forceinline__ constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------------
// Tile staging: ROWS x TWS elements of the compute type TS, smem column 0 <-> input x `x_start`,
// row 0 <-> input y `y_start`.  Interior tiles (footprint inside the boundary window, rows 4-pixel
// aligned) take the branch-free vector path; border tiles fetch every element through the boundary
// mode, so the compute code that follows never sees a boundary.
// ------------------------------------------------------------------------------------------------
// CH > 1: the image holds CH interleaved channels per pixel (uchar4); columns count channel ELEMENTS, the boundary
// window `w` counts PIXELS, so the remap works on the pixel index and keeps the channel.
template <typename T, int CH>
__device__ __forceinline__ T fetch_bh_elem(const ImgRef<T> &im, const Window &w, int ex, int y, T cval) {
    if (CH == 1) return fetch_bh(im, w, ex, y, cval);
    int px = ex >= 0 ? ex / CH : -((CH - 1 - ex) / CH);   // floor(ex / CH)
    const int ch = ex - px * CH;
    if (w.mode == HB_BOUNDARY_CONSTANT) {
        if (px < w.lo_x || px >= w.hi_x || y < w.lo_y || y >= w.hi_y) return cval;
    } else {
        px = remap_idx(px, w.lo_x, w.hi_x, w.mode);
        y = remap_idx(y, w.lo_y, w.hi_y, w.mode);
        px = min(max(px, 0), im.iw / CH - 1);
        y = min(max(y, 0), im.ih - 1);
    }
    return im.p[(size_t)y * im.stride + px * CH + ch];
}

template <typename TI, typename TS, int ROWS, int TWS, int NTHREADS, int CH = 1>
__device__ __forceinline__ void stage_tile(TS *tile, const TI *__restrict__ in, int in_stride, int in_iw, int in_ih,
                                           const Window w, TI cval, int x_start, int y_start, int tid) {
    const bool interior = x_start >= w.lo_x * CH && x_start + TWS <= w.hi_x * CH && y_start >= w.lo_y && y_start + ROWS <= w.hi_y;
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) + (size_t)x_start * sizeof(TI)) % (4 * sizeof(TI)) == 0) && (in_stride % 4 == 0);
    if (interior && aligned) {
        constexpr int VPR = TWS / 4;
        const TI *base = in + (size_t)y_start * in_stride + x_start;
#pragma unroll 2
        for (int v = tid; v < ROWS * VPR; v += NTHREADS) {
            const int r = v / VPR, c4 = v - r * VPR;
            TI t[4];
            load4(base + (size_t)r * in_stride + 4 * c4, t);
            TS s[4] = {(TS)t[0], (TS)t[1], (TS)t[2], (TS)t[3]};
            store4(tile + r * TWS + 4 * c4, s);
        }
    } else {
        ImgRef<TI> im{in, in_stride, in_iw, in_ih};
        for (int e = tid; e < ROWS * TWS; e += NTHREADS) {
            const int r = e / TWS, c = e - r * TWS;
            tile[e] = (TS)fetch_bh_elem<TI, CH>(im, w, x_start + c, y_start + r, cval);
        }
    }
}

}  // namespace hb
