// hb_pyramid.cu -- Gaussian / Laplacian pyramid level transitions (float) for sm_100a.
//
// The reference sample (samples-public/5_Other/Gaussian_Laplacian_Pyramid/src/main.cpp:199-248)
// issues per level: Gaussian (fine -> tmp), Subsample (NN, tmp -> coarse), DifferenceOfGaussian
// (fine - LF(coarse) -> lap) on the way down and Restore (LF(coarse) + lap -> fine), Blend
// (LF(coarse lap) + lap/2 -> lap) on the way up: 37 bytes of HBM traffic per fine pixel.
//
//   hb_pyr_down : when `tmp` is not requested the blur is evaluated ONLY at the NN-sampled positions
//                 (one quarter of the pixels, no tmp round trip); taps are folded in the same
//                 row-major order with separately rounded mul/add, so `coarse` is bit-identical to
//                 blur-then-subsample.  The DoG is then one streaming pass.
//   hb_pyr_up   : Restore and Blend in ONE pass (lap(l) is read once, both outputs written).
// Interpolation follows dsl/image.hpp:390-422 (cell-centred mapping, LF on x_mapped - 0.5 clamped at
// 0, neighbours through the DSL's default CLAMP, dsl/image.hpp:616-620).
#include "hb_common.cuh"
#include "hb_internal.h"
#include "hb_tma.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

namespace hb {

struct PlaneRef {
    float *p;
    int stride, iw, ih;
    int w, h, ox, oy;
};
static PlaneRef plane_of(const hb_view &v) {
    return PlaneRef{static_cast<float *>(v.data), v.stride, v.img_width, v.img_height, v.width, v.height, v.offset_x, v.offset_y};
}

struct PyrDownParams {
    PlaneRef fine, coarse;
    int size;
    float coef[49];
};

// coarse(x,y) = sum_{taps row-major} coef * fine(clamp(sx + dx), clamp(sy + dy)),  (sx,sy) = NN-sampled position
template <int S>
__global__ void __launch_bounds__(256) pyr_blur_subsample_kernel(const __grid_constant__ PyrDownParams p) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= p.coarse.w || y >= p.coarse.h) return;
    constexpr int H = S / 2;
    const float stride_x = __fdiv_rn((float)p.fine.w, (float)p.coarse.w);
    const float stride_y = __fdiv_rn((float)p.fine.h, (float)p.coarse.h);
    const int sx = __float2int_rz(__fadd_rn(__fadd_rn((float)p.fine.ox, __fdiv_rn(stride_x, 2.0f)), __fmul_rn(stride_x, (float)x)));
    const int sy = __float2int_rz(__fadd_rn(__fadd_rn((float)p.fine.oy, __fdiv_rn(stride_y, 2.0f)), __fmul_rn(stride_y, (float)y)));
    // NN fetch itself goes through CLAMP on the tmp image (same extent as fine)
    const int lo_x = p.fine.ox, hi_x = p.fine.ox + p.fine.w, lo_y = p.fine.oy, hi_y = p.fine.oy + p.fine.h;
    const int cx = min(max(sx, lo_x), hi_x - 1), cy = min(max(sy, lo_y), hi_y - 1);
    float acc = 0.0f;
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
        const int yy = min(max(cy + dy - H, lo_y), hi_y - 1);
        const float *row = p.fine.p + (size_t)yy * p.fine.stride;
#pragma unroll
        for (int dx = 0; dx < S; ++dx) {
            const int xx = min(max(cx + dx - H, lo_x), hi_x - 1);
            acc = __fadd_rn(acc, __fmul_rn(__ldg(row + xx), p.coef[dy * S + dx]));  // in(mask) * mask()
        }
    }
    p.coarse.p[(size_t)(p.coarse.oy + y) * p.coarse.stride + p.coarse.ox + x] = acc;
}

// bilinear sample of `c` for output pixel (gx,gy) of an iteration space (is_w x is_h)
__device__ __forceinline__ float lf_sample(const PlaneRef &c, int gx, int gy, float stride_x, float stride_y) {
    const float x_mapped = __fadd_rn(__fadd_rn((float)c.ox, __fdiv_rn(stride_x, 2.0f)), __fmul_rn(stride_x, (float)gx));
    const float y_mapped = __fadd_rn(__fadd_rn((float)c.oy, __fdiv_rn(stride_y, 2.0f)), __fmul_rn(stride_y, (float)gy));
    float xb = __fadd_rn(x_mapped, -0.5f), yb = __fadd_rn(y_mapped, -0.5f);
    if (xb < 0.0f) xb = 0.0f;
    if (yb < 0.0f) yb = 0.0f;
    const int x_int = __float2int_rz(xb), y_int = __float2int_rz(yb);
    const float xf = __fadd_rn(xb, -(float)x_int), yf = __fadd_rn(yb, -(float)y_int);
    const float omx = __fadd_rn(1.0f, -xf), omy = __fadd_rn(1.0f, -yf);
    const int lo_x = c.ox, hi_x = c.ox + c.w - 1, lo_y = c.oy, hi_y = c.oy + c.h - 1;
    const int x0 = min(max(x_int, lo_x), hi_x), x1 = min(max(x_int + 1, lo_x), hi_x);
    const int y0 = min(max(y_int, lo_y), hi_y), y1 = min(max(y_int + 1, lo_y), hi_y);
    const float *r0 = c.p + (size_t)y0 * c.stride, *r1 = c.p + (size_t)y1 * c.stride;
    float r = __fmul_rn(__fmul_rn(omx, omy), __ldg(r0 + x0));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xf, omy), __ldg(r0 + x1)));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(omx, yf), __ldg(r1 + x0)));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xf, yf), __ldg(r1 + x1)));
    return r;
}

struct PyrDogParams {
    PlaneRef fine, coarse, lap;
};
// lap(x,y) = fine(x,y) - LF(coarse)(x,y)
__global__ void __launch_bounds__(256) pyr_dog_kernel(const __grid_constant__ PyrDogParams p) {
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = (blockIdx.y * 8 + threadIdx.y) * 4;
    if (x >= p.lap.w) return;
    const float stride_x = __fdiv_rn((float)p.coarse.w, (float)p.lap.w);
    const float stride_y = __fdiv_rn((float)p.coarse.h, (float)p.lap.h);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = y0 + j;
        if (y >= p.lap.h) return;
        const float f = __ldg(p.fine.p + (size_t)(p.fine.oy + y) * p.fine.stride + p.fine.ox + x);
        p.lap.p[(size_t)(p.lap.oy + y) * p.lap.stride + p.lap.ox + x] = __fadd_rn(f, -lf_sample(p.coarse, x, y, stride_x, stride_y));
    }
}

struct PyrUpParams {
    PlaneRef cg, cl, fg, fl;
};
// fine_gaus = LF(coarse_gaus) + lap ; lap = LF(coarse_lap) + lap / 2      (one pass, lap read once)
__global__ void __launch_bounds__(256) pyr_up_kernel(const __grid_constant__ PyrUpParams p) {
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = (blockIdx.y * 8 + threadIdx.y) * 4;
    if (x >= p.fg.w) return;
    const float gsx = __fdiv_rn((float)p.cg.w, (float)p.fg.w), gsy = __fdiv_rn((float)p.cg.h, (float)p.fg.h);
    const float lsx = __fdiv_rn((float)p.cl.w, (float)p.fl.w), lsy = __fdiv_rn((float)p.cl.h, (float)p.fl.h);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = y0 + j;
        if (y >= p.fg.h) return;
        float *lp = p.fl.p + (size_t)(p.fl.oy + y) * p.fl.stride + p.fl.ox + x;
        const float l = *lp;
        p.fg.p[(size_t)(p.fg.oy + y) * p.fg.stride + p.fg.ox + x] = __fadd_rn(lf_sample(p.cg, x, y, gsx, gsy), l);
        *lp = __fadd_rn(lf_sample(p.cl, x, y, lsx, lsy), __fdiv_rn(l, 2.0f));
    }
}


// ================================================================================================
// Exact-halving transitions (fine = 2 x coarse in both dimensions, the case of every level of a
// power-of-two pyramid): the accessor strides are exactly 2 (NN) and 0.5 (LF), so the float mapping
// of dsl/image.hpp:390-422 collapses to integers with exact weights:
//   NN : coarse (cx,cy) samples fine (2cx+1, 2cy+1)
//   LF : fine x = 2m   (m > 0): x_int = m-1, xf = 0.75 ;  x = 0: x_int = 0, xf = 0 (x_mapped-0.5 clamps at 0)
//        fine x = 2m+1        : x_int = m,   xf = 0.25 ;  neighbour x_int+1 through CLAMP
// (0.25 + 0.5*x, the subtraction of 0.5 and the weight products are all exact in float, so using the
// constants is bit-identical to evaluating the general formula.)  The sum keeps the DSL's order:
// ((omx*omy)*p00 + (xf*omy)*p10) + (omx*yf)*p01) + (xf*yf)*p11.
// ================================================================================================

// generic (edge-safe) bilinear sample for the exact-halving case; `at(cx,cy)` returns the coarse value.
// Row-strip sharding: `top_edge` says fine row 0 is the global top (x_mapped - 0.5 clamps at 0 only there; an
// interior strip's row 0 blends with coarse row -1, a ghost row), `cy_max` is the last coarse row that exists
// (ch - 1 at the global bottom, ch - 1 + ghost rows otherwise).
template <typename F>
__device__ __forceinline__ float lf_half(F at, int x, int y, int cw, int cy_max, bool top_edge) {
    int x0, x1, y0, y1;
    float xf, yf;
    if (x & 1) { x0 = x >> 1; x1 = min(x0 + 1, cw - 1); xf = 0.25f; }
    else if (x == 0) { x0 = 0; x1 = min(1, cw - 1); xf = 0.0f; }
    else { x0 = (x >> 1) - 1; x1 = x0 + 1; xf = 0.75f; }
    if (y & 1) { y0 = y >> 1; y1 = min(y0 + 1, cy_max); yf = 0.25f; }
    else if (y == 0 && top_edge) { y0 = 0; y1 = min(1, cy_max); yf = 0.0f; }
    else { y0 = (y >> 1) - 1; y1 = y0 + 1; yf = 0.75f; }
    const float omx = __fadd_rn(1.0f, -xf), omy = __fadd_rn(1.0f, -yf);
    float r = __fmul_rn(__fmul_rn(omx, omy), at(x0, y0));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xf, omy), at(x1, y0)));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(omx, yf), at(x0, y1)));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xf, yf), at(x1, y1)));
    return r;
}

// Interior fast path: c[j][i] = coarse(m-1+i, k-1+j) for the thread's 4x4 fine block at (2m, 2k), m,k >= 1 (or
// clamped duplicates at the right / bottom edge).  Pixel (i,j) of the block uses columns (i+1)/2, (i+1)/2+1 and
// rows (j+1)/2, (j+1)/2+1 with weights (0.25,0.75) for even and (0.75,0.25) for odd offsets.
__device__ __forceinline__ float lf_block(const float (&c)[4][4], int i, int j) {
    const int ci = (i + 1) >> 1, cj = (j + 1) >> 1;
    const float xf = (i & 1) ? 0.25f : 0.75f, yf = (j & 1) ? 0.25f : 0.75f;
    const float omx = 1.0f - xf, omy = 1.0f - yf;   // exact
    float r = __fmul_rn(omx * omy, c[cj][ci]);      // weight products are exact constants
    r = __fadd_rn(r, __fmul_rn(xf * omy, c[cj][ci + 1]));
    r = __fadd_rn(r, __fmul_rn(omx * yf, c[cj + 1][ci]));
    r = __fadd_rn(r, __fmul_rn(xf * yf, c[cj + 1][ci + 1]));
    return r;
}

constexpr int FD_TW = 128, FD_TH = 32;                 // fine tile of one CTA
constexpr int FD_CCOLS = 68, FD_CROWS = 18;            // coarse tile incl. halo 1 (66 used columns, computed as 34 pairs)
constexpr int FD_FCOLS = 144;                          // staged fine columns (column 0 <-> x = X0 - 4)
constexpr int FD_NT = 256;

struct PyrFusedParams {
    const float *fine;
    float *coarse, *lap;
    int fine_stride, coarse_stride, lap_stride;
    int fw, fh, cw, ch;
    int fgt, fgb;   // ghost rows of real neighbour data above / below the fine region (row-strip sharding)
    float coef[49];
};

// blur(fine) sampled at (2cx+1, 2cy+1) -> coarse, and lap = fine - LF(coarse), one pass: 4 B read + 1 B + 4 B
// written per fine pixel (the unfused sequence moves 19).  Phase 1 computes the coarse tile with a 1-pixel halo
// (the halo is recomputed by the neighbouring CTAs, 16 % extra FP work) from the staged fine tile: each thread
// owns 2 coarse columns x 3 coarse rows and walks the S+4 fine rows once (row-stationary, 16-byte LDS).  Phase 2
// is the DoG on 4x4 fine blocks from the two tiles.
// COHERENT: the tile function runs inside the multi-level kernel below, where the fine level was written earlier in the
// SAME launch by other CTAs -- its loads must then bypass the non-coherent read-only path (ld.global.cg instead of .nc)
template <bool COHERENT> __device__ __forceinline__ float4 ld_f4(const float4 *p) { return COHERENT ? __ldcg(p) : __ldg(p); }
template <bool COHERENT> __device__ __forceinline__ float2 ld_f2(const float2 *p) { return COHERENT ? __ldcg(p) : __ldg(p); }
template <bool COHERENT> __device__ __forceinline__ float ld_f1(const float *p) { return COHERENT ? __ldcg(p) : __ldg(p); }

template <int S> struct DownSmem {
    static constexpr int FROWS = 35 + 2 * (S / 2);     // row 0 <-> y = Y0 - 1 - H
    static constexpr int FLOATS = FROWS * FD_FCOLS + FD_CROWS * FD_CCOLS;
};

// one 128 x 32 fine tile at tile coordinates (bx, by); every thread of the CTA must call it (it synchronises)
// DOG = false: blur + subsample only (lap is not written; the DifferenceOfGaussian runs as its own kernel, pyr_dog_half_kernel)
template <int S, bool COHERENT, bool DOG = true>
__device__ __forceinline__ void pyr_down_tile(const PyrFusedParams &p, const int bx, const int by, float *ftile, float *ctile,
                                              const CUtensorMap *tmap = nullptr, uint64_t *bar = nullptr) {
    constexpr int H = S / 2;
    constexpr int FROWS = DownSmem<S>::FROWS;
    const int tid = threadIdx.x;
    const int X0 = bx * FD_TW, Y0 = by * FD_TH;
    const int CX0 = X0 >> 1, CY0 = Y0 >> 1;

    // ---- stage the fine tile, CLAMP applied here (Gaussian's BoundaryCondition), so phase 1 is branch-free.
    // Tiles whose staged columns lie inside the image (block-uniform test) take 16-byte loads with only the row
    // index clamped; all loads of a thread are issued before the first shared-memory store.
    {
        constexpr int VPR = FD_FCOLS / 4;
        constexpr int NV = FROWS * VPR, PER = (NV + FD_NT - 1) / FD_NT;
        const int xs = X0 - 4, ys = Y0 - 1 - H;
        if (!COHERENT && tmap && xs >= 0 && xs + FD_FCOLS <= p.fw && ys >= -p.fgt && ys + FROWS - 1 <= p.fh - 1 + p.fgb) {
            // interior tile (no row of the box needs the CLAMP): one thread requests the 144 x FROWS box with
            // cp.async.bulk.tensor.2d; the map starts at the first ghost row, hence the row coordinate ys + fgt
            if (tid == 0) {
                mbar_init(bar, 1);
                mbar_fence_init();
                mbar_arrive_expect_tx(bar, FROWS * FD_FCOLS * (unsigned)sizeof(float));
                tma_load_2d(ftile, tmap, xs, ys + p.fgt, bar);
            }
            __syncthreads();   // the initialised barrier is visible to the waiting threads
            mbar_wait(bar, 0);
        } else if (xs >= 0 && xs + FD_FCOLS <= p.fw) {
            float4 t[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int v = tid + k * FD_NT;
                if (v < NV) {
                    const int r = v / VPR, c4 = v - r * VPR;
                    const int gy = min(max(ys + r, -p.fgt), p.fh - 1 + p.fgb);
                    t[k] = ld_f4<COHERENT>(reinterpret_cast<const float4 *>(p.fine + (size_t)gy * p.fine_stride + xs) + c4);
                }
            }
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int v = tid + k * FD_NT;
                if (v < NV) reinterpret_cast<float4 *>(ftile)[v] = t[k];
            }
        } else {
            for (int v = tid; v < NV; v += FD_NT) {
                const int r = v / VPR, c4 = v - r * VPR;
                const int gy = min(max(ys + r, -p.fgt), p.fh - 1 + p.fgb);
                const int gx = xs + 4 * c4;
                const float *row = p.fine + (size_t)gy * p.fine_stride;
                float4 t;
                if (gx >= 0 && gx + 3 < p.fw) {
                    t = ld_f4<COHERENT>(reinterpret_cast<const float4 *>(row + gx));
                } else {
                    t.x = ld_f1<COHERENT>(row + min(max(gx, 0), p.fw - 1));
                    t.y = ld_f1<COHERENT>(row + min(max(gx + 1, 0), p.fw - 1));
                    t.z = ld_f1<COHERENT>(row + min(max(gx + 2, 0), p.fw - 1));
                    t.w = ld_f1<COHERENT>(row + min(max(gx + 3, 0), p.fw - 1));
                }
                reinterpret_cast<float4 *>(ftile)[v] = t;
            }
        }
    }
    __syncthreads();

    // ---- phase 1: coarse tile.  thread -> column pair pc (coarse tile columns 2pc, 2pc+1), row group rg (3 rows)
    if (tid < 34 * 6) {
        const int pc = tid % 34, rg = tid / 34;
        float acc[3][2];
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[j][0] = acc[j][1] = 0.0f;
#pragma unroll
        for (int fr = 0; fr < S + 4; ++fr) {
            float w[12];
            const float *row = ftile + (6 * rg + fr) * FD_FCOLS + 4 * pc;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 t = *reinterpret_cast<const float4 *>(row + 4 * q);
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int dy = fr - 2 * j;
                if (dy < 0 || dy >= S) continue;
#pragma unroll
                for (int dx = 0; dx < S; ++dx) {
                    const float cf = p.coef[dy * S + dx];
                    acc[j][0] = __fadd_rn(acc[j][0], __fmul_rn(w[3 - H + dx], cf));
                    acc[j][1] = __fadd_rn(acc[j][1], __fmul_rn(w[5 - H + dx], cf));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int tr = 3 * rg + j;
            *reinterpret_cast<float2 *>(ctile + tr * FD_CCOLS + 2 * pc) = make_float2(acc[j][0], acc[j][1]);
            const int cy = CY0 - 1 + tr;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int tc = 2 * pc + e, cx = CX0 - 1 + tc;
                if (tc >= 1 && tc <= FD_TW / 2 && tr >= 1 && tr <= FD_TH / 2 && cx < p.cw && cy < p.ch)
                    p.coarse[(size_t)cy * p.coarse_stride + cx] = acc[j][e];
            }
        }
    }
    if (!DOG) return;
    __syncthreads();

    // ---- phase 2: lap = fine - LF(coarse) on the thread's 4 x 4 fine block
    const int tx = tid & 31, ty = tid >> 5;
    const int x = X0 + 4 * tx, y = Y0 + 4 * ty;
    if (x < p.fw && y < p.fh) {
        // interior strips (ghost rows present) have no vertical edge: their halo coarse rows -1 / ch were computed
        // in phase 1 from the ghost rows
        const bool edge = X0 == 0 || (Y0 == 0 && p.fgt == 0) || X0 + FD_TW >= p.fw || (Y0 + FD_TH >= p.fh && (p.fgb == 0 || Y0 + FD_TH > p.fh));
        float o[4][4];
        if (!edge) {
            float c[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 a = *reinterpret_cast<const float2 *>(ctile + (2 * ty + j) * FD_CCOLS + 2 * tx);
                const float2 b = *reinterpret_cast<const float2 *>(ctile + (2 * ty + j) * FD_CCOLS + 2 * tx + 2);
                c[j][0] = a.x; c[j][1] = a.y; c[j][2] = b.x; c[j][3] = b.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 f = *reinterpret_cast<const float4 *>(ftile + (4 * ty + j + 1 + H) * FD_FCOLS + 4 + 4 * tx);
                o[j][0] = __fadd_rn(f.x, -lf_block(c, 0, j));
                o[j][1] = __fadd_rn(f.y, -lf_block(c, 1, j));
                o[j][2] = __fadd_rn(f.z, -lf_block(c, 2, j));
                o[j][3] = __fadd_rn(f.w, -lf_block(c, 3, j));
            }
        } else {
            auto at = [&](int cx, int cy) { return ctile[(cy - CY0 + 1) * FD_CCOLS + (cx - CX0 + 1)]; };
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int xx = min(x + i, p.fw - 1), yy = min(y + j, p.fh - 1);
                    const float f = ftile[(yy - Y0 + 1 + H) * FD_FCOLS + 4 + (xx - X0)];
                    o[j][i] = __fadd_rn(f, -lf_half(at, xx, yy, p.cw, p.ch - 1 + (p.fgb > 0 ? 1 : 0), p.fgt == 0));
                }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (y + j >= p.fh) break;
            float *dst = p.lap + (size_t)(y + j) * p.lap_stride + x;
            if (x + 3 < p.fw) {
                __stcs(reinterpret_cast<float4 *>(dst), make_float4(o[j][0], o[j][1], o[j][2], o[j][3]));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (x + i < p.fw) dst[i] = o[j][i];
            }
        }
    }
}

// MINB = resident CTAs per SM the kernel is compiled for.  All-threads staging keeps six 16-byte loads per thread in flight:
// 48 registers, 5 CTAs (492 vs 513 us at level 0 of the 16384^2 pyramid; 4: 64 registers; 6: spills, 526 us).  With TMA staging
// those registers are free on the interior tiles: 40 registers, 6 CTAs, 432 us (5 CTAs: 455 us; the spills sit in the border
// tiles' loader).
template <int S, bool DOG, int MINB>
__global__ void __launch_bounds__(FD_NT, MINB) pyr_down_fused_kernel(const __grid_constant__ PyrFusedParams p, const __grid_constant__ CUtensorMap tmap,
                                                                            const int use_tma) {
    __shared__ __align__(128) float smem[DownSmem<S>::FLOATS];
    __shared__ __align__(8) uint64_t bar;
    pyr_down_tile<S, false, DOG>(p, blockIdx.x, blockIdx.y, smem, smem + DownSmem<S>::FROWS * FD_FCOLS, use_tma ? &tmap : nullptr, &bar);
}

struct PyrUpHalfParams {
    const float *cg, *cl;
    float *fg, *fl;
    int cg_stride, cl_stride, fg_stride, fl_stride;
    int fw, fh, cw, ch;
    int cgt, cgb;   // ghost rows of the coarse planes (row-strip sharding; filled by the halo exchange)
};

// Restore + Blend for the exact-halving case: 4 x 4 fine block per thread, the two 4 x 4 coarse neighbourhoods
// come through the read-only path (each coarse value is used by ~16 fine pixels: L1/L2 hits), lap is read once
// with 16-byte streaming loads, both outputs are written with 16-byte stores.
template <bool COHERENT>
__device__ __forceinline__ void pyr_up_tile(const PyrUpHalfParams &p, const int bx, const int by) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = bx * 128 + 4 * tx, y = by * 32 + 4 * ty;
    if (x >= p.fw || y >= p.fh) return;
    float l[4][4], og[4][4], ol[4][4];
    const bool full = x + 3 < p.fw && y + 3 < p.fh;
    if (full) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 t = __ldcs(reinterpret_cast<const float4 *>(p.fl + (size_t)(y + j) * p.fl_stride + x));
            l[j][0] = t.x; l[j][1] = t.y; l[j][2] = t.z; l[j][3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) l[j][i] = __ldcs(p.fl + (size_t)min(y + j, p.fh - 1) * p.fl_stride + min(x + i, p.fw - 1));
    }
    const int cy_max = p.ch - 1 + p.cgb;
    if (full && x > 0 && (y > 0 || p.cgt > 0)) {
        const int m = x >> 1, k = y >> 1;
        float g[4][4], c[4][4];
        const int c0 = m - 1, c3 = min(m + 2, p.cw - 1), c2 = min(m + 1, p.cw - 1);  // m even: (m, m+1) is an 8-byte aligned pair
        const bool pair = (p.cg_stride % 2 == 0) && (p.cl_stride % 2 == 0) && m + 1 < p.cw &&
                          ((reinterpret_cast<uintptr_t>(p.cg) | reinterpret_cast<uintptr_t>(p.cl)) % 8 == 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cy = min(k - 1 + j, cy_max);   // >= cy_min: y == 0 takes this path only with ghost rows
            const float *rg = p.cg + (ptrdiff_t)cy * p.cg_stride, *rl = p.cl + (ptrdiff_t)cy * p.cl_stride;
            g[j][0] = ld_f1<COHERENT>(rg + c0); c[j][0] = ld_f1<COHERENT>(rl + c0);
            if (pair) {
                const float2 a = ld_f2<COHERENT>(reinterpret_cast<const float2 *>(rg + m)), b = ld_f2<COHERENT>(reinterpret_cast<const float2 *>(rl + m));
                g[j][1] = a.x; g[j][2] = a.y; c[j][1] = b.x; c[j][2] = b.y;
            } else {
                g[j][1] = ld_f1<COHERENT>(rg + m); g[j][2] = ld_f1<COHERENT>(rg + c2); c[j][1] = ld_f1<COHERENT>(rl + m); c[j][2] = ld_f1<COHERENT>(rl + c2);
            }
            g[j][3] = ld_f1<COHERENT>(rg + c3); c[j][3] = ld_f1<COHERENT>(rl + c3);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                og[j][i] = __fadd_rn(lf_block(g, i, j), l[j][i]);
                ol[j][i] = __fadd_rn(lf_block(c, i, j), __fdiv_rn(l[j][i], 2.0f));
            }
    } else {
        auto atg = [&](int cx, int cy) { return ld_f1<COHERENT>(p.cg + (ptrdiff_t)cy * p.cg_stride + cx); };
        auto atl = [&](int cx, int cy) { return ld_f1<COHERENT>(p.cl + (ptrdiff_t)cy * p.cl_stride + cx); };
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int xx = min(x + i, p.fw - 1), yy = min(y + j, p.fh - 1);
                og[j][i] = __fadd_rn(lf_half(atg, xx, yy, p.cw, cy_max, p.cgt == 0), l[j][i]);
                ol[j][i] = __fadd_rn(lf_half(atl, xx, yy, p.cw, cy_max, p.cgt == 0), __fdiv_rn(l[j][i], 2.0f));
            }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (y + j >= p.fh) break;
        float *dg = p.fg + (size_t)(y + j) * p.fg_stride + x, *dl = p.fl + (size_t)(y + j) * p.fl_stride + x;
        if (x + 3 < p.fw) {
            __stcs(reinterpret_cast<float4 *>(dg), make_float4(og[j][0], og[j][1], og[j][2], og[j][3]));
            __stcs(reinterpret_cast<float4 *>(dl), make_float4(ol[j][0], ol[j][1], ol[j][2], ol[j][3]));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x + i < p.fw) { dg[i] = og[j][i]; dl[i] = ol[j][i]; }
        }
    }
}

__global__ void __launch_bounds__(256) pyr_up_half_kernel(const __grid_constant__ PyrUpHalfParams p) { pyr_up_tile<false>(p, blockIdx.x, blockIdx.y); }

// DifferenceOfGaussian on its own for the exact-halving case: lap = fine - LF(coarse), 4 x 4 fine block per thread, the
// coarse neighbourhood through the read-only path (ghost rows of the coarse plane like pyr_up_half_kernel).  Same
// arithmetic as phase 2 of the fused down kernel.  PyrUpHalfParams: cg = coarse, fg = fine (read), fl = lap (written).
__global__ void __launch_bounds__(256) pyr_dog_half_kernel(const __grid_constant__ PyrUpHalfParams p) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = blockIdx.x * 128 + 4 * tx, y = blockIdx.y * 32 + 4 * ty;
    if (x >= p.fw || y >= p.fh) return;
    float f[4][4], o[4][4];
    const bool full = x + 3 < p.fw && y + 3 < p.fh;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (full) {
            const float4 t = __ldcs(reinterpret_cast<const float4 *>(p.fg + (size_t)(y + j) * p.fg_stride + x));
            f[j][0] = t.x; f[j][1] = t.y; f[j][2] = t.z; f[j][3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) f[j][i] = __ldg(p.fg + (size_t)min(y + j, p.fh - 1) * p.fg_stride + min(x + i, p.fw - 1));
        }
    }
    const int cy_max = p.ch - 1 + p.cgb;
    if (full && x > 0 && (y > 0 || p.cgt > 0)) {
        const int m = x >> 1, k = y >> 1;
        float g[4][4];
        const int c0 = m - 1, c3 = min(m + 2, p.cw - 1), c2 = min(m + 1, p.cw - 1);
        const bool pair = (p.cg_stride % 2 == 0) && m + 1 < p.cw && (reinterpret_cast<uintptr_t>(p.cg) % 8 == 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cy = min(k - 1 + j, cy_max);
            const float *rg = p.cg + (ptrdiff_t)cy * p.cg_stride;
            g[j][0] = __ldg(rg + c0);
            if (pair) {
                const float2 a = __ldg(reinterpret_cast<const float2 *>(rg + m));
                g[j][1] = a.x; g[j][2] = a.y;
            } else {
                g[j][1] = __ldg(rg + m); g[j][2] = __ldg(rg + c2);
            }
            g[j][3] = __ldg(rg + c3);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) o[j][i] = __fadd_rn(f[j][i], -lf_block(g, i, j));
    } else {
        auto atg = [&](int cx, int cy) { return __ldg(p.cg + (ptrdiff_t)cy * p.cg_stride + cx); };
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int xx = min(x + i, p.fw - 1), yy = min(y + j, p.fh - 1);
                o[j][i] = __fadd_rn(f[j][i], -lf_half(atg, xx, yy, p.cw, cy_max, p.cgt == 0));
            }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (y + j >= p.fh) break;
        float *dl = p.fl + (size_t)(y + j) * p.fl_stride + x;
        if (x + 3 < p.fw) {
            __stcs(reinterpret_cast<float4 *>(dl), make_float4(o[j][0], o[j][1], o[j][2], o[j][3]));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x + i < p.fw) dl[i] = o[j][i];
        }
    }
}

// ================================================================================================
// The coarse end of a pyramid in ONE launch.  Below ~1024^2 every level transition is a few microseconds of work
// behind a launch of its own (levels 4-7 of the 16384^2 pyramid: 6 launches, ~60 us of a 1.55 ms traversal, and the
// part every rank repeats when the pyramid is sharded).  This cooperative kernel walks the whole tail -- way down
// through all transitions, then way up -- with a grid-wide barrier between transitions: the same tile functions as
// the per-level kernels (same arithmetic, same order: bit-identical), tiles handed out round-robin to the resident
// CTAs, loads through the coherent path because a level written in one phase is read in the next.
// ================================================================================================
constexpr int kMaxCoarseTransitions = 7;
struct PyrCoarseParams {
    int n;                                   // transitions: levels 0 .. n of the small pyramid
    PyrFusedParams down[kMaxCoarseTransitions];   // down[t]: level t -> t + 1
    PyrUpHalfParams up[kMaxCoarseTransitions];    // up[t]  : level t + 1 -> t
    unsigned *bar;                           // [0] arrivals (monotonic within the launch), [1] exit ticket
};

__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        __threadfence();                     // this CTA's stores are visible before it arrives
        atomicAdd(bar, 1u);
        const unsigned target = epoch * gridDim.x;
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

template <int S>
__global__ void __launch_bounds__(FD_NT) pyr_coarse_kernel(const __grid_constant__ PyrCoarseParams p) {
    __shared__ __align__(16) float smem[DownSmem<S>::FLOATS];
    float *ftile = smem, *ctile = smem + DownSmem<S>::FROWS * FD_FCOLS;
    unsigned epoch = 0;
    for (int t = 0; t < p.n; ++t) {          // way down
        const PyrFusedParams &q = p.down[t];
        const int ntx = (q.fw + FD_TW - 1) / FD_TW, nty = (q.fh + FD_TH - 1) / FD_TH;
        for (int i = blockIdx.x; i < ntx * nty; i += gridDim.x) {
            pyr_down_tile<S, true>(q, i % ntx, i / ntx, ftile, ctile);
            __syncthreads();                 // the tiles are reused by the next iteration
        }
        grid_barrier(p.bar, epoch);
    }
    for (int t = p.n - 1; t >= 0; --t) {     // way up
        const PyrUpHalfParams &q = p.up[t];
        const int ntx = (q.fw + 127) / 128, nty = (q.fh + 31) / 32;
        for (int i = blockIdx.x; i < ntx * nty; i += gridDim.x) pyr_up_tile<true>(q, i % ntx, i / ntx);
        if (t > 0) grid_barrier(p.bar, epoch);
    }
    // the last CTA to leave resets the barrier for the next launch (every CTA has passed the last barrier by then)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.bar + 1, 1u) == gridDim.x - 1) {
            p.bar[0] = 0;
            p.bar[1] = 0;
            __threadfence();
        }
    }
}

// region base pointer + alignment test of a plane for the 16-byte paths
static inline float *region_base(const PlaneRef &r) { return r.p + (size_t)r.oy * r.stride + r.ox; }
static inline bool vec_ok(const PlaneRef &r) { return r.stride % 4 == 0 && reinterpret_cast<uintptr_t>(region_base(r)) % 16 == 0; }

}  // namespace hb

using namespace hb;

extern "C" int hb_pyr_down(const hb_pyr_down_desc *d, void *stream) {
    HB_REQUIRE(d && d->coef_f32, HB_ERR_INVALID, "hb_pyr_down: null descriptor / mask");
    hb_view fine = norm_view(d->fine), coarse = norm_view(d->coarse);
    HB_REQUIRE(view_ok(fine) && view_ok(coarse) && fine.dtype == HB_F32 && coarse.dtype == HB_F32, HB_ERR_INVALID, "hb_pyr_down: needs valid f32 views");
    HB_REQUIRE(d->size == 3 || d->size == 5 || d->size == 7, HB_ERR_UNSUPPORTED, "hb_pyr_down: mask size %d unsupported (3,5,7); no CPU fallback", d->size);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = HB_OK;
    if (!d->tmp.data && fine.width == 2 * coarse.width && fine.height == 2 * coarse.height) {
        // exact-halving transition: blur + subsample (+ DoG when lap_fine is given) in one kernel
        const bool dog = d->lap_fine.data != nullptr;
        hb_view lap = dog ? norm_view(d->lap_fine) : fine;
        HB_REQUIRE(!dog || (view_ok(lap) && lap.dtype == HB_F32 && lap.width == fine.width && lap.height == fine.height), HB_ERR_INVALID,
                   "hb_pyr_down: lap_fine must be an f32 view of the fine level's size");
        const PlaneRef f = plane_of(fine), c = plane_of(coarse), l = plane_of(lap);
        if (vec_ok(f) && vec_ok(l) && c.stride % 2 == 0 && reinterpret_cast<uintptr_t>(region_base(c)) % 8 == 0) {
            PyrFusedParams p;
            memset(&p, 0, sizeof(p));
            p.fine = region_base(f); p.coarse = region_base(c); p.lap = dog ? region_base(l) : nullptr;
            p.fine_stride = f.stride; p.coarse_stride = c.stride; p.lap_stride = l.stride;
            p.fw = f.w; p.fh = f.h; p.cw = c.w; p.ch = c.h;
            // ghost rows (row-strip sharding): the coarse halo row above needs fine rows down to -(H+1), the one
            // below up to fh + H + 1; fewer ghost rows than that cannot reproduce the unsharded result
            p.fgt = fine.ghost_top; p.fgb = fine.ghost_bottom;
            HB_REQUIRE((p.fgt == 0 || p.fgt >= d->size / 2 + 1) && (p.fgb == 0 || p.fgb >= d->size / 2 + 2), HB_ERR_INVALID,
                       "hb_pyr_down: a sharded fine level needs >= %d ghost rows above and >= %d below", d->size / 2 + 1, d->size / 2 + 2);
            for (int k = 0; k < d->size * d->size; ++k) p.coef[k] = d->coef_f32[k];
            OpScope scope(s, dog ? "hb_pyr_down(fused blur+subsample+DoG)" : "hb_pyr_down(fused blur+subsample)");
            dim3 grid((p.fw + FD_TW - 1) / FD_TW, (p.fh + FD_TH - 1) / FD_TH);
            // interior tiles are staged by TMA (one request per CTA); the tensor map starts at the first ghost row
            CUtensorMap tmap;
            memset(&tmap, 0, sizeof(tmap));
            static int no_tma = -1;
            if (no_tma < 0) { const char *e = getenv("HB_PYR_NO_TMA"); no_tma = (e && atoi(e)) ? 1 : 0; }   // A/B knob
            const int frows = 35 + 2 * (d->size / 2);
            const int use_tma = !no_tma && make_tile_map(&tmap, p.fine - (ptrdiff_t)p.fgt * p.fine_stride, HB_F32, p.fw, p.fh + p.fgt + p.fgb, p.fine_stride, FD_FCOLS, frows);
#define HB_PD(S_, DOG_) (use_tma ? pyr_down_fused_kernel<S_, DOG_, 6><<<grid, FD_NT, 0, s>>>(p, tmap, 1) : pyr_down_fused_kernel<S_, DOG_, 5><<<grid, FD_NT, 0, s>>>(p, tmap, 0))
            if (dog) {
                if (d->size == 3) HB_PD(3, true);
                else if (d->size == 5) HB_PD(5, true);
                else HB_PD(7, true);
            } else {
                if (d->size == 3) HB_PD(3, false);
                else if (d->size == 5) HB_PD(5, false);
                else HB_PD(7, false);
            }
#undef HB_PD
            g_launches++;
            return scope.finish();
        }
    }
    HB_REQUIRE(fine.ghost_top == 0 && fine.ghost_bottom == 0, HB_ERR_UNSUPPORTED,
               "hb_pyr_down: ghost rows (row-strip sharding) need the fused exact-halving path (even level sizes, 16-byte aligned rows, no tmp)");
    if (d->tmp.data) {
        // unfused form (tmp is part of the visible state): blur into tmp, then NN subsample
        hb_local_desc l;
        memset(&l, 0, sizeof(l));
        l.in = fine; l.out = norm_view(d->tmp);
        l.kind = HB_LOCAL_CONVOLVE; l.reduce_mode = HB_REDUCE_SUM; l.tap = HB_TAP_MUL; l.acc_dtype = HB_F32;
        l.size_x = l.size_y = d->size; l.coef_f32 = d->coef_f32; l.boundary = HB_BOUNDARY_CLAMP; l.epilogue = HB_EPI_CAST;
        rc = hb_local_op(&l, stream);
        if (rc) return rc;
        hb_point_desc pt;
        memset(&pt, 0, sizeof(pt));
        pt.in[0] = norm_view(d->tmp); pt.interp[0] = HB_INTERP_NN; pt.n_in = 1; pt.out = coarse; pt.op = HB_POINT_COPY;
        rc = hb_point_op(&pt, stream);
        if (rc) return rc;
    } else {
        PyrDownParams p;
        memset(&p, 0, sizeof(p));
        p.fine = plane_of(fine); p.coarse = plane_of(coarse); p.size = d->size;
        for (int k = 0; k < d->size * d->size; ++k) p.coef[k] = d->coef_f32[k];
        OpScope scope(s, "hb_pyr_down(blur+subsample)");
        dim3 grid((coarse.width + 31) / 32, (coarse.height + 7) / 8);
        if (d->size == 3) pyr_blur_subsample_kernel<3><<<grid, dim3(32, 8), 0, s>>>(p);
        else if (d->size == 5) pyr_blur_subsample_kernel<5><<<grid, dim3(32, 8), 0, s>>>(p);
        else pyr_blur_subsample_kernel<7><<<grid, dim3(32, 8), 0, s>>>(p);
        g_launches++;
        rc = scope.finish();
        if (rc) return rc;
    }
    if (d->lap_fine.data) {
        hb_view lap = norm_view(d->lap_fine);
        HB_REQUIRE(view_ok(lap) && lap.dtype == HB_F32 && lap.width == fine.width && lap.height == fine.height, HB_ERR_INVALID,
                   "hb_pyr_down: lap_fine must be an f32 view of the fine level's size");
        PyrDogParams p{plane_of(fine), plane_of(coarse), plane_of(lap)};
        OpScope scope(s, "hb_pyr_down(DoG)");
        dim3 grid((lap.width + 31) / 32, (lap.height + 31) / 32);
        pyr_dog_kernel<<<grid, dim3(32, 8), 0, s>>>(p);
        g_launches++;
        rc = scope.finish();
    }
    return rc;
}

extern "C" int hb_pyr_up(const hb_pyr_up_desc *d, void *stream) {
    HB_REQUIRE(d, HB_ERR_INVALID, "hb_pyr_up: null descriptor");
    hb_view cg = norm_view(d->coarse_gaus), cl = norm_view(d->coarse_lap), fg = norm_view(d->fine_gaus), fl = norm_view(d->fine_lap);
    HB_REQUIRE(view_ok(cg) && view_ok(cl) && view_ok(fg) && view_ok(fl), HB_ERR_INVALID, "hb_pyr_up: malformed view");
    HB_REQUIRE(cg.dtype == HB_F32 && cl.dtype == HB_F32 && fg.dtype == HB_F32 && fl.dtype == HB_F32, HB_ERR_UNSUPPORTED, "hb_pyr_up: f32 only; no CPU fallback");
    HB_REQUIRE(fg.width == fl.width && fg.height == fl.height, HB_ERR_INVALID, "hb_pyr_up: fine gaus / lap sizes differ");
    cudaStream_t s = (cudaStream_t)stream;
    {
        const PlaneRef g = plane_of(cg), l = plane_of(cl), G = plane_of(fg), L = plane_of(fl);
        if (G.w == 2 * g.w && G.h == 2 * g.h && L.w == 2 * l.w && L.h == 2 * l.h && g.w == l.w && g.h == l.h && vec_ok(G) && vec_ok(L)) {
            HB_REQUIRE(cg.ghost_top == cl.ghost_top && cg.ghost_bottom == cl.ghost_bottom, HB_ERR_INVALID, "hb_pyr_up: coarse gaus / lap ghost rows differ");
            PyrUpHalfParams q{region_base(g), region_base(l), region_base(G), region_base(L), g.stride, l.stride, G.stride, L.stride, G.w, G.h, g.w, g.h,
                              cg.ghost_top, cg.ghost_bottom};
            OpScope scope(s, "hb_pyr_up(exact halving)");
            dim3 grid((G.w + 127) / 128, (G.h + 31) / 32);
            pyr_up_half_kernel<<<grid, 256, 0, s>>>(q);
            g_launches++;
            return scope.finish();
        }
    }
    HB_REQUIRE(cg.ghost_top == 0 && cg.ghost_bottom == 0 && cl.ghost_top == 0 && cl.ghost_bottom == 0, HB_ERR_UNSUPPORTED,
               "hb_pyr_up: ghost rows (row-strip sharding) need the exact-halving path (even level sizes, 16-byte aligned rows)");
    PyrUpParams p{plane_of(cg), plane_of(cl), plane_of(fg), plane_of(fl)};
    OpScope scope(s, "hb_pyr_up");
    dim3 grid((fg.width + 31) / 32, (fg.height + 31) / 32);
    pyr_up_kernel<<<grid, dim3(32, 8), 0, s>>>(p);
    g_launches++;
    return scope.finish();
}

// DifferenceOfGaussian of the pyramid sample on its own: lap_fine = fine - LF(coarse)
extern "C" int hb_pyr_dog(const hb_pyr_dog_desc *d, void *stream) {
    HB_REQUIRE(d, HB_ERR_INVALID, "hb_pyr_dog: null descriptor");
    hb_view fine = norm_view(d->fine), coarse = norm_view(d->coarse), lap = norm_view(d->lap_fine);
    HB_REQUIRE(view_ok(fine) && view_ok(coarse) && view_ok(lap), HB_ERR_INVALID, "hb_pyr_dog: malformed view");
    HB_REQUIRE(fine.dtype == HB_F32 && coarse.dtype == HB_F32 && lap.dtype == HB_F32, HB_ERR_UNSUPPORTED, "hb_pyr_dog: f32 only; no CPU fallback");
    HB_REQUIRE(lap.width == fine.width && lap.height == fine.height, HB_ERR_INVALID, "hb_pyr_dog: lap_fine must have the fine level's size");
    cudaStream_t s = (cudaStream_t)stream;
    const PlaneRef f = plane_of(fine), c = plane_of(coarse), l = plane_of(lap);
    if (f.w == 2 * c.w && f.h == 2 * c.h && vec_ok(f) && vec_ok(l)) {
        PyrUpHalfParams q{region_base(c), region_base(c), region_base(f), region_base(l), c.stride, c.stride, f.stride, l.stride, f.w, f.h, c.w, c.h,
                          coarse.ghost_top, coarse.ghost_bottom};
        OpScope scope(s, "hb_pyr_dog(exact halving)");
        dim3 grid((f.w + 127) / 128, (f.h + 31) / 32);
        pyr_dog_half_kernel<<<grid, 256, 0, s>>>(q);
        g_launches++;
        return scope.finish();
    }
    HB_REQUIRE(coarse.ghost_top == 0 && coarse.ghost_bottom == 0, HB_ERR_UNSUPPORTED,
               "hb_pyr_dog: ghost rows (row-strip sharding) need the exact-halving path (even level sizes, 16-byte aligned rows)");
    PyrDogParams p{plane_of(fine), plane_of(coarse), plane_of(lap)};
    OpScope scope(s, "hb_pyr_dog");
    dim3 grid((lap.width + 31) / 32, (lap.height + 31) / 32);
    pyr_dog_kernel<<<grid, dim3(32, 8), 0, s>>>(p);
    g_launches++;
    return scope.finish();
}

// ---- the coarse end of a pyramid in one cooperative launch -------------------------------------------------------
namespace hb {
static std::mutex g_bar_mutex;
static std::map<std::pair<int, cudaStream_t>, unsigned *> g_bars;   // grid-barrier words, one pair per (device, stream)

static int coarse_barrier(unsigned **out, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_bar_mutex);
    auto it = g_bars.find({dev, s});
    if (it == g_bars.end()) {
        HB_REQUIRE(!stream_is_capturing(s), HB_ERR_INVALID,
                   "hb_pyr_traverse_coarse: the barrier words of this stream cannot be allocated inside a capture; begin it with hb_graph_begin");
        unsigned *b = nullptr;
        int rc = check_cuda(cudaMalloc(&b, 2 * sizeof(unsigned)), "cudaMalloc(grid barrier)");
        if (!rc) rc = check_cuda(cudaMemset(b, 0, 2 * sizeof(unsigned)), "cudaMemset(grid barrier)");
        if (rc) return rc;
        it = g_bars.emplace(std::make_pair(dev, s), b).first;
    }
    *out = it->second;
    return HB_OK;
}
int reserve_pyramid_scratch(cudaStream_t s) {
    unsigned *b = nullptr;
    return coarse_barrier(&b, s);
}
}  // namespace hb

extern "C" int hb_pyr_traverse_coarse(const hb_pyr_coarse_desc *d, void *stream) {
    HB_REQUIRE(d && d->coef_f32, HB_ERR_INVALID, "hb_pyr_traverse_coarse: null descriptor / mask");
    HB_REQUIRE(d->levels >= 2 && d->levels <= kMaxCoarseTransitions + 1, HB_ERR_UNSUPPORTED, "hb_pyr_traverse_coarse: 2..%d levels", kMaxCoarseTransitions + 1);
    HB_REQUIRE(d->size == 3 || d->size == 5 || d->size == 7, HB_ERR_UNSUPPORTED, "hb_pyr_traverse_coarse: mask size %d unsupported (3,5,7); no CPU fallback", d->size);
    PyrCoarseParams p;
    memset(&p, 0, sizeof(p));
    p.n = d->levels - 1;
    PlaneRef g[kMaxCoarseTransitions + 1], l[kMaxCoarseTransitions + 1];
    for (int i = 0; i < d->levels; ++i) {
        const hb_view vg = norm_view(d->gaus[i]), vl = norm_view(d->lap[i]);
        HB_REQUIRE(view_ok(vg) && view_ok(vl) && vg.dtype == HB_F32 && vl.dtype == HB_F32, HB_ERR_INVALID, "hb_pyr_traverse_coarse: level %d needs valid f32 views", i);
        HB_REQUIRE(vg.ghost_top == 0 && vg.ghost_bottom == 0 && vl.ghost_top == 0 && vl.ghost_bottom == 0, HB_ERR_UNSUPPORTED,
                   "hb_pyr_traverse_coarse: ghost rows are not supported (the coarse levels are replicated, not sharded)");
        HB_REQUIRE(vg.width == vl.width && vg.height == vl.height, HB_ERR_INVALID, "hb_pyr_traverse_coarse: level %d gaus / lap sizes differ", i);
        g[i] = plane_of(vg); l[i] = plane_of(vl);
    }
    for (int t = 0; t < p.n; ++t) {
        const PlaneRef &f = g[t], &c = g[t + 1], &lf = l[t], &lc = l[t + 1];
        // the fused tile functions need exact halving and 16-byte (fine) / 8-byte (coarse) aligned rows; anything else is the
        // caller's per-level path (hb_pyr_down / hb_pyr_up)
        HB_REQUIRE(f.w == 2 * c.w && f.h == 2 * c.h && vec_ok(f) && vec_ok(lf) && c.stride % 2 == 0 && lc.stride % 2 == 0 &&
                       reinterpret_cast<uintptr_t>(region_base(c)) % 8 == 0 && reinterpret_cast<uintptr_t>(region_base(lc)) % 8 == 0,
                   HB_ERR_UNSUPPORTED, "hb_pyr_traverse_coarse: transition %d is not an aligned exact halving; use hb_pyr_down / hb_pyr_up per level", t);
        PyrFusedParams &q = p.down[t];
        q.fine = region_base(f); q.coarse = region_base(c); q.lap = region_base(lf);
        q.fine_stride = f.stride; q.coarse_stride = c.stride; q.lap_stride = lf.stride;
        q.fw = f.w; q.fh = f.h; q.cw = c.w; q.ch = c.h;
        for (int k = 0; k < d->size * d->size; ++k) q.coef[k] = d->coef_f32[k];
        p.up[t] = PyrUpHalfParams{region_base(c), region_base(lc), region_base(f), region_base(lf), c.stride, lc.stride, f.stride, lf.stride, f.w, f.h, c.w, c.h, 0, 0};
    }
    cudaStream_t s = (cudaStream_t)stream;
    int rc = coarse_barrier(&p.bar, s);
    if (rc) return rc;
    const void *kern = d->size == 3 ? (const void *)pyr_coarse_kernel<3> : d->size == 5 ? (const void *)pyr_coarse_kernel<5> : (const void *)pyr_coarse_kernel<7>;
    // every CTA must be resident (grid-wide barrier): at most one CTA per SM, never more than the first level has tiles
    const int tiles0 = ((p.down[0].fw + FD_TW - 1) / FD_TW) * ((p.down[0].fh + FD_TH - 1) / FD_TH);
    static int per_sm = -1;   // resident CTAs per SM the grid may use (HB_COARSE_CTAS_PER_SM, tuning knob)
    static int max_per_sm[3] = {0, 0, 0};
    if (per_sm < 0) {
        const char *e = getenv("HB_COARSE_CTAS_PER_SM");
        per_sm = e ? atoi(e) : 2;
        if (per_sm < 1) per_sm = 1;
    }
    const int ki = d->size == 3 ? 0 : d->size == 5 ? 1 : 2;
    if (max_per_sm[ki] == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, FD_NT, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
        max_per_sm[ki] = nb;
    }
    const int cap = sm_count() * (per_sm < max_per_sm[ki] ? per_sm : max_per_sm[ki]);
    int grid = cap < tiles0 ? cap : tiles0;
    if (grid < 1) grid = 1;
    OpScope scope(s, "hb_pyr_traverse_coarse");
    void *args[] = {&p};
    rc = check_cuda(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(FD_NT), args, 0, s), "cudaLaunchCooperativeKernel(pyr_coarse_kernel)");
    if (rc) return rc;
    g_launches++;
    return scope.finish();
}
