// hb_harris.cu -- fused Harris corner detector (uchar -> uchar) for sm_100a.
//
// The reference sample (samples-public/3_Preprocessing/Harris_Corner/src/main.cpp:230-305) runs
// nine kernels over eight full-size images: Sobel dx, dy (uchar -> short, /6), Square1 x2, Square2,
// three 3x3 binomial Gaussians (short -> short, /16, CLAMP on the *intermediate* images) and the
// HarrisCorner point operator -- 39 bytes of HBM traffic per pixel.  Here the whole pipeline is one
// kernel with a 5x5 receptive field: 2 bytes per pixel.
//
// Exactness ("Fusing across a boundary-handled intermediate"): the Gaussian
// stage applies CLAMP to the coordinates of the intermediate images, so the halo ring of
// intermediates is evaluated at the CLAMPED in-image position (and the Sobel stage applies its own
// CLAMP there) -- never at out-of-image coordinates.  All integer stages are exact; short stores of
// the reference cannot wrap here (|dx| <= 127, products <= 16129) and the response uses separately
// rounded float ops like the C++ expression.  Results are bit-identical to the unfused pipeline.
#include "hb_common.cuh"
#include "hb_internal.h"
#include "hb_tma.cuh"

#include <cstdlib>
#include <cstring>

namespace hb {

struct HarrisParams {
    const uchar *in;
    uchar *out;
    int in_stride, in_iw, in_ih;
    Window win;  // CLAMP window of the input accessor
    int in_ox, in_oy;
    int out_stride, out_ox, out_oy, w, h;
    float k, threshold;
};

constexpr int HTW = 128, HRPT = 4, HBX = 32, HBY = 8, HTH = HBY * HRPT;
constexpr int HIN_COLS = HTW + 8, HIN_ROWS = HTH + 4;        // input tile, halo 2 (4 columns staged for alignment)
constexpr int HMID_COLS = HTW + 8, HMID_ROWS = HTH + 2;      // intermediates, halo 1 (column 0 <-> tile x = -4)

__global__ void __launch_bounds__(HBX *HBY) harris_fused_kernel(const __grid_constant__ HarrisParams p) {
    __shared__ __align__(16) int tin[HIN_ROWS * HIN_COLS];
    __shared__ __align__(16) short sxx[HMID_ROWS * HMID_COLS];
    __shared__ __align__(16) short syy[HMID_ROWS * HMID_COLS];
    __shared__ __align__(16) short sxy[HMID_ROWS * HMID_COLS];

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * HBX + tx;
    const int gx0 = blockIdx.x * HTW, gy0 = blockIdx.y * HTH;

    // stage A: input tile (column 0 <-> IS-relative x = gx0 - 4, row 0 <-> y = gy0 - 2)
    stage_tile<uchar, int, HIN_ROWS, HIN_COLS, HBX * HBY>(tin, p.in, p.in_stride, p.in_iw, p.in_ih, p.win, (uchar)0,
                                                         p.in_ox + gx0 - 4, p.in_oy + gy0 - 2, tid);
    __syncthreads();

    // stage B: structure-tensor products at tile positions [-1, HTW] x [-1, HTH], each evaluated at the
    // CLAMPED position of the intermediate image (w x h)
    for (int q = tid; q < HMID_ROWS * (HTW + 2); q += HBX * HBY) {
        const int jy = q / (HTW + 2) - 1, jx = q - (jy + 1) * (HTW + 2) - 1;
        int cx = gx0 + jx, cy = gy0 + jy;
        // the intermediates live on the same extent as the input window: [0, w) horizontally and, with ghost
        // rows (row-strip sharding), [-ghost_top, h + ghost_bottom) vertically -- CLAMP only at the global edge
        cx = min(max(cx, 0), p.w - 1) - gx0;  // tile-local clamped position
        cy = min(max(cy, p.win.lo_y - p.in_oy), p.win.hi_y - 1 - p.in_oy) - gy0;
        const int *c = tin + (cy + 2) * HIN_COLS + (cx + 4);
        const int a00 = c[-HIN_COLS - 1], a01 = c[-HIN_COLS], a02 = c[-HIN_COLS + 1];
        const int a10 = c[-1], a12 = c[1];
        const int a20 = c[HIN_COLS - 1], a21 = c[HIN_COLS], a22 = c[HIN_COLS + 1];
        const int dx = ((a02 - a00) + (a12 - a10) + (a22 - a20)) / 6;  // short sum / 6
        const int dy = ((a20 - a00) + (a21 - a01) + (a22 - a02)) / 6;
        const int o = (jy + 1) * HMID_COLS + (jx + 4);
        sxx[o] = (short)(dx * dx);
        syy[o] = (short)(dy * dy);
        sxy[o] = (short)(dx * dy);
    }
    __syncthreads();

    // stage C: 3x3 binomial on the three planes (int accumulate, /16) + response; 4 px x 4 rows per thread
    const int r0 = ty * HRPT;
    int gxx[HRPT][4], gyy[HRPT][4], gxy[HRPT][4];
#pragma unroll
    for (int r = 0; r < HRPT; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) gxx[r][i] = gyy[r][i] = gxy[r][i] = 0;
#pragma unroll
    for (int ir = 0; ir < HRPT + 2; ++ir) {
        // horizontal [1 2 1] of this intermediate row for the thread's 4 pixels (columns 4tx+3 .. 4tx+8)
        const int base = (r0 + ir) * HMID_COLS + 4 * tx + 2;  // even index -> 4-byte aligned short2 loads
        int hx[4], hy[4], hxy[4];
        {
            short v[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const short2 t = *reinterpret_cast<const short2 *>(sxx + base + 2 * q); v[2 * q] = t.x; v[2 * q + 1] = t.y; }
#pragma unroll
            for (int i = 0; i < 4; ++i) hx[i] = v[i + 1] + 2 * v[i + 2] + v[i + 3];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const short2 t = *reinterpret_cast<const short2 *>(syy + base + 2 * q); v[2 * q] = t.x; v[2 * q + 1] = t.y; }
#pragma unroll
            for (int i = 0; i < 4; ++i) hy[i] = v[i + 1] + 2 * v[i + 2] + v[i + 3];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const short2 t = *reinterpret_cast<const short2 *>(sxy + base + 2 * q); v[2 * q] = t.x; v[2 * q + 1] = t.y; }
#pragma unroll
            for (int i = 0; i < 4; ++i) hxy[i] = v[i + 1] + 2 * v[i + 2] + v[i + 3];
        }
#pragma unroll
        for (int r = 0; r < HRPT; ++r) {
            const int dy = ir - r;
            if (dy < 0 || dy > 2) continue;
            const int wgt = dy == 1 ? 2 : 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                gxx[r][i] += wgt * hx[i];
                gyy[r][i] += wgt * hy[i];
                gxy[r][i] += wgt * hxy[i];
            }
        }
    }

    const int gx = gx0 + 4 * tx;
#pragma unroll
    for (int r = 0; r < HRPT; ++r) {
        const int gy = gy0 + r0 + r;
        if (gy >= p.h || gx >= p.w) continue;
        uchar o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = (int)(short)(gxx[r][i] / 16), y = (int)(short)(gyy[r][i] / 16), xy = (int)(short)(gxy[r][i] / 16);
            const float det = (float)(x * y - xy * xy);
            const float s = (float)(x + y);
            const float tr = __fmul_rn(__fmul_rn(p.k, s), s);
            o[i] = __fadd_rn(det, -tr) > p.threshold ? 1 : 0;
        }
        uchar *dst = p.out + (size_t)(p.out_oy + gy) * p.out_stride + p.out_ox + gx;
        if (gx + 3 < p.w && (reinterpret_cast<uintptr_t>(dst) % 4 == 0)) {
            store4(dst, o);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (gx + i < p.w) dst[i] = o[i];
        }
    }
}


// ================================================================================================
// Version 2 of the fused kernel: the same arithmetic on the integer dot-product instructions.
//   stage A  input tile as BYTES (5 KB instead of 20 KB), CLAMP applied by the loader;
//   stage B  per intermediate position  dx*6 and dy*6 are five IDP.4A (dp4a.u32.s32) on byte windows of the three
//            input rows -- the 3x3 Sobel sums fold into the accumulator operand, no adds; d/6 by multiply-high
//            (the compiler's multiply-high); products stored as packed 16-bit pairs (sxx, syy unsigned, sxy signed);
//            each thread owns 4 adjacent positions x 5 rows and slides the byte windows down the rows;
//   fix-up   (border tiles only) intermediates outside the image take the value at the CLAMPED position;
//   stage C  3x3 binomial of the three planes = six IDP.2A (dp2a) per pixel and plane on 16-bit pairs: the
//            horizontal [1 2 1] and the vertical weight ride in the coefficient bytes, 32-bit results, no unpacking;
//            then /16 and the response as before.
// ~68 instructions per pixel instead of ~119, most of them on the FMA pipe (IDP / IMAD) that version 1 left idle
// while it saturated the ALU pipe.  Results are bit-identical (same integer values at every stage).
// ================================================================================================
constexpr int H2_TIN_STRIDE = 144, H2_TIN_ROWS = 37;  // byte (r, t) <-> input (gy0 - 2 + r, gx0 - 8 + t); staged r < 36, 4 <= t < 140
constexpr int H2_PL_COLS = 136, H2_PL_ROWS = 35;      // plane (q, c) <-> intermediate position (jy, jx) = (q - 1, c - 4)
constexpr int H2_NT = 256;

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned mad_u32(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_uu(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_uu(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_ss(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_ss(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__global__ void __launch_bounds__(H2_NT, 6) harris_fused2_kernel(const __grid_constant__ HarrisParams p) {
    __shared__ __align__(16) unsigned char tin[H2_TIN_ROWS * H2_TIN_STRIDE];
    __shared__ __align__(16) unsigned short sxx[H2_PL_ROWS * H2_PL_COLS];
    __shared__ __align__(16) unsigned short syy[H2_PL_ROWS * H2_PL_COLS];
    __shared__ __align__(16) short sxy[H2_PL_ROWS * H2_PL_COLS];

    const int tid = threadIdx.x;
    const int gx0 = blockIdx.x * HTW, gy0 = blockIdx.y * HTH;

    // ---- stage A: 36 rows x 34 words of input bytes
    {
        const int x_start = p.in_ox + gx0 - 4, y_start = p.in_oy + gy0 - 2;   // image coordinates of (r = 0, t = 4)
        const bool interior = x_start >= p.win.lo_x && x_start + 136 <= p.win.hi_x && y_start >= p.win.lo_y && y_start + 36 <= p.win.hi_y;
        const bool aligned = ((reinterpret_cast<uintptr_t>(p.in) + (size_t)x_start) % 4 == 0) && (p.in_stride % 4 == 0);
        if (interior && aligned) {
            // all of a thread's loads are in flight before its first shared-memory store
            const uchar *base = p.in + (size_t)y_start * p.in_stride + x_start;
            constexpr int NV = 36 * 34, PER = (NV + H2_NT - 1) / H2_NT;
            unsigned t[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int v = tid + k * H2_NT;
                if (v < NV) {
                    const int r = v / 34, w = v - r * 34;
                    t[k] = __ldg(reinterpret_cast<const unsigned *>(base + (size_t)r * p.in_stride) + w);
                }
            }
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int v = tid + k * H2_NT;
                if (v < NV) {
                    const int r = v / 34, w = v - r * 34;
                    *reinterpret_cast<unsigned *>(tin + r * H2_TIN_STRIDE + 4 + 4 * w) = t[k];
                }
            }
        } else {
            ImgRef<uchar> im{p.in, p.in_stride, p.in_iw, p.in_ih};
            for (int e = tid; e < 36 * 136; e += H2_NT) {
                const int r = e / 136, c = e - r * 136;
                tin[r * H2_TIN_STRIDE + 4 + c] = fetch_bh(im, p.win, x_start + c, y_start + r, (uchar)0);
            }
        }
    }
    __syncthreads();

    // ---- stage B: group g = 4 adjacent plane columns 4g .. 4g+3, chunk k = plane rows 5k .. 5k+4
    if (tid < 34 * 7) {
        const int g = tid % 34, k = tid / 34;
        unsigned win[3][4];   // byte windows (a[i-1], a[i], a[i+1], a[i+2]) of the last three input rows
#pragma unroll
        for (int rr = 0; rr < 7; ++rr) {
            const unsigned *row = reinterpret_cast<const unsigned *>(tin + (5 * k + rr) * H2_TIN_STRIDE) + g;
            const unsigned w0 = row[0], w1 = row[1], w2 = row[2];
            unsigned cur[4] = {__byte_perm(w0, w1, 0x6543), w1, __byte_perm(w1, w2, 0x4321), __byte_perm(w1, w2, 0x5432)};
            if (rr >= 2) {
                const int q = 5 * k + rr - 2;   // plane row; input rows: win[0] = above, win[1] = centre, cur = below
                unsigned pxx[4], pyy[4], pxy[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int dx6 = dp4a_us(cur[i], 0x000100FF, dp4a_us(win[1][i], 0x000100FF, dp4a_us(win[0][i], 0x000100FF, 0)));
                    const int dy6 = dp4a_us(cur[i], 0x00010101, dp4a_us(win[0][i], 0x00FFFFFF, 0));
                    const int qx = dx6 / 6, qy = dy6 / 6;   // C truncating division (multiply-high by the compiler)
                    pxx[i] = (unsigned)(qx * qx);
                    pyy[i] = (unsigned)(qy * qy);
                    pxy[i] = (unsigned)(qx * qy);
                }
                const int o = q * H2_PL_COLS + 4 * g;
                // pack with IMAD (FMA pipe) rather than shift + or (ALU pipe, the busier one in this kernel)
                *reinterpret_cast<uint2 *>(sxx + o) = make_uint2(mad_u32(pxx[1], 65536u, pxx[0]), mad_u32(pxx[3], 65536u, pxx[2]));
                *reinterpret_cast<uint2 *>(syy + o) = make_uint2(mad_u32(pyy[1], 65536u, pyy[0]), mad_u32(pyy[3], 65536u, pyy[2]));
                *reinterpret_cast<uint2 *>(sxy + o) = make_uint2(__byte_perm(pxy[0], pxy[1], 0x5410), __byte_perm(pxy[2], pxy[3], 0x5410));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { win[0][i] = win[1][i]; win[1][i] = cur[i]; }
        }
    }
    __syncthreads();

    // ---- fix-up: the Gaussian stage applies CLAMP to the coordinates of the intermediate images
    {
        const int ylo = p.win.lo_y - p.in_oy, yhi = p.win.hi_y - p.in_oy;   // intermediates live on [0, w) x [ylo, yhi)
        if (gx0 == 0 || gx0 + HTW > p.w - 1 || gy0 - 1 < ylo || gy0 + HTH > yhi - 1) {
            for (int e = tid; e < 34 * 130; e += H2_NT) {
                const int jy = e / 130 - 1, jx = e - (jy + 1) * 130 - 1;
                const int X = gx0 + jx, Y = gy0 + jy;
                const int cx = min(max(X, 0), p.w - 1), cy = min(max(Y, ylo), yhi - 1);
                if (cx != X || cy != Y) {
                    const int src = (cy - gy0 + 1) * H2_PL_COLS + (cx - gx0 + 4), dst = (jy + 1) * H2_PL_COLS + (jx + 4);
                    sxx[dst] = sxx[src]; syy[dst] = syy[src]; sxy[dst] = sxy[src];
                }
            }
            __syncthreads();
        }
    }

    // ---- stage C: 4 px x 4 rows per thread; plane rows 4ty .. 4ty+5, plane columns 4tx+2 .. 4tx+9
    const int tx = tid & 31, ty = tid >> 5;
    int gxx[4][4], gyy[4][4], gxy[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) gxx[r][i] = gyy[r][i] = gxy[r][i] = 0;
#pragma unroll
    for (int iq = 0; iq < 6; ++iq) {
        const int o = (4 * ty + iq) * H2_PL_COLS + 4 * tx;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            const unsigned short *plane = pl == 0 ? sxx : pl == 1 ? syy : reinterpret_cast<const unsigned short *>(sxy);
            const unsigned wa = *reinterpret_cast<const unsigned *>(plane + o + 2);
            const uint2 wbc = *reinterpret_cast<const uint2 *>(plane + o + 4);
            const unsigned wd = *reinterpret_cast<const unsigned *>(plane + o + 8);
            const unsigned p34 = __byte_perm(wa, wbc.x, 0x5432), p56 = __byte_perm(wbc.x, wbc.y, 0x5432), p78 = __byte_perm(wbc.y, wd, 0x5432);
            const unsigned A[4] = {p34, wbc.x, p56, wbc.y};   // (v[i-1], v[i])
            const unsigned B[4] = {p56, wbc.y, p78, wd};      // (v[i+1], v[i+2])
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int dy = iq - r;
                if (dy < 0 || dy > 2) continue;
                const unsigned cw = dy == 1 ? 0x00020402u : 0x00010201u;   // bytes (w, 2w, w, 0), w = vertical weight
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (pl == 0) gxx[r][i] = dp2a_hi_uu(B[i], cw, dp2a_lo_uu(A[i], cw, gxx[r][i]));
                    else if (pl == 1) gyy[r][i] = dp2a_hi_uu(B[i], cw, dp2a_lo_uu(A[i], cw, gyy[r][i]));
                    else gxy[r][i] = dp2a_hi_ss(B[i], cw, dp2a_lo_ss(A[i], cw, gxy[r][i]));
                }
            }
        }
    }

    const int gx = gx0 + 4 * tx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int gy = gy0 + 4 * ty + r;
        if (gy >= p.h || gx >= p.w) continue;
        uchar o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = gxx[r][i] >> 4, y = gyy[r][i] >> 4;                       // non-negative: >> 4 == / 16
            const int xy = abs(gxy[r][i]) >> 4;                                     // |truncating / 16|: only xy * xy is used
            const float det = (float)(x * y - xy * xy);
            const float s = (float)(x + y);
            const float tr = __fmul_rn(__fmul_rn(p.k, s), s);
            o[i] = __fadd_rn(det, -tr) > p.threshold ? 1 : 0;
        }
        uchar *dst = p.out + (size_t)(p.out_oy + gy) * p.out_stride + p.out_ox + gx;
        if (gx + 3 < p.w && (reinterpret_cast<uintptr_t>(dst) % 4 == 0)) {
            store4(dst, o);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (gx + i < p.w) dst[i] = o[i];
        }
    }
}


// ================================================================================================
// Version 3: the same integer values at every stage as version 2, fewer instructions around them.
//   stage A  interior tiles: ONE thread requests the 160 x 36 byte box with cp.async.bulk.tensor.2d (SASS: UTMALDG) on an
//            mbarrier -- no per-thread address arithmetic, loads or shared-memory stores (version 2: ~140 of ~1400
//            instructions per thread); the co-resident CTAs compute meanwhile.  Border tiles: all threads, through CLAMP.
//   stage B  two IDP.4A per position and input row (D = a[i+1] - a[i-1], S = a[i-1] + a[i] + a[i+1]) kept for three rows:
//            dx*6 = D0 + D1 + D2 is one three-input add, dy*6 = S2 - S0 one subtraction (version 2: five IDP.4A); products
//            as packed 16-bit pairs straight from the multiplier -- the odd position's quotients are scaled by 256, so
//            x1*x1 + x0*x0 IS the packed pair; the xy plane is stored with a bias: |dx*6| + |dy*6| <= 1020 for any 3x3
//            byte window (the two sums share their corner pixels), so |qx * qy| <= 85 * 85 = 7225 and xy + 8192 is an
//            unsigned 14-bit value like xx and yy.
//   stage C  VERTICAL pass first, on the packed pairs: a + 2b + c of three plane rows is two 32-bit integer operations for
//            two pixels (every half stays below 2^16: 4 * 16129, 4 * 15417); the horizontal [1 2 1] is then two IDP.2A per
//            pixel and plane on the 16-bit column sums (two coefficient words, no byte permutes) -- 16 instructions per
//            output row and plane instead of 28.5, and no accumulators that live across rows (version 2 spilled two of
//            its 48).
// 60 instructions per pixel instead of 86; 363 -> 528 Gpx/s on the 32768^2 image.
// ================================================================================================
constexpr int H3_TIN_STRIDE = 160, H3_TIN_ROWS = 37;  // byte (r, t) <-> input (gy0 - 2 + r, gx0 - 16 + t); the TMA box is 160 x 36
constexpr int H3_PL_COLS = 136, H3_PL_ROWS = 35;      // plane (q, c) <-> intermediate position (jy, jx) = (q - 1, c - 6)
constexpr int H3_NT = 256;
constexpr int H3_XY_BIAS = 8192;
constexpr unsigned H3_BOX_BYTES = 160 * 36;

// the three stages of version 3 on one 128 x 32 tile whose input bytes are staged in `tin`; every thread of the CTA calls
// it (it synchronises); gx0 / gy0 = tile origin
__device__ __forceinline__ void h3_tile(const HarrisParams &p, const unsigned char *tin, unsigned short *sxx, unsigned short *syy, unsigned short *sxy,
                                        const int tid, const int gx0, const int gy0) {
    // ---- stage B: group g = intermediate columns jx = 4g - 2 .. 4g + 1 (plane columns 4g + 4 ..), chunk k = plane rows 5k .. 5k + 4
    if (tid < 33 * 7) {
        const int g = tid % 33, k = tid / 33;
        int D[2][4], Ssum[2][4];   // per-row sums of the two input rows above: D = a[i+1] - a[i-1], S = a[i-1] + a[i] + a[i+1]
#pragma unroll
        for (int rr = 0; rr < 7; ++rr) {
            const unsigned *row = reinterpret_cast<const unsigned *>(tin + (5 * k + rr) * H3_TIN_STRIDE) + g + 3;   // bytes t = 4g + 12 .. 4g + 19
            const unsigned w0 = row[0], w1 = row[1];
            const unsigned cur[4] = {__byte_perm(w0, w1, 0x4321), __byte_perm(w0, w1, 0x5432), __byte_perm(w0, w1, 0x6543), w1};
            int Dc[4], Sc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { Dc[i] = dp4a_us(cur[i], 0x000100FF, 0); Sc[i] = dp4a_us(cur[i], 0x00010101, 0); }
            if (rr >= 2) {
                const int q = 5 * k + rr - 2;   // plane row; rows: [0] = above, [1] = centre, c = below
                int qx[4], qy[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int dx6 = D[0][i] + D[1][i] + Dc[i];
                    const int dy6 = Sc[i] - Ssum[0][i];
                    qx[i] = dx6 / 6; qy[i] = dy6 / 6;   // C truncating division (multiply-high by the compiler)
                }
                // packed products: the odd position's factors are scaled by 256, so its product lands in the upper half
                unsigned pk[3][2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int x0 = qx[2 * h], y0 = qy[2 * h], x1 = qx[2 * h + 1] << 8, y1 = qy[2 * h + 1] << 8;
                    pk[0][h] = (unsigned)(x1 * x1 + x0 * x0);
                    pk[1][h] = (unsigned)(y1 * y1 + y0 * y0);
                    pk[2][h] = (unsigned)(x1 * y1 + (x0 * y0 + (H3_XY_BIAS * 65537)));
                }
                const int o = q * H3_PL_COLS + 4 * g + 4;
                *reinterpret_cast<uint2 *>(sxx + o) = make_uint2(pk[0][0], pk[0][1]);
                *reinterpret_cast<uint2 *>(syy + o) = make_uint2(pk[1][0], pk[1][1]);
                *reinterpret_cast<uint2 *>(sxy + o) = make_uint2(pk[2][0], pk[2][1]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { D[0][i] = D[1][i]; D[1][i] = Dc[i]; Ssum[0][i] = Ssum[1][i]; Ssum[1][i] = Sc[i]; }
        }
    }
    __syncthreads();

    // ---- fix-up: the Gaussian stage applies CLAMP to the coordinates of the intermediate images
    {
        const int ylo = p.win.lo_y - p.in_oy, yhi = p.win.hi_y - p.in_oy;   // intermediates live on [0, w) x [ylo, yhi)
        if (gx0 == 0 || gx0 + HTW > p.w - 1 || gy0 - 1 < ylo || gy0 + HTH > yhi - 1) {
            for (int e = tid; e < 34 * 130; e += H3_NT) {
                const int jy = e / 130 - 1, jx = e - (jy + 1) * 130 - 1;
                const int X = gx0 + jx, Y = gy0 + jy;
                const int cx = min(max(X, 0), p.w - 1), cy = min(max(Y, ylo), yhi - 1);
                if (cx != X || cy != Y) {
                    const int src = (cy - gy0 + 1) * H3_PL_COLS + (cx - gx0 + 6), dst = (jy + 1) * H3_PL_COLS + (jx + 6);
                    sxx[dst] = sxx[src]; syy[dst] = syy[src]; sxy[dst] = sxy[src];
                }
            }
            __syncthreads();
        }
    }

    // ---- stage C + response: 4 px x 4 rows per thread, one output row at a time over a rolling window of plane rows
    const int tx = tid & 31, ty = tid >> 5;
    const int gx = gx0 + 4 * tx, gyb = gy0 + 4 * ty;
    if (gx >= p.w || gyb >= p.h) return;
    const int nrows = p.h - gyb;
    uchar *dst = p.out + (size_t)(p.out_oy + gyb) * p.out_stride + p.out_ox + gx;
    const bool vec = gx + 3 < p.w && ((reinterpret_cast<uintptr_t>(dst) | (unsigned)p.out_stride) & 3u) == 0;
    constexpr unsigned cw = 0x00010201u;   // dp2a.lo: (1, 2) on (v[i-1], v[i]); dp2a.hi: (1, 0) on (v[i+1], v[i+2])
    constexpr unsigned ce = 0x01020100u;   // dp2a.lo: (0, 1) on (v[i-2], v[i-1]); dp2a.hi: (2, 1) on (v[i], v[i+1])
    const int o0 = 4 * ty * H3_PL_COLS + 4 * tx + 4;   // plane columns 4tx + 4 .. 4tx + 11 (8-byte aligned)
    const unsigned short *pl0 = sxx + o0, *pl1 = syy + o0, *pl2 = sxy + o0;
    uint2 ra[3][2], rb[3][2];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const unsigned short *pp = pl == 0 ? pl0 : pl == 1 ? pl1 : pl2;
        ra[pl][0] = *reinterpret_cast<const uint2 *>(pp); ra[pl][1] = *reinterpret_cast<const uint2 *>(pp + 4);
        rb[pl][0] = *reinterpret_cast<const uint2 *>(pp + H3_PL_COLS); rb[pl][1] = *reinterpret_cast<const uint2 *>(pp + H3_PL_COLS + 4);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r >= nrows) return;   // rows below the image
        {
            int G[3][4];
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
                const unsigned short *pp = (pl == 0 ? pl0 : pl == 1 ? pl1 : pl2) + (r + 2) * H3_PL_COLS;
                const uint2 c0 = *reinterpret_cast<const uint2 *>(pp), c1 = *reinterpret_cast<const uint2 *>(pp + 4);
                // column sums a + 2b + c of the three rows, two pixels per register (no carry between the halves)
                const unsigned V0 = mad_u32(rb[pl][0].x, 2u, ra[pl][0].x) + c0.x, V1 = mad_u32(rb[pl][0].y, 2u, ra[pl][0].y) + c0.y;
                const unsigned V2 = mad_u32(rb[pl][1].x, 2u, ra[pl][1].x) + c1.x, V3 = mad_u32(rb[pl][1].y, 2u, ra[pl][1].y) + c1.y;
                ra[pl][0] = rb[pl][0]; ra[pl][1] = rb[pl][1]; rb[pl][0] = c0; rb[pl][1] = c1;
                // Vk = (v[2k], v[2k+1]); pixel i needs v[i+1] + 2 v[i+2] + v[i+3].  Odd pixels start on a register boundary
                // (weights (1, 2) | (1, 0)); even pixels straddle it, which the OTHER coefficient word absorbs
                // (weights (0, 1) | (2, 1)) -- no byte permutes to realign the pairs
                const int init = pl == 2 ? -16 * H3_XY_BIAS : 0;
                G[pl][0] = dp2a_hi_uu(V1, ce, dp2a_lo_uu(V0, ce, init));
                G[pl][1] = dp2a_hi_uu(V2, cw, dp2a_lo_uu(V1, cw, init));
                G[pl][2] = dp2a_hi_uu(V2, ce, dp2a_lo_uu(V1, ce, init));
                G[pl][3] = dp2a_hi_uu(V3, cw, dp2a_lo_uu(V2, cw, init));
            }
            unsigned o = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int x = G[0][i] >> 4, y = G[1][i] >> 4;                           // non-negative: >> 4 == / 16
                const int xy = abs(G[2][i]) >> 4;                                       // |truncating / 16|: only xy * xy is used
                const float det = (float)(x * y - xy * xy);
                const float s = (float)(x + y);
                const float tr = __fmul_rn(__fmul_rn(p.k, s), s);
                if (__fadd_rn(det, -tr) > p.threshold) o |= 1u << (8 * i);
            }
            if (vec) {
                *reinterpret_cast<unsigned *>(dst) = o;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (gx + i < p.w) dst[i] = (uchar)(o >> (8 * i));
            }
            dst += p.out_stride;
        }
    }
}

// One tile per CTA: the co-resident CTAs (5 per SM) cover each other's TMA latency and barriers.  Two vertically adjacent
// tiles per CTA with both TMA requests issued up front were measured slower (476-490 vs 512 Gpx/s: more registers or spills,
// and a CTA's two tiles serialise on its plane buffers).
template <int MINB>
__global__ void __launch_bounds__(H3_NT, MINB) harris_fused3_kernel(const __grid_constant__ HarrisParams p, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(128) unsigned char tin[H3_TIN_ROWS * H3_TIN_STRIDE];
    __shared__ __align__(16) unsigned short sxx[H3_PL_ROWS * H3_PL_COLS];
    __shared__ __align__(16) unsigned short syy[H3_PL_ROWS * H3_PL_COLS];
    __shared__ __align__(16) unsigned short sxy[H3_PL_ROWS * H3_PL_COLS];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int gx0 = blockIdx.x * HTW, gy0 = blockIdx.y * HTH;

    // ---- stage A
    const int x_need = p.in_ox + gx0 - 4, y_start = p.in_oy + gy0 - 2;   // image coordinates of (r = 0, t = 12)
    const bool interior = x_need >= p.win.lo_x && x_need + 136 <= p.win.hi_x && y_start >= p.win.lo_y && y_start + 36 <= p.win.hi_y;
    if (interior) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&bar, H3_BOX_BYTES);
            tma_load_2d(tin, &tmap, x_need - 12, y_start, &bar);   // 16-byte aligned origin (launch condition: in_ox % 16 == 0)
        }
        __syncthreads();   // the initialised barrier is visible to the waiting threads
        mbar_wait(&bar, 0);
    } else {   // border tiles: all threads, through CLAMP
        ImgRef<uchar> im{p.in, p.in_stride, p.in_iw, p.in_ih};
        for (int e = tid; e < 36 * 136; e += H3_NT) {
            const int r = e / 136, c = e - r * 136;
            tin[r * H3_TIN_STRIDE + 12 + c] = fetch_bh(im, p.win, x_need + c, y_start + r, (uchar)0);
        }
        __syncthreads();
    }
    h3_tile(p, tin, sxx, syy, sxy, tid, gx0, gy0);
}

}  // namespace hb

using namespace hb;

extern "C" int hb_harris(const hb_harris_desc *d, void *stream) {
    HB_REQUIRE(d, HB_ERR_INVALID, "hb_harris: null descriptor");
    hb_view in = norm_view(d->in), out = norm_view(d->out);
    HB_REQUIRE(view_ok(in) && view_ok(out) && in.dtype == HB_U8 && out.dtype == HB_U8, HB_ERR_INVALID, "hb_harris: needs valid u8 views");
    HB_REQUIRE(in.width == out.width && in.height == out.height, HB_ERR_INVALID, "hb_harris: input and output regions must have the same size");
    // the fused Sobel + Gaussian reads two rows beyond an interior strip edge: with fewer ghost rows the row index would
    // be clamped inside the window and the result would silently differ from the unsharded pipeline
    HB_REQUIRE((in.ghost_top == 0 || in.ghost_top >= 2) && (in.ghost_bottom == 0 || in.ghost_bottom >= 2), HB_ERR_INVALID,
               "hb_harris: a sharded strip needs >= 2 ghost rows per interior side (got %d / %d)", in.ghost_top, in.ghost_bottom);
    HarrisParams p;
    memset(&p, 0, sizeof(p));
    p.in = static_cast<const uchar *>(in.data); p.out = static_cast<uchar *>(out.data);
    p.in_stride = in.stride; p.in_iw = in.img_width; p.in_ih = in.img_height;
    p.win = Window{in.offset_x, in.offset_x + in.width, in.offset_y - in.ghost_top, in.offset_y + in.height + in.ghost_bottom, HB_BOUNDARY_CLAMP};
    p.in_ox = in.offset_x; p.in_oy = in.offset_y;
    p.out_stride = out.stride; p.out_ox = out.offset_x; p.out_oy = out.offset_y; p.w = out.width; p.h = out.height;
    p.k = d->k; p.threshold = d->threshold;
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_harris");
    dim3 grid((p.w + HTW - 1) / HTW, (p.h + HTH - 1) / HTH);
    static int ver = -1;
    if (ver < 0) {
        const char *e1 = getenv("HB_HARRIS_V1"), *e = getenv("HB_HARRIS_VERSION");   // A/B knobs: the earlier versions of the fused kernel
        ver = (e1 && atoi(e1)) ? 1 : (e && atoi(e) >= 1 && atoi(e) <= 3) ? atoi(e) : 3;
    }
    // version 3 stages interior tiles with TMA: 16-byte aligned base, pitch and box origins, else version 2 (all-threads loader)
    CUtensorMap tmap;
    const bool tma_ok = ver == 3 && (p.in_ox % 16 == 0) && make_tile_map(&tmap, p.in, HB_U8, p.in_iw, p.in_ih, p.in_stride, H3_TIN_STRIDE, 36);
    if (ver == 1) harris_fused_kernel<<<grid, dim3(HBX, HBY), 0, s>>>(p);
    else if (!tma_ok) harris_fused2_kernel<<<grid, H2_NT, 0, s>>>(p);
    else {
        static int minb = -1;
        if (minb < 0) {
            const char *e = getenv("HB_HARRIS_CTAS");   // tuning knob: resident CTAs per SM the kernel is compiled for (register cap)
            minb = e ? atoi(e) : 5;
        }
        if (minb == 6) harris_fused3_kernel<6><<<grid, H3_NT, 0, s>>>(p, tmap);
        else if (minb == 4) harris_fused3_kernel<4><<<grid, H3_NT, 0, s>>>(p, tmap);
        else harris_fused3_kernel<5><<<grid, H3_NT, 0, s>>>(p, tmap);   // 48 registers, no spills: 516 Gpx/s at 32768^2 (4 CTAs: 512; 6 CTAs, 40 registers with spills: 481)
    }
    g_launches++;
    return scope.finish();
}
