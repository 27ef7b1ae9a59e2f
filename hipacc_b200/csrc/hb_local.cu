// hb_local.cu -- local-operator engine (tiled 2-D stencils) for sm_100a.
//
// Replaces the generated local-operator kernels of Hipacc's CUDA backend (kernel text
// lib/Rewrite/Rewrite.cpp:2726-2883, body lib/AST/ASTTranslate.cpp:510-1197; one pixel per
// thread, per-block `goto` border variants) -- paths relative to the Hipacc tree.
//
// Design (DESIGN.md section "Local-operator engine"):
//   * one CTA = one 128 x 32 output tile; the input tile + halo is staged ONCE into shared
//     memory, already converted to the accumulation type, so the boundary mode is a property of
//     the loader only: interior tiles take a branch-free 16-byte vector path, border tiles the
//     remapping path; the compute loop is identical and branch-free for both.
//   * each thread owns 4 adjacent pixels x 4 rows (register blocking); it walks the input rows
//     once (row-stationary): every staged row is read from shared memory with 16-byte loads and
//     feeds all output rows it contributes to.  Per output pixel taps are still folded in
//     row-major order, so results are bit-identical to the DSL's sequential fold.
//   * mask coefficients live in the kernel parameter (constant) bank and the tap loops are fully
//     unrolled for 3x3 / 5x5 / 7x7; other sizes use the generic kernel below.
#include "hb_local.cuh"

#include <cstdlib>
#include <cstring>

namespace hb {

// ------------------------------------------------------------------------------------------------
// Tiled kernel: SX x SY compile-time.  VAR 0: SUM of coef*in (no domain test, hot path).
//                                        VAR 1: run-time reduce mode / tap kind / domain holes.
// ------------------------------------------------------------------------------------------------
constexpr int TW = 128, RPT = 4, BX = 32, BY = 8, TH = BY * RPT;

// CH = interleaved channels per pixel (uchar4: 4): columns then count channel elements, horizontal taps are CH
// elements apart and every thread's 4 adjacent outputs are the 4 channels of one pixel (or 4 pixels for CH = 1).
template <typename TI, typename TS, typename TO, int SX, int SY, int VAR, int CH = 1>
__global__ void __launch_bounds__(BX *BY) local_tiled_kernel(const __grid_constant__ LocalParams p) {
    constexpr int HX = (SX / 2) * CH, HY = SY / 2;
    constexpr int HXP = round_up(HX, 4);
    constexpr int TWS = TW + 2 * HXP;
    constexpr int ROWS = TH + SY - 1;
    constexpr int WIN = 4 + 2 * HXP;
    __shared__ __align__(16) TS tile[ROWS * TWS];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * BX + tx;
    const int gx0 = blockIdx.x * TW, gy0 = blockIdx.y * TH;

    stage_tile<TI, TS, ROWS, TWS, BX * BY, CH>(tile, static_cast<const TI *>(p.in), p.in_stride, p.in_iw, p.in_ih, p.win,
                                               (TI)cval_of<TS>(p), p.in_ox + gx0 - HXP, p.in_oy + gy0 - HY, tid);
    __syncthreads();

    const int mode = VAR != 1 ? (int)HB_REDUCE_SUM : p.reduce_mode;
    TS acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[r][i] = fold_identity<TS>(mode);

    const int r0 = ty * RPT;
#pragma unroll
    for (int ir = 0; ir < RPT + SY - 1; ++ir) {
        TS w[WIN];
        const TS *row = tile + (r0 + ir) * TWS + 4 * tx;
#pragma unroll
        for (int q = 0; q < WIN / 4; ++q) {
            TS t[4];
            load4(row + 4 * q, t);
            w[4 * q] = t[0]; w[4 * q + 1] = t[1]; w[4 * q + 2] = t[2]; w[4 * q + 3] = t[3];
        }
        if (VAR == 0 && sizeof(TS) == 4 && DtypeOf<TS>::v == HB_F32 && SX <= 7 && SY <= 7) {
            // float SUM of products with packed multiplies; per accumulator the products are still added one by one in
            // tap order (row-major), so the result is bit-identical to the scalar form
            float(&wf)[WIN] = reinterpret_cast<float(&)[WIN]>(w);
            float(&af)[RPT][4] = reinterpret_cast<float(&)[RPT][4]>(acc);
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int dy = ir - r;
                if (dy < 0 || dy >= SY) continue;
                if (CH == 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        constexpr int base = HXP - HX;
                        const int odd = (base + i) & 1;   // compile-time after unrolling: first tap in an odd register
                        // the first visited tap initialises the accumulator (dsl/kernel.hpp:250), it is not added to zero
                        if (odd) {
                            const float p0 = __fmul_rn(p.coef.f[dy * SX], wf[base + i]);
                            af[r][i] = dy == 0 ? p0 : __fadd_rn(af[r][i], p0);
                        }
#pragma unroll
                        for (int dx = odd; dx + 1 < SX; dx += 2) {
                            float p0, p1;
                            mul2_rn(wf[base + i + dx], wf[base + i + dx + 1], &p.cpair[odd][dy][dx - odd], p0, p1);
                            af[r][i] = __fadd_rn(dy == 0 && dx == 0 ? p0 : __fadd_rn(af[r][i], p0), p1);
                        }
                        if (((SX - odd) & 1) != 0) af[r][i] = __fadd_rn(af[r][i], __fmul_rn(p.coef.f[dy * SX + SX - 1], wf[base + i + SX - 1]));
                    }
                } else {
#pragma unroll
                    for (int dx = 0; dx < SX; ++dx)
#pragma unroll
                        for (int i = 0; i < 4; i += 2) {   // two adjacent channel elements share the coefficient
                            float p0, p1;
                            mul2_rn(wf[HXP - HX + i + dx * CH], wf[HXP - HX + i + 1 + dx * CH], p.cdup[dy * SX + dx], p0, p1);
                            af[r][i] = dy == 0 && dx == 0 ? p0 : __fadd_rn(af[r][i], p0);
                            af[r][i + 1] = dy == 0 && dx == 0 ? p1 : __fadd_rn(af[r][i + 1], p1);
                        }
                }
            }
            continue;
        }
        if (VAR == 2) {
            // separable integer mask m = v * h^T (p.coef.i[0..SX) = h, [SX..SX+SY) = v): horizontal pass once per staged
            // row, vertical weights per output row.  Integer sums are exact in any order, so this equals the 2-D fold.
            TS h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                h[i] = 0;
#pragma unroll
                for (int dx = 0; dx < SX; ++dx) h[i] = add_rn(h[i], mul_rn(coef_of<TS>(p, dx), w[HXP - HX + i + dx * CH]));
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int dy = ir - r;
                if (dy < 0 || dy >= SY) continue;
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[r][i] = add_rn(acc[r][i], mul_rn(coef_of<TS>(p, SX + dy), h[i]));
            }
            continue;
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int dy = ir - r;
            if (dy < 0 || dy >= SY) continue;
#pragma unroll
            for (int dx = 0; dx < SX; ++dx) {
                const int k = dy * SX + dx;
                if (VAR == 1 && !((p.dom[k >> 5] >> (k & 31)) & 1u)) continue;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const TS pix = w[HXP - HX + i + dx * CH];
                    TS v;
                    if (VAR == 0 || p.tap == HB_TAP_MUL) v = mul_rn(coef_of<TS>(p, k), pix);
                    else v = pix;
                    // first visited tap initialises (dsl/kernel.hpp:250,279): tap 0 when every tap is visited
                    acc[r][i] = (VAR == 0 ? k == 0 : k == p.first_tap) ? v : fold<TS>(acc[r][i], v, mode);
                }
            }
        }
    }

    TO *out = static_cast<TO *>(p.out);
    const int gx = gx0 + 4 * tx;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int gy = gy0 + r0 + r;
        if (gy >= p.is_h || gx >= p.is_w) continue;
        TO o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = epilogue<TO>(acc[r][i], p);
        TO *dst = out + (size_t)(p.out_oy + gy) * p.out_stride + p.out_ox + gx;
        if (gx + 3 < p.is_w && (reinterpret_cast<uintptr_t>(dst) % (4 * sizeof(TO)) == 0)) {
            store4(dst, o);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (gx + i < p.is_w) dst[i] = o[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic kernel: any mask size up to 13x13 (run-time), one pixel per thread, 32 x 8 tile.
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TS, typename TO>
__global__ void __launch_bounds__(256) local_generic_kernel(const __grid_constant__ LocalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TS *tile = reinterpret_cast<TS *>(smem_raw);
    const int sx = p.size_x, sy = p.size_y, hx = sx / 2, hy = sy / 2;
    const int pw = 32 + sx - 1, ph = 8 + sy - 1;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int gx0 = blockIdx.x * 32, gy0 = blockIdx.y * 8;
    {
        const TI *in = static_cast<const TI *>(p.in);
        ImgRef<TI> im{in, p.in_stride, p.in_iw, p.in_ih};
        const TI cv = (TI)cval_of<TS>(p);
        const int x_start = p.in_ox + gx0 - hx, y_start = p.in_oy + gy0 - hy;
        for (int e = tid; e < pw * ph; e += 256) {
            const int r = e / pw, c = e - r * pw;
            tile[e] = (TS)fetch_bh(im, p.win, x_start + c, y_start + r, cv);
        }
    }
    __syncthreads();
    const int gx = gx0 + tx, gy = gy0 + ty;
    if (gx >= p.is_w || gy >= p.is_h) return;
    const int mode = p.reduce_mode;
    TS acc = fold_identity<TS>(mode);
    for (int dy = 0; dy < sy; ++dy)
        for (int dx = 0; dx < sx; ++dx) {
            const int k = dy * sx + dx;
            if (!((p.dom[k >> 5] >> (k & 31)) & 1u)) continue;
            const TS pix = tile[(ty + dy) * pw + tx + dx];
            const TS v = p.tap == HB_TAP_MUL ? mul_rn(coef_of<TS>(p, k), pix) : pix;
            acc = k == p.first_tap ? v : fold<TS>(acc, v, mode);
        }
    static_cast<TO *>(p.out)[(size_t)(p.out_oy + gy) * p.out_stride + p.out_ox + gx] = epilogue<TO>(acc, p);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// 4-channel images (uchar4, float4, ...): the tiled kernel on the 4x wider channel-element image (p.* already in element
// units, p.win in pixels)
template <typename TI, typename TS, typename TO>
static int launch_local_x4(const LocalParams &p, bool fast, cudaStream_t s) {
    dim3 block(BX, BY);
    dim3 grid((p.is_w + TW - 1) / TW, (p.is_h + TH - 1) / TH);
#define HB_TILED4(SXV, SYV)                                                                        \
    if (p.size_x == SXV && p.size_y == SYV) {                                                      \
        if (fast) local_tiled_kernel<TI, TS, TO, SXV, SYV, 0, 4><<<grid, block, 0, s>>>(p);        \
        else local_tiled_kernel<TI, TS, TO, SXV, SYV, 1, 4><<<grid, block, 0, s>>>(p);             \
        g_launches++;                                                                              \
        return HB_OK;                                                                              \
    }
    HB_TILED4(3, 3)
    HB_TILED4(5, 5)
    HB_TILED4(7, 7)
#undef HB_TILED4
    return HB_ERR_UNSUPPORTED;
}

// register-blocked two-pass variant for separable INTEGER masks (VAR 2); p.coef.i holds h then v
template <typename TI, typename TO>
static int launch_local_separable(const LocalParams &p, cudaStream_t s) {
    dim3 block(BX, BY);
    dim3 grid((p.is_w + TW - 1) / TW, (p.is_h + TH - 1) / TH);
#define HB_SEP(SXV, SYV)                                                          \
    if (p.size_x == SXV && p.size_y == SYV) {                                     \
        local_tiled_kernel<TI, int, TO, SXV, SYV, 2><<<grid, block, 0, s>>>(p);   \
        g_launches++;                                                             \
        return HB_OK;                                                             \
    }
    HB_SEP(3, 3)
    HB_SEP(5, 5)
    HB_SEP(7, 7)
#undef HB_SEP
    return HB_ERR_UNSUPPORTED;
}

// m[j][i] == v[j] * h[i] with integer h, v?  (rank-1 factorisation through the first non-zero entry)
static bool factor_separable(const int *m, int sx, int sy, int *h, int *v) {
    int r0 = -1, c0 = -1;
    for (int k = 0; k < sx * sy && r0 < 0; ++k)
        if (m[k]) { r0 = k / sx; c0 = k % sx; }
    if (r0 < 0) return false;
    auto gcd = [](long long a, long long b) { a = a < 0 ? -a : a; b = b < 0 ? -b : b; while (b) { long long t = a % b; a = b; b = t; } return a; };
    long long g = 0;
    for (int i = 0; i < sx; ++i) g = gcd(g, m[r0 * sx + i]);
    for (int i = 0; i < sx; ++i) h[i] = (int)(m[r0 * sx + i] / g);
    for (int j = 0; j < sy; ++j) {
        if (m[j * sx + c0] % h[c0]) return false;
        v[j] = m[j * sx + c0] / h[c0];
    }
    for (int j = 0; j < sy; ++j)
        for (int i = 0; i < sx; ++i)
            if ((long long)v[j] * h[i] != m[j * sx + i]) return false;
    return true;
}

template <typename TI, typename TS, typename TO>
static int launch_local(const LocalParams &p, bool fast, cudaStream_t s) {
    dim3 block(BX, BY);
    dim3 grid((p.is_w + TW - 1) / TW, (p.is_h + TH - 1) / TH);
#define HB_TILED(SXV, SYV)                                                                   \
    if (p.size_x == SXV && p.size_y == SYV) {                                                \
        if (fast) local_tiled_kernel<TI, TS, TO, SXV, SYV, 0><<<grid, block, 0, s>>>(p);     \
        else local_tiled_kernel<TI, TS, TO, SXV, SYV, 1><<<grid, block, 0, s>>>(p);          \
        g_launches++;                                                                        \
        return HB_OK;                                                                        \
    }
    HB_TILED(3, 3)
    HB_TILED(5, 5)
    HB_TILED(7, 7)
#undef HB_TILED
    dim3 g2((p.is_w + 31) / 32, (p.is_h + 7) / 8);
    size_t smem = (size_t)(32 + p.size_x - 1) * (8 + p.size_y - 1) * sizeof(TS);
    local_generic_kernel<TI, TS, TO><<<g2, dim3(32, 8), smem, s>>>(p);
    g_launches++;
    return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" int hb_local_op(const hb_local_desc *d, void *stream) {
    HB_REQUIRE(d, HB_ERR_INVALID, "hb_local_op: null descriptor");
    hb_view in = norm_view(d->in), out = norm_view(d->out);
    HB_REQUIRE(view_ok(in) && view_ok(out), HB_ERR_INVALID, "hb_local_op: malformed view");
    const bool x4 = is_x4(in.dtype);
    HB_REQUIRE(x4 == is_x4(out.dtype), HB_ERR_UNSUPPORTED, "hb_local_op: a 4-channel input needs a 4-channel output (and vice versa)");
    const int win_lo_x = in.offset_x, win_hi_x = in.offset_x + in.width;   // boundary window in PIXELS
    if (x4) { in = as_channels(in); out = as_channels(out); }              // everything else in channel elements
    HB_REQUIRE(d->size_x > 0 && d->size_y > 0 && (d->size_x & 1) && (d->size_y & 1) && d->size_x * d->size_y <= kMaxTaps &&
                   d->size_x <= 13 && d->size_y <= 13,
               HB_ERR_UNSUPPORTED, "hb_local_op: mask %dx%d unsupported (odd sizes up to 13x13)", d->size_x, d->size_y);
    HB_REQUIRE(d->tap == HB_TAP_IN || d->coef_f32 || d->coef_s32, HB_ERR_INVALID, "hb_local_op: HB_TAP_MUL needs coefficients");
    HB_REQUIRE(d->reduce_mode >= HB_REDUCE_SUM && d->reduce_mode <= HB_REDUCE_PROD, HB_ERR_INVALID, "hb_local_op: bad reduce mode");
    HB_REQUIRE(d->boundary >= HB_BOUNDARY_UNDEFINED && d->boundary <= HB_BOUNDARY_CONSTANT, HB_ERR_INVALID, "hb_local_op: bad boundary mode");

    LocalParams p;
    memset(&p, 0, sizeof(p));
    p.in = in.data; p.out = out.data;
    p.in_stride = in.stride; p.in_iw = in.img_width; p.in_ih = in.img_height;
    p.win = Window{win_lo_x, win_hi_x, in.offset_y - in.ghost_top, in.offset_y + in.height + in.ghost_bottom, d->boundary};
    p.in_ox = in.offset_x; p.in_oy = in.offset_y;
    p.out_stride = out.stride; p.out_ox = out.offset_x; p.out_oy = out.offset_y; p.is_w = out.width; p.is_h = out.height;
    p.size_x = d->size_x; p.size_y = d->size_y;
    p.reduce_mode = d->reduce_mode; p.tap = d->tap; p.acc_s16 = d->acc_dtype == HB_S16; p.epilogue = d->epilogue;
    for (int i = 0; i < 3; ++i) { p.epi_f[i] = (float)d->epi_p[i]; p.epi_i[i] = (int)d->epi_p[i]; }
    HB_REQUIRE(!(d->epilogue == HB_EPI_DIVI_CAST && p.epi_i[0] == 0), HB_ERR_INVALID, "hb_local_op: division by zero in epilogue");
    p.cval_f = (float)d->boundary_const; p.cval_i = (int)d->boundary_const;

    const bool facc = d->acc_dtype == HB_F32;
    HB_REQUIRE(facc || d->acc_dtype == HB_S32 || d->acc_dtype == HB_S16, HB_ERR_UNSUPPORTED, "hb_local_op: accumulator type %d unsupported", d->acc_dtype);
    HB_REQUIRE(!(d->tap == HB_TAP_MUL && !facc && !d->coef_s32), HB_ERR_UNSUPPORTED,
               "hb_local_op: float mask with an integer accumulator is not supported");
    const int n = d->size_x * d->size_y;
    int visited = 0;
    p.first_tap = -1;
    for (int k = 0; k < n; ++k) {
        bool on = true;
        if (d->kind == HB_LOCAL_REDUCE_DOMAIN) {
            if (d->domain) on = d->domain[k] != 0;
            else if (d->tap == HB_TAP_MUL) on = d->coef_f32 ? d->coef_f32[k] != 0.0f : d->coef_s32[k] != 0;
        }
        if (on) { p.dom[k >> 5] |= 1u << (k & 31); ++visited; if (p.first_tap < 0) p.first_tap = k; }
        if (d->tap == HB_TAP_MUL) {
            // holes contribute coef 0 in the SUM fast path (exact: x + 0 == x)
            if (facc) p.coef.f[k] = on ? (d->coef_f32 ? d->coef_f32[k] : (float)d->coef_s32[k]) : 0.0f;
            else p.coef.i[k] = on ? d->coef_s32[k] : 0;
        }
    }
    HB_REQUIRE(visited > 0, HB_ERR_INVALID, "hb_local_op: empty domain");
    if (facc && d->tap == HB_TAP_MUL && d->size_x <= 7 && d->size_y <= 7) {   // operands of the packed multiplies
        for (int dy = 0; dy < d->size_y; ++dy)
            for (int dx = 0; dx < 8; ++dx) {
                p.cpair[0][dy][dx] = dx < d->size_x ? p.coef.f[dy * d->size_x + dx] : 0.0f;
                p.cpair[1][dy][dx] = dx + 1 < d->size_x ? p.coef.f[dy * d->size_x + dx + 1] : 0.0f;
            }
        for (int k = 0; k < n; ++k) p.cdup[k][0] = p.cdup[k][1] = p.coef.f[k];
    }
    // SUM-of-products fast variant: Domain holes ride along as zero coefficients.  That is exact for integer pixels
    // (0 * pixel adds nothing); for FLOAT pixels a hole over an inf / NaN pixel would poison the sum although the DSL
    // never visits that tap, so float images with holes take the domain-testing variant.
    const bool sum_of_products = d->reduce_mode == HB_REDUCE_SUM && d->tap == HB_TAP_MUL;
    const bool fast = sum_of_products && (visited == n || in.dtype != HB_F32);

    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_local_op");
    int rc = HB_ERR_UNSUPPORTED;
    const int it = in.dtype, ot = out.dtype;
    static int no_pair = -1;
    if (no_pair < 0) { const char *e = getenv("HB_NO_PAIR"); no_pair = (e && atoi(e)) ? 1 : 0; }
    const bool pair_ok = facc && fast && !no_pair;   // float SUM of products over every tap: two pixels per FMUL2 / FADD2
    if (x4) {
        // element-wise per channel (dsl/types.hpp): uchar4 -> uchar4 (Gaussian / Laplace / Dilate / Box _RGBA), uchar4 -> int4 /
        // short4 (Sobel_RGBA's intermediates), float4 -> float4
        if (it == HB_U8 && ot == HB_U8) {
            if (pair_ok) rc = launch_local_pair(p, HB_U8, HB_U8, 4, s);
            if (rc == HB_ERR_UNSUPPORTED) rc = facc ? launch_local_x4<uchar, float, uchar>(p, fast, s) : launch_local_x4<uchar, int, uchar>(p, fast, s);
        } else if (it == HB_U8 && ot == HB_S32 && !facc) rc = launch_local_x4<uchar, int, int>(p, fast, s);
        else if (it == HB_U8 && ot == HB_S16 && !facc) rc = launch_local_x4<uchar, int, short>(p, fast, s);
        else if (it == HB_F32 && ot == HB_F32 && facc) rc = launch_local_x4<float, float, float>(p, fast, s);
        HB_REQUIRE(rc != HB_ERR_UNSUPPORTED, HB_ERR_UNSUPPORTED,
                   "hb_local_op: 4-channel images support 3x3, 5x5 and 7x7 masks on uchar4 -> uchar4 / short4 / int4 and float4 -> float4; no CPU fallback");
    } else if (facc) {
        if (it == HB_U8 && ot == HB_U8) {
            if (pair_ok) rc = launch_local_pair(p, it, ot, 1, s);
            if (rc == HB_ERR_UNSUPPORTED) rc = launch_local<uchar, float, uchar>(p, fast, s);
        } else if (it == HB_F32 && ot == HB_F32) {
            // hot path: persistent TMA-pipelined kernel (hb_local_tma.cu); anything it does not take runs staged
            // (the TMA kernel skips Domain holes itself: constexpr masks drop them, run-time masks are only taken when
            // every tap is visited)
            if (sum_of_products && d->epilogue == HB_EPI_CAST) rc = launch_local_tma_f32(p, visited == n, s);
            if (rc == HB_ERR_UNSUPPORTED && pair_ok) rc = launch_local_pair(p, it, ot, 1, s);
            if (rc == HB_ERR_UNSUPPORTED) rc = launch_local<float, float, float>(p, fast, s);
        } else if (it == HB_S8 && ot == HB_S8) {
            if (pair_ok) rc = launch_local_pair(p, it, ot, 1, s);
            if (rc == HB_ERR_UNSUPPORTED) rc = launch_local<signed char, float, signed char>(p, fast, s);
        }
    } else {
        // separable integer masks (Sobel 3x3 / 5x5 / 7x7, binomial Gaussians): two-pass variant, SX + SY instead of
        // SX * SY multiply-adds per pixel; exact because integer sums do not depend on the order
        static int no_sep = -1;
        if (no_sep < 0) { const char *e = getenv("HB_NO_SEPARABLE"); no_sep = (e && atoi(e)) ? 1 : 0; }
        int hv[2 * 13];
        if (fast && !no_sep && p.size_x == p.size_y && p.size_x <= 7 && p.size_x >= 3 &&
            factor_separable(p.coef.i, p.size_x, p.size_y, hv, hv + p.size_x)) {
            LocalParams q = p;
            for (int k = 0; k < p.size_x + p.size_y; ++k) q.coef.i[k] = hv[k];
            if (it == HB_U8 && ot == HB_U8) rc = launch_local_separable<uchar, uchar>(q, s);
            else if (it == HB_U8 && ot == HB_S32) rc = launch_local_separable<uchar, int>(q, s);
            else if (it == HB_U8 && ot == HB_S16) rc = launch_local_separable<uchar, short>(q, s);
            else if (it == HB_S16 && ot == HB_S16) rc = launch_local_separable<short, short>(q, s);
        }
        if (rc == HB_ERR_UNSUPPORTED) {
            if (it == HB_U8 && ot == HB_U8) rc = launch_local<uchar, int, uchar>(p, fast, s);
            else if (it == HB_U8 && ot == HB_S32) rc = launch_local<uchar, int, int>(p, fast, s);
            else if (it == HB_U8 && ot == HB_S16) rc = launch_local<uchar, int, short>(p, fast, s);
            else if (it == HB_S16 && ot == HB_S16) rc = launch_local<short, int, short>(p, fast, s);
        }
    }
    HB_REQUIRE(rc != HB_ERR_UNSUPPORTED, HB_ERR_UNSUPPORTED,
               "hb_local_op: no device kernel for (in %d, acc %d, out %d); there is no CPU fallback", it, d->acc_dtype, ot);
    return scope.finish();
}
