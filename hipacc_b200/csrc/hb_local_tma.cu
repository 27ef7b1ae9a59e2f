// hb_local_tma.cu -- TMA-pipelined local-operator kernel (float -> float, SUM of coef * in) for sm_100a.
//
// This is the hot kernel of BASELINE config C2 (Sobel / Laplace 3x3 float, MIRROR) and of the float
// Gaussians of the pyramid (C5).  It replaces the generated Hipacc CUDA kernel for such operators
// (lib/Rewrite/Rewrite.cpp:2726-2883, lib/AST/ASTTranslate.cpp:510-1197).
//
//   * one CTA per 128 x TH output tile, handed out by the hardware block scheduler (a persistent grid with a multi-
//     stage ring was measured 22 % slower: its CTAs march over memory in lockstep, see launch_variant);
//   * the input box (tile + halo) of an INTERIOR tile is fetched by one thread with a single
//     cp.async.bulk.tensor.2d (TMA) into shared memory and completes on an mbarrier, so no thread spends
//     instructions on staging; the copy overlaps the arithmetic of the co-resident CTAs;
//     BORDER tiles (box crosses the accessor's boundary window) get their out-of-window cells patched by all threads
//     through the boundary-mode index remap (lib/AST/BorderHandling.cpp:41-120) -- same compute code afterwards;
//   * each thread owns 4 adjacent pixels x RPT rows and walks the staged rows once (row-stationary), taps
//     are folded per pixel in row-major order with separately rounded multiply and add, first visited tap
//     initialises (dsl/kernel.hpp:241-296): results are bit-identical to the DSL's sequential fold;
//   * masks the reference's samples define at compile time (Sobel, Laplace; Hipacc bakes constant masks into
//     the kernel text, lib/Rewrite/Rewrite.cpp:2617-2665) have constexpr-coefficient instantiations: zero
//     taps vanish and +-1 taps need no multiply.  Any other mask runs with coefficients from the parameter
//     (constant) bank.
#include "hb_local.cuh"
#include "hb_tma.cuh"

#include <cstdlib>
#include <cstring>

namespace hb {

namespace {

constexpr int TW = 128;  // tile width: 32 lanes x 4 pixels

struct MaskRuntime {
    static constexpr bool kRuntime = true;
    __host__ __device__ static constexpr float coef(int) { return 0.0f; }
};
#define HB_CONST_MASK(NAME, N, ...)                                   \
    struct NAME {                                                     \
        static constexpr bool kRuntime = false;                       \
        static constexpr int kTaps = N;                               \
        __host__ __device__ static constexpr float coef(int k) {      \
            constexpr float c[N] = {__VA_ARGS__};                     \
            return c[k];                                              \
        }                                                             \
    };
// samples-public/3_Preprocessing/Sobel/src/main.cpp:128-140, 1_Local_Operators/Laplace/src/main.cpp:106-124
HB_CONST_MASK(MaskSobel3X, 9, -1, 0, 1, -2, 0, 2, -1, 0, 1)
HB_CONST_MASK(MaskSobel3Y, 9, -1, -2, -1, 0, 0, 0, 1, 2, 1)
HB_CONST_MASK(MaskLaplace3D, 9, 2, 0, 2, 0, -8, 0, 2, 0, 2)
HB_CONST_MASK(MaskLaplace3N, 9, 0, 1, 0, 1, -4, 1, 0, 1, 0)
HB_CONST_MASK(MaskIdentity3, 9, 0, 0, 0, 0, 1, 0, 0, 0, 0)   // data-movement ceiling of the pipeline (tools/repro_local.py)
HB_CONST_MASK(MaskSobel5X, 25, -1, -2, 0, 2, 1, -4, -8, 0, 8, 4, -6, -12, 0, 12, 6, -4, -8, 0, 8, 4, -1, -2, 0, 2, 1)
HB_CONST_MASK(MaskSobel5Y, 25, -1, -4, -6, -4, -1, -2, -8, -12, -8, -2, 0, 0, 0, 0, 0, 2, 8, 12, 8, 2, 1, 4, 6, 4, 1)
HB_CONST_MASK(MaskLaplace5, 25, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -24, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1)

template <int SX, int SY, int RPT, int BY, int NST>
struct Geo {
    static constexpr int HX = SX / 2, HY = SY / 2;
    static constexpr int TH = BY * RPT;
    // TMA needs the box origin on a 16-byte boundary (measured: an x coordinate with x * 4 % 16 != 0 raises
    // "illegal instruction"), so the staged box starts HXP = round_up(HX, 4) columns left of the tile
    static constexpr int HXP = round_up(HX, 4);
    static constexpr int TWS = TW + 2 * HXP;               // staged columns
    static constexpr int C0 = HXP - HX;                    // first column a tile really needs ...
    static constexpr int C1 = HXP + TW + HX;               // ... and one past the last
    static constexpr int ROWS = TH + SY - 1;
    static constexpr unsigned TILE_BYTES = ROWS * TWS * sizeof(float);
    static constexpr unsigned STAGE_BYTES = (TILE_BYTES + 127u) / 128u * 128u;
    static constexpr unsigned SMEM_BYTES = NST * STAGE_BYTES + 128;  // + slack to align the base to 128 bytes
};

// One CTA: NST shared-memory stages, tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the box of tile j + NST - 1
// is requested before tile j is computed.
template <int SX, int SY, typename MASK, int RPT, int BY, int NST>
__global__ void __launch_bounds__(32 * BY, NST == 1 ? 8 : 1) local_tma_f32_kernel(const __grid_constant__ LocalParams p,
                                                                const __grid_constant__ CUtensorMap tmap, const int ntx,
                                                                const int ntiles, const int stream_stores) {
    typedef Geo<SX, SY, RPT, BY, NST> G;
    constexpr int HX = G::HX, HY = G::HY, HXP = G::HXP, TH = G::TH, TWS = G::TWS, ROWS = G::ROWS;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar[NST];
    unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);  // 128-byte aligned TMA destination

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const float *in = static_cast<const float *>(p.in);
    float *out = static_cast<float *>(p.out);
    const bool vec_store = (p.out_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) + (size_t)p.out_ox * 4) % 16 == 0);

    auto origin = [&](int t, int &gx0, int &gy0, int &xs, int &ys) {
        const int by = t / ntx, bx = t - by * ntx;
        gx0 = bx * TW; gy0 = by * TH;
        xs = p.in_ox + gx0 - HXP; ys = p.in_oy + gy0 - HY;   // box origin (16-byte aligned: in_ox % 4 == 0 is a launch condition)
    };
    auto request = [&](int t, int stage) {   // thread 0 only
        int gx0, gy0, xs, ys;
        origin(t, gx0, gy0, xs, ys);
        mbar_arrive_expect_tx(&bar[stage], G::TILE_BYTES);
        tma_load_2d(smem + stage * G::STAGE_BYTES, &tmap, xs, ys, &bar[stage]);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST - 1; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < ntiles) request((int)t, s);
        }
    }
    unsigned phases = 0;
    int stage = 0;
    for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const int t = (int)tl;
        {   // refill the stage released by the barrier that ended the previous iteration
            const long long tn = tl + (long long)(NST - 1) * gridDim.x;
            if (tid == 0 && tn < ntiles) request((int)tn, stage == 0 ? NST - 1 : stage - 1);
        }
        int gx0, gy0, xs, ys;
        origin(t, gx0, gy0, xs, ys);
        float *tile = reinterpret_cast<float *>(smem + stage * G::STAGE_BYTES);
        mbar_wait(&bar[stage], (phases >> stage) & 1u);
        phases ^= 1u << stage;

        // Border tiles: the box crosses the accessor's boundary window.  TMA delivered the in-window part (and
        // zeros / neighbouring data elsewhere); overwrite just the out-of-window cells that valid outputs read
        // with the boundary-mode value (lib/AST/BorderHandling.cpp:41-120).  Interior tiles skip all of this.
        const int c_end = min(G::C1, p.is_w - gx0 + HXP + HX), r_end = min(ROWS, p.is_h - gy0 + 2 * HY);
        const int cl = p.win.lo_x - xs, cr = p.win.hi_x - xs, rt = p.win.lo_y - ys, rb = p.win.hi_y - ys;  // window in tile coordinates
        if (cl > G::C0 || cr < c_end || rt > 0 || rb < r_end) {
            ImgRef<float> im{in, p.in_stride, p.in_iw, p.in_ih};
            const int nl = max(min(cl, c_end) - G::C0, 0), nr = max(c_end - max(cr, G::C0), 0);   // out-of-window columns left / right
            const int ncol = nl + nr;
            for (int e = tid; e < ncol * r_end; e += 32 * BY) {       // column strips, all rows
                const int r = e / ncol, k = e - r * ncol;
                const int c = k < nl ? G::C0 + k : max(cr, G::C0) + (k - nl);
                tile[r * TWS + c] = fetch_bh(im, p.win, xs + c, ys + r, p.cval_f);
            }
            const int nt = max(min(rt, r_end), 0), nb = max(r_end - max(rb, 0), 0);     // out-of-window rows top / bottom
            const int wcol = c_end - G::C0;
            for (int e = tid; e < (nt + nb) * wcol; e += 32 * BY) {   // row strips, all columns
                const int k = e / wcol, c = G::C0 + (e - k * wcol);
                const int r = k < nt ? k : max(rb, 0) + (k - nt);
                tile[r * TWS + c] = fetch_bh(im, p.win, xs + c, ys + r, p.cval_f);
            }
            fence_proxy_async();  // generic-proxy writes are ordered before the next TMA write into this stage
            __syncthreads();
        }

        float acc[RPT][4];
        const int r0 = ty * RPT;
#pragma unroll
        for (int ir = 0; ir < RPT + SY - 1; ++ir) {
            // window of 4 + 2*HX pixels: the 4 centre pixels with one 16-byte load, the halos with the
            // narrowest aligned loads that cover them
            float w[4 + 2 * HX];
            const float *row = tile + (r0 + ir) * TWS + 4 * tx + HXP;  // centre pixel 0
            {
                const float4 v = *reinterpret_cast<const float4 *>(row);
                w[HX] = v.x; w[HX + 1] = v.y; w[HX + 2] = v.z; w[HX + 3] = v.w;
            }
            if (HX == 1) {
                w[0] = row[-1];
                w[5] = row[4];
            } else if (HX == 2) {
                const float2 a = *reinterpret_cast<const float2 *>(row - 2), b = *reinterpret_cast<const float2 *>(row + 4);
                w[0] = a.x; w[1] = a.y; w[6] = b.x; w[7] = b.y;
            } else {
                const float4 a = *reinterpret_cast<const float4 *>(row - 4), b = *reinterpret_cast<const float4 *>(row + 4);
                w[0] = a.y; w[1] = a.z; w[2] = a.w; w[7] = b.x; w[8] = b.y; w[9] = b.z;
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int dy = ir - r;
                if (dy < 0 || dy >= SY) continue;
#pragma unroll
                for (int dx = 0; dx < SX; ++dx) {
                    const int k = dy * SX + dx;
                    if (MASK::kRuntime) {
                        const float c = p.coef.f[k];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float v = __fmul_rn(c, w[i + dx]);
                            acc[r][i] = k == 0 ? v : __fadd_rn(acc[r][i], v);
                        }
                    } else {
                        const float c = MASK::coef(k);
                        if (c == 0.0f) continue;  // Domain hole: not visited (dsl/mask.hpp:112-126)
                        bool first = true;        // is k the first non-zero tap?
                        for (int j = 0; j < k; ++j) first = first && (MASK::coef(j) == 0.0f);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float pix = w[i + dx];
                            const float v = c == 1.0f ? pix : c == -1.0f ? -pix : __fmul_rn(c, pix);
                            acc[r][i] = first ? v : __fadd_rn(acc[r][i], v);
                        }
                    }
                }
            }
        }

        const int gx = gx0 + 4 * tx;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int gy = gy0 + r0 + r;
            if (gy >= p.is_h || gx >= p.is_w) continue;
            float *dst = out + (size_t)(p.out_oy + gy) * p.out_stride + p.out_ox + gx;
            if (vec_store && gx + 3 < p.is_w) {
                const float4 o = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                if (stream_stores) __stcs(reinterpret_cast<float4 *>(dst), o);   // written once, not re-read by this kernel
                else *reinterpret_cast<float4 *>(dst) = o;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (gx + i < p.is_w) dst[i] = acc[r][i];
            }
        }
        __syncthreads();  // every thread is done reading this stage: it may be refilled by the next request
        stage = stage + 1 == NST ? 0 : stage + 1;
    }
}

template <int SX, int SY, typename MASK, int RPT, int BY, int NST>
int launch_variant(const LocalParams &p, cudaStream_t s) {
    typedef Geo<SX, SY, RPT, BY, NST> G;
    auto kern = local_tma_f32_kernel<SX, SY, MASK, RPT, BY, NST>;
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES), "cudaFuncSetAttribute()")) return HB_ERR_CUDA;
        int n = 0;
        if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 32 * BY, G::SMEM_BYTES), "cudaOccupancyMaxActiveBlocksPerMultiprocessor()")) return HB_ERR_CUDA;
        ctas_per_sm = n > 0 ? n : 1;
        if (const char *e = getenv("HB_TMA_CTAS")) {  // tuning knob: cap the resident CTAs per SM
            const int cap = atoi(e);
            if (cap > 0 && cap < ctas_per_sm) ctas_per_sm = cap;
        }
    }
    CUtensorMap tmap;
    if (!make_tile_map(&tmap, p.in, HB_F32, p.in_iw, p.in_ih, p.in_stride, G::TWS, G::ROWS)) return HB_ERR_UNSUPPORTED;
    const int ntx = (p.is_w + TW - 1) / TW, nty = (p.is_h + G::TH - 1) / G::TH;
    const long long ntiles = (long long)ntx * nty;
    if (ntiles > 0x7fffffffLL) return HB_ERR_UNSUPPORTED;
    // One tile per CTA by default: measured on B200 (profiles/r1p_tma_grid_sweep.txt), a grid of ntiles CTAs that the
    // block scheduler hands out as CTAs retire streams at the HBM copy peak (809 Gpx/s), while a persistent grid of
    // resident CTAs marching over the tiles in lockstep reaches 636 Gpx/s with the same kernel -- reads and writes of
    // all SMs fall into phase.  HB_TMA_GRID_MULT = m > 0 restores a capped grid of m x (resident CTAs) for comparison;
    // the kernel's tile loop and mbarrier ring handle either.
    static int grid_mult = -1;
    if (grid_mult < 0) {
        const char *e = getenv("HB_TMA_GRID_MULT");
        grid_mult = e ? atoi(e) : 0;
        if (grid_mult < 0) grid_mult = 0;
    }
    long long grid = grid_mult ? (long long)sm_count() * ctas_per_sm * grid_mult : ntiles;
    if (grid > ntiles) grid = ntiles;
    static int stream_stores = -1;
    if (stream_stores < 0) {
        const char *e = getenv("HB_TMA_STCS");   // tuning knob: st.cs for the output (no measurable effect)
        stream_stores = e ? atoi(e) : 0;
    }
    kern<<<(unsigned)grid, dim3(32, BY), G::SMEM_BYTES, s>>>(p, tmap, ntx, (int)ntiles, stream_stores);
    g_launches++;
    return HB_OK;
}

template <typename MASK>
bool mask_equals(const LocalParams &p) {
    for (int k = 0; k < MASK::kTaps; ++k) {
        const bool on = (p.dom[k >> 5] >> (k & 31)) & 1u;
        const float c = MASK::coef(k);
        if (on != (c != 0.0f)) return false;       // specialised kernels skip exactly the Domain holes
        if (on && p.coef.f[k] != c) return false;
    }
    return true;
}

// tile height variant (rows per thread x warps per CTA); HB_LOCAL_TMA_SHAPE=RPTxBY overrides for tuning
template <int SX, int SY, typename MASK>
int launch_shape(const LocalParams &p, cudaStream_t s) {
    static int shape = -1;
    if (shape < 0) {
        const char *e = getenv("HB_LOCAL_TMA_SHAPE");
        shape = e ? atoi(e) : 0;
    }
    switch (shape) {
#ifdef HB_TUNE_SHAPES
    case 1: return launch_variant<SX, SY, MASK, 4, 8, 2>(p, s);
    case 2: return launch_variant<SX, SY, MASK, 4, 8, 3>(p, s);
    case 3: return launch_variant<SX, SY, MASK, 4, 8, 6>(p, s);
    case 4: return launch_variant<SX, SY, MASK, 8, 4, 4>(p, s);
#endif
    case 6: return launch_variant<SX, SY, MASK, 4, 8, 2>(p, s);
    case 7: return launch_variant<SX, SY, MASK, 4, 8, 4>(p, s);   // the 4-stage ring of the persistent design
    default: return launch_variant<SX, SY, MASK, 4, 8, 1>(p, s);  // one stage: with one tile per CTA the overlap comes from the co-resident CTAs
    }
}

}  // namespace

int launch_local_tma_f32(const LocalParams &p, bool all_taps_visited, cudaStream_t s) {
    static int disabled = -1;
    if (disabled < 0) {
        const char *e = getenv("HB_DISABLE_TMA");
        disabled = (e && atoi(e)) ? 1 : 0;
    }
    if (disabled) return HB_ERR_UNSUPPORTED;
    if (p.size_x != p.size_y) return HB_ERR_UNSUPPORTED;
    if (!tma_addressable(p.in, HB_F32, p.in_stride) || (p.in_ox & 3)) return HB_ERR_UNSUPPORTED;  // box origins must be 16-byte aligned
    if (p.size_x == 3) {
        if (mask_equals<MaskSobel3X>(p)) return launch_shape<3, 3, MaskSobel3X>(p, s);
        if (mask_equals<MaskSobel3Y>(p)) return launch_shape<3, 3, MaskSobel3Y>(p, s);
        if (mask_equals<MaskLaplace3D>(p)) return launch_shape<3, 3, MaskLaplace3D>(p, s);
        if (mask_equals<MaskLaplace3N>(p)) return launch_shape<3, 3, MaskLaplace3N>(p, s);
        if (mask_equals<MaskIdentity3>(p)) return launch_shape<3, 3, MaskIdentity3>(p, s);
        if (all_taps_visited) return launch_shape<3, 3, MaskRuntime>(p, s);
    } else if (p.size_x == 5) {
        if (mask_equals<MaskSobel5X>(p)) return launch_shape<5, 5, MaskSobel5X>(p, s);
        if (mask_equals<MaskSobel5Y>(p)) return launch_shape<5, 5, MaskSobel5Y>(p, s);
        if (mask_equals<MaskLaplace5>(p)) return launch_shape<5, 5, MaskLaplace5>(p, s);
        if (all_taps_visited) return launch_shape<5, 5, MaskRuntime>(p, s);
    } else if (p.size_x == 7) {
        if (all_taps_visited) return launch_shape<7, 7, MaskRuntime>(p, s);
    }
    return HB_ERR_UNSUPPORTED;
}

}  // namespace hb
