// hb_local_pair.cu -- float SUM-of-products stencils with PACKED multiplies AND packed additions (sm_100a FMUL2 / FADD2).
//
// The DSL folds `mask() * input(mask)` over the taps in row-major order with separately rounded float operations
// (dsl/kernel.hpp:241-267); the contract is bit-identity with that fold, so the per-pixel operation count cannot be
// reduced -- but two PIXELS can share every instruction: a thread keeps the accumulators of two outputs in one 64-bit
// register pair, forms both products with one FMUL2 (input pair x duplicated coefficient) and adds them with one FADD2.
// 25 taps cost 12.5 + 12 issue slots per pixel instead of 12.5 + 25 (hb_local.cu) or 50 scalar.
//
// Which two pixels?  A register pair must be (even, odd) aligned, and a window that slides by one column (or row) per
// tap would need every staged value in an even AND an odd register: twice the shared-memory traffic, which then bounds
// the kernel (measured: the adjacent-pixel version was l1tex-bound at 75 %).  So the pair is (x, y) and (x + 64, y):
// the two halves of the 128-wide tile.  The tile is staged as float2 { column c, column c + 64 }, every tap of every
// pixel pair is then ONE aligned 64-bit operand, and each staged value is read once per row it contributes to.
//
// ptxas contracts `mul.rn.f32x2` feeding `add.rn.f32x2` into FFMA2 even with --fmad=false, which would round once
// instead of twice.  It does not when the product reaches the addition with its halves exchanged (the exchange is a
// free operand modifier, `.F32x2.LO_HI`), so the accumulator pairs are kept in EXCHANGED order: lo = right-half pixel,
// hi = left-half pixel.  tests/test_abi.py greps the SASS of this object: FFMA2 must not appear.
//
// Shared-memory layout of one staged row: 16-byte units (2 float2) u = 0 .. QC/2-1; a thread reads units 2*lx + k, i.e.
// with a lane stride of two units, which would be a 2-way bank conflict -- even units are stored in the first half of
// the row and odd units in the second, so the lanes of a quarter-warp read consecutive units.
#include "hb_local.cuh"

#include <cstdlib>

namespace hb {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// acc (exchanged order) += halves-exchanged(prod)
__device__ __forceinline__ u64 add2_x(u64 acc, u64 prod) {
    float lo, hi;
    upk(prod, lo, hi);
    return add2(acc, pk(hi, lo));
}

constexpr int PTW = 128, PHALF = PTW / 2, PTH = 32;   // tile 128 x 32; (PHALF / 4) x (PTH / PRPT) threads, PRPT = rows per thread

template <typename TO, int EPI>
__device__ __forceinline__ TO pair_epilogue(float acc, const LocalParams &p) {
    if (EPI == HB_EPI_CAST) return cast_out<TO, float>(acc);
    if (EPI == HB_EPI_ADD_CAST) return cast_out<TO, float>(__fadd_rn(acc, p.epi_f[0]));
    return epilogue<TO>(acc, p);
}

// float index of unit u within a staged row of QC column pairs (even units first, then the odd ones)
template <int QC>
__device__ __forceinline__ constexpr int unit_pos(int u) { return ((u >> 1) + (u & 1) * (QC / 4)) * 4; }

// EPI: HB_EPI_CAST / HB_EPI_ADD_CAST compile-time, -1 = run-time switch (hb_local.cuh)
template <typename TI, typename TO, int SX, int SY, int CH, int EPI, int PRPT>
__global__ void __launch_bounds__((PHALF / 4) * (PTH / PRPT)) local_pair_kernel(const __grid_constant__ LocalParams p) {
    constexpr int PNT = (PHALF / 4) * (PTH / PRPT);
    constexpr int HX = (SX / 2) * CH, HY = SY / 2;
    constexpr int HXP = round_up(HX, 4);
    constexpr int QC = PHALF + 2 * HXP;     // column pairs per staged row (multiple of 4)
    constexpr int ROWS = PTH + SY - 1;
    constexpr int BASE = HXP - HX;          // window column of pixel 0's first tap
    constexpr int PITCH = QC * 2;           // floats per staged row
    __shared__ __align__(16) float q[ROWS * PITCH];

    const int tid = threadIdx.x;
    const int lx = tid & 15, ly = tid >> 4;
    const int gx0 = blockIdx.x * PTW, gy0 = blockIdx.y * PTH;

    // ---- stage: q[r][c] = { in(x_start + c, y_start + r), in(x_start + c + 64, y_start + r) } as float
    {
        const TI *in = static_cast<const TI *>(p.in);
        const int x_start = p.in_ox + gx0 - HXP, y_start = p.in_oy + gy0 - HY;
        const Window w = p.win;
        const bool interior = x_start >= w.lo_x * CH && x_start + PTW + 2 * HXP <= w.hi_x * CH && y_start >= w.lo_y && y_start + ROWS <= w.hi_y;
        const bool aligned = ((reinterpret_cast<uintptr_t>(in) + (size_t)x_start * sizeof(TI)) % (4 * sizeof(TI)) == 0) && (p.in_stride % 4 == 0);
        if (interior && aligned) {
            constexpr int VPR = QC / 4;   // vectors of 4 column pairs per row = 2 units
            constexpr int NV = ROWS * VPR;
            // byte pixels: a whole tile is PER vectors per thread; wider pixels go PER at a time (registers)
            constexpr int PER = sizeof(TI) == 1 ? (NV + PNT - 1) / PNT : 2;
            const TI *base = in + (size_t)y_start * p.in_stride + x_start;
            // all of a thread's global loads are in flight before the first conversion (the staging phase is latency-bound)
            for (int v0 = tid; v0 < NV; v0 += PER * PNT) {
            typename Vec4<TI>::type ga[PER], gb[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int v = v0 + k * PNT;
                if (v < NV) {
                    const int r = v / VPR, m = v - r * VPR;
                    const TI *src = base + (size_t)r * p.in_stride + 4 * m;
                    ga[k] = __ldg(reinterpret_cast<const typename Vec4<TI>::type *>(src));
                    gb[k] = __ldg(reinterpret_cast<const typename Vec4<TI>::type *>(src + PHALF));
                }
            }
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int v = v0 + k * PNT;
                if (v < NV) {
                    const int r = v / VPR, m = v - r * VPR;
                    float *dst = q + r * PITCH + 4 * m;   // unit 2m -> even plane slot m, unit 2m+1 -> odd plane slot m
                    *reinterpret_cast<float4 *>(dst) = make_float4((float)ga[k].x, (float)gb[k].x, (float)ga[k].y, (float)gb[k].y);
                    *reinterpret_cast<float4 *>(dst + QC) = make_float4((float)ga[k].z, (float)gb[k].z, (float)ga[k].w, (float)gb[k].w);
                }
            }
            }
        } else {
            ImgRef<TI> im{in, p.in_stride, p.in_iw, p.in_ih};
            const TI cv = (TI)p.cval_f;
            for (int e = tid; e < ROWS * QC * 2; e += PNT) {
                const int r = e / (QC * 2), rem = e - r * (QC * 2);
                const int half = rem / QC, c = rem - half * QC;
                const float val = (float)fetch_bh_elem<TI, CH>(im, w, x_start + c + half * PHALF, y_start + r, cv);
                q[r * PITCH + unit_pos<QC>(c >> 1) + (c & 1) * 2 + half] = val;
            }
        }
    }
    __syncthreads();

    u64 acc[PRPT][4];   // [row][pixel]: lo = right-half pixel, hi = left-half pixel (exchanged order, see the file header)
    const int r0 = ly * PRPT;
    const float *qt = q + r0 * PITCH + 4 * lx;   // unit 2*lx + k  ->  float offset unit_pos(k) + 4*lx
#pragma unroll
    for (int ir = 0; ir < PRPT + SY - 1; ++ir) {
        const float *row = qt + ir * PITCH;
        constexpr int NCOL = 4 + (SX - 1) * CH;                   // window columns BASE .. BASE + NCOL - 1
        constexpr int U0 = BASE / 2, U1 = (BASE + NCOL - 1) / 2;  // units that hold them
        if (CH == 1) {
            u64 w[2 * (U1 - U0 + 1)];   // w[j - 2*U0] = column pair j
#pragma unroll
            for (int u = U0; u <= U1; ++u) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(row + unit_pos<QC>(u));
                w[2 * (u - U0)] = v.x; w[2 * (u - U0) + 1] = v.y;
            }
#pragma unroll
            for (int r = 0; r < PRPT; ++r) {
                const int dy = ir - r;
                if (dy < 0 || dy >= SY) continue;
#pragma unroll
                for (int dx = 0; dx < SX; ++dx) {
                    const int k = dy * SX + dx;
                    const u64 c2 = *reinterpret_cast<const u64 *>(p.cdup[k]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const u64 a = w[BASE + i + dx - 2 * U0];
                        if (k == 0) {
                            // the first tap initialises the accumulator (dsl/kernel.hpp:250): two scalar products
                            // written straight into the exchanged order
                            float lo, hi;
                            upk(a, lo, hi);
                            acc[r][i] = pk(__fmul_rn(hi, p.coef.f[0]), __fmul_rn(lo, p.coef.f[0]));
                        } else {
                            acc[r][i] = add2_x(acc[r][i], mul2(a, c2));
                        }
                    }
                }
            }
        } else {
            // interleaved channels: tap dx reads columns BASE + dx*CH .. + 3 (CH is a multiple of 4: two whole units)
#pragma unroll
            for (int dx = 0; dx < SX; ++dx) {
                const int u = (BASE + dx * CH) / 2;
                const ulonglong2 v0 = *reinterpret_cast<const ulonglong2 *>(row + unit_pos<QC>(u));
                const ulonglong2 v1 = *reinterpret_cast<const ulonglong2 *>(row + unit_pos<QC>(u + 1));
                const u64 w[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
                for (int r = 0; r < PRPT; ++r) {
                    const int dy = ir - r;
                    if (dy < 0 || dy >= SY) continue;
                    const int k = dy * SX + dx;
                    const u64 c2 = *reinterpret_cast<const u64 *>(p.cdup[k]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (k == 0) {
                            float lo, hi;
                            upk(w[i], lo, hi);
                            acc[r][i] = pk(__fmul_rn(hi, p.coef.f[0]), __fmul_rn(lo, p.coef.f[0]));
                        } else {
                            acc[r][i] = add2_x(acc[r][i], mul2(w[i], c2));
                        }
                    }
                }
            }
        }
    }

    TO *out = static_cast<TO *>(p.out);
#pragma unroll
    for (int r = 0; r < PRPT; ++r) {
        const int gy = gy0 + r0 + r;
        if (gy >= p.is_h) continue;
        float a[2][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) upk(acc[r][i], a[1][i], a[0][i]);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int gx = gx0 + 4 * lx + half * PHALF;
            if (gx >= p.is_w) continue;
            TO o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = pair_epilogue<TO, EPI>(a[half][i], p);
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
}

template <typename TI, typename TO, int CH, int PRPT>
static int launch_pair_t(const LocalParams &p, cudaStream_t s) {
    constexpr int PNT = (PHALF / 4) * (PTH / PRPT);
    dim3 grid((p.is_w + PTW - 1) / PTW, (p.is_h + PTH - 1) / PTH);
#define HB_PAIR(SXV, SYV)                                                                                             \
    if (p.size_x == SXV && p.size_y == SYV) {                                                                         \
        if (p.epilogue == HB_EPI_CAST) local_pair_kernel<TI, TO, SXV, SYV, CH, HB_EPI_CAST, PRPT><<<grid, PNT, 0, s>>>(p);   \
        else if (p.epilogue == HB_EPI_ADD_CAST) local_pair_kernel<TI, TO, SXV, SYV, CH, HB_EPI_ADD_CAST, PRPT><<<grid, PNT, 0, s>>>(p); \
        else local_pair_kernel<TI, TO, SXV, SYV, CH, -1, PRPT><<<grid, PNT, 0, s>>>(p);                                \
        g_launches++;                                                                                                 \
        return HB_OK;                                                                                                 \
    }
    HB_PAIR(3, 3)
    HB_PAIR(5, 5)
    HB_PAIR(7, 7)
#undef HB_PAIR
    return HB_ERR_UNSUPPORTED;
}

// float SUM of coef * pixel over every tap (VAR 0 of hb_local.cu), masks 3x3 / 5x5 / 7x7.  in_dtype / out_dtype after
// as_channels(); ch = 4 for uchar4 images.  HB_ERR_UNSUPPORTED = not taken, the caller falls through.
int launch_local_pair(const LocalParams &p, int in_dtype, int out_dtype, int ch, cudaStream_t s) {
    // rows per thread: 4 (128-thread CTAs, fewest shared-memory reads per pixel) for large images; 2 (256-thread CTAs, half
    // the serial work per thread) for small ones, where the time of the last wave of CTAs -- one CTA's latency chain --
    // is a visible part of the kernel.  HB_PAIR_RPT = 2 / 4 overrides.
    static int force = -1;
    if (force < 0) { const char *e = getenv("HB_PAIR_RPT"); force = e ? atoi(e) : 0; }
    const long long tiles = (long long)((p.is_w + PTW - 1) / PTW) * ((p.is_h + PTH - 1) / PTH);
    const bool small = force ? force == 2 : tiles < 256LL * sm_count();
#define HB_PAIR_T(TI, TO, CH) return small ? launch_pair_t<TI, TO, CH, 2>(p, s) : launch_pair_t<TI, TO, CH, 4>(p, s)
    if (ch == 4) HB_PAIR_T(uchar, uchar, 4);
    if (in_dtype == HB_U8 && out_dtype == HB_U8) HB_PAIR_T(uchar, uchar, 1);
    if (in_dtype == HB_S8 && out_dtype == HB_S8) HB_PAIR_T(signed char, signed char, 1);
    if (in_dtype == HB_F32 && out_dtype == HB_F32) HB_PAIR_T(float, float, 1);
#undef HB_PAIR_T
    return HB_ERR_UNSUPPORTED;
}

}  // namespace hb
