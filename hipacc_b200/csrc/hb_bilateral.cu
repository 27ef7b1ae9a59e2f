// hb_bilateral.cu -- bilateral filter (the `iterate(dom, ...)` body of
// samples-public/3_Preprocessing/Bilateral_Filter/src/main.cpp:63-77) for sm_100a.
//
// 169 exponentials per pixel at 13x13: this operator is bound by the MUFU pipe (ex2, 16 per clock per
// SM), not by HBM.  The tile + halo is staged once into shared memory as float; each
// thread owns 4 x 4 pixels and walks the staged rows once (row-stationary), so shared-memory
// traffic is ~5 16-byte loads per 52 taps.  The per-pixel tap order (row-major) follows the sample.
// Per tap the sample computes  s = expf(-c_r*diff*diff) * mask ; d += s ; p += s*in.  To keep the FP32
// pipe below the MUFU pipe the kernel evaluates the same value as
//     s = ex2( (diff*diff) * (-c_r*log2 e) + log2(mask) )          (one FMA + one MUFU.EX2)
//     d += s ; p = fma(s, in, p)
// i.e. 5 FP32 instructions per tap instead of 9.  This is a float pipeline under the 1e-5 relative
// contract (measured error vs libm ~1e-6, tests/test_gpu_parity.py); on uchar output a 1-LSB flip at
// rounding ties is possible where the sample itself tolerates |diff| <= 1.  Masks with a non-positive
// coefficient take the unfolded form  s = ex2(...) * mask.
#include "hb_common.cuh"
#include "hb_internal.h"

#include <cmath>
#include <cstring>

namespace hb {

struct BilateralParams {
    const void *in;
    void *out;
    int in_stride, in_iw, in_ih;
    Window win;
    int in_ox, in_oy;
    int out_stride, out_ox, out_oy, is_w, is_h;
    float k2;      // -c_r * log2(e)
    float cval;
    float coef[169];  // FOLD: log2(mask), else mask
};

constexpr int BTW = 128, BRPT = 4, BBX = 32, BBY = 8, BTH = BBY * BRPT;

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <typename T, int S, bool FOLD>
__global__ void __launch_bounds__(BBX *BBY) bilateral_kernel(const __grid_constant__ BilateralParams p) {
    constexpr int H = S / 2;
    constexpr int HXP = round_up(H, 4);
    constexpr int TWS = BTW + 2 * HXP;
    constexpr int ROWS = BTH + S - 1;
    constexpr int WIN = 4 + 2 * HXP;
    extern __shared__ __align__(16) float btile[];

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * BBX + tx;
    const int gx0 = blockIdx.x * BTW, gy0 = blockIdx.y * BTH;
    stage_tile<T, float, ROWS, TWS, BBX * BBY>(btile, static_cast<const T *>(p.in), p.in_stride, p.in_iw, p.in_ih, p.win, (T)p.cval,
                                               p.in_ox + gx0 - HXP, p.in_oy + gy0 - H, tid);
    __syncthreads();

    const int r0 = ty * BRPT;
    float center[BRPT][4], dsum[BRPT][4], psum[BRPT][4];
#pragma unroll
    for (int r = 0; r < BRPT; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            center[r][i] = btile[(r0 + r + H) * TWS + HXP + 4 * tx + i];
            dsum[r][i] = 0.0f;
            psum[r][i] = 0.0f;
        }

    const float k2 = p.k2;
#pragma unroll 1
    for (int ir = 0; ir < BRPT + S - 1; ++ir) {
        float w[WIN];
        const float *row = btile + (r0 + ir) * TWS + 4 * tx;
#pragma unroll
        for (int q = 0; q < WIN / 4; ++q) {
            float t[4];
            load4(row + 4 * q, t);
            w[4 * q] = t[0]; w[4 * q + 1] = t[1]; w[4 * q + 2] = t[2]; w[4 * q + 3] = t[3];
        }
#pragma unroll
        for (int r = 0; r < BRPT; ++r) {
            const int dy = ir - r;
            if (dy < 0 || dy >= S) continue;  // warp-uniform
#pragma unroll
            for (int dx = 0; dx < S; ++dx) {
                const float m = p.coef[dy * S + dx];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float v = w[HXP - H + i + dx];
                    const float diff = __fadd_rn(v, -center[r][i]);
                    const float t = __fmul_rn(diff, diff);
                    float s;
                    if (FOLD) {
                        s = ex2_approx(__fmaf_rn(t, k2, m));
                    } else {
                        s = __fmul_rn(ex2_approx(__fmul_rn(t, k2)), m);
                    }
                    dsum[r][i] = __fadd_rn(dsum[r][i], s);
                    psum[r][i] = __fmaf_rn(s, v, psum[r][i]);
                }
            }
        }
    }

    T *out = static_cast<T *>(p.out);
    const int gx = gx0 + 4 * tx;
#pragma unroll
    for (int r = 0; r < BRPT; ++r) {
        const int gy = gy0 + r0 + r;
        if (gy >= p.is_h || gx >= p.is_w) continue;
        T o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float q = __fdiv_rn(psum[r][i], dsum[r][i]);
            if (DtypeOf<T>::v == HB_F32) o[i] = (T)q;
            else o[i] = cast_out<T, float>(__fadd_rn(q, 0.5f));
        }
        T *dst = out + (size_t)(p.out_oy + gy) * p.out_stride + p.out_ox + gx;
        if (gx + 3 < p.is_w && (reinterpret_cast<uintptr_t>(dst) % (4 * sizeof(T)) == 0)) {
            store4(dst, o);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (gx + i < p.is_w) dst[i] = o[i];
        }
    }
}

template <typename T, int S, bool FOLD>
static void launch_bilateral_v(const BilateralParams &p, cudaStream_t s) {
    constexpr int HXP = round_up(S / 2, 4);
    constexpr size_t smem = (size_t)(BTH + S - 1) * (BTW + 2 * HXP) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(bilateral_kernel<T, S, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    dim3 grid((p.is_w + BTW - 1) / BTW, (p.is_h + BTH - 1) / BTH);
    bilateral_kernel<T, S, FOLD><<<grid, dim3(BBX, BBY), smem, s>>>(p);
    g_launches++;
}

template <typename T, int S>
static void launch_bilateral(const BilateralParams &p, bool fold, cudaStream_t s) {
    if (fold) launch_bilateral_v<T, S, true>(p, s);
    else launch_bilateral_v<T, S, false>(p, s);
}

template <typename T>
static int dispatch_bilateral(const BilateralParams &p, int size, bool fold, cudaStream_t s) {
    switch (size) {
    case 3: launch_bilateral<T, 3>(p, fold, s); return HB_OK;
    case 5: launch_bilateral<T, 5>(p, fold, s); return HB_OK;
    case 7: launch_bilateral<T, 7>(p, fold, s); return HB_OK;
    case 13: launch_bilateral<T, 13>(p, fold, s); return HB_OK;
    default: return HB_ERR_UNSUPPORTED;
    }
}

}  // namespace hb

using namespace hb;

extern "C" int hb_bilateral(const hb_bilateral_desc *d, void *stream) {
    HB_REQUIRE(d && d->coef_f32, HB_ERR_INVALID, "hb_bilateral: null descriptor / mask");
    hb_view in = norm_view(d->in), out = norm_view(d->out);
    HB_REQUIRE(view_ok(in) && view_ok(out) && in.dtype == out.dtype, HB_ERR_INVALID, "hb_bilateral: malformed views");
    HB_REQUIRE(d->sigma_r != 0, HB_ERR_INVALID, "hb_bilateral: sigma_r == 0");
    BilateralParams p;
    memset(&p, 0, sizeof(p));
    p.in = in.data; p.out = out.data;
    p.in_stride = in.stride; p.in_iw = in.img_width; p.in_ih = in.img_height;
    p.win = Window{in.offset_x, in.offset_x + in.width, in.offset_y - in.ghost_top, in.offset_y + in.height + in.ghost_bottom, d->boundary};
    p.in_ox = in.offset_x; p.in_oy = in.offset_y;
    p.out_stride = out.stride; p.out_ox = out.offset_x; p.out_oy = out.offset_y; p.is_w = out.width; p.is_h = out.height;
    const float c_r = 0.5f / (float)(d->sigma_r * d->sigma_r);  // float c_r = 0.5f/(sigma_r*sigma_r)
    p.k2 = (float)(-(double)c_r * 1.4426950408889634);
    p.cval = (float)d->boundary_const;
    if (d->size * d->size > 169 || d->size <= 0) {
        log_msg(2, "hb_bilateral: mask size %d unsupported", d->size);
        return HB_ERR_UNSUPPORTED;
    }
    bool fold = true;
    for (int k = 0; k < d->size * d->size; ++k) fold = fold && d->coef_f32[k] > 0.0f;
    for (int k = 0; k < d->size * d->size; ++k) p.coef[k] = fold ? (float)log2((double)d->coef_f32[k]) : d->coef_f32[k];
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_bilateral");
    int rc = HB_ERR_UNSUPPORTED;
    if (in.dtype == HB_U8) rc = dispatch_bilateral<uchar>(p, d->size, fold, s);
    else if (in.dtype == HB_F32) rc = dispatch_bilateral<float>(p, d->size, fold, s);
    HB_REQUIRE(rc == HB_OK, HB_ERR_UNSUPPORTED, "hb_bilateral: no device kernel for dtype %d size %d (sizes 3,5,7,13; u8/f32); no CPU fallback",
               in.dtype, d->size);
    return scope.finish();
}
