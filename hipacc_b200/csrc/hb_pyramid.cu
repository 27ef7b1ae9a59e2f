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

#include <cstring>

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

}  // namespace hb

using namespace hb;

extern "C" int hb_pyr_down(const hb_pyr_down_desc *d, void *stream) {
    HB_REQUIRE(d && d->coef_f32, HB_ERR_INVALID, "hb_pyr_down: null descriptor / mask");
    hb_view fine = norm_view(d->fine), coarse = norm_view(d->coarse);
    HB_REQUIRE(view_ok(fine) && view_ok(coarse) && fine.dtype == HB_F32 && coarse.dtype == HB_F32, HB_ERR_INVALID, "hb_pyr_down: needs valid f32 views");
    HB_REQUIRE(d->size == 3 || d->size == 5 || d->size == 7, HB_ERR_UNSUPPORTED, "hb_pyr_down: mask size %d unsupported (3,5,7); no CPU fallback", d->size);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = HB_OK;
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
    PyrUpParams p{plane_of(cg), plane_of(cl), plane_of(fg), plane_of(fl)};
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_pyr_up");
    dim3 grid((fg.width + 31) / 32, (fg.height + 31) / 32);
    pyr_up_kernel<<<grid, dim3(32, 8), 0, s>>>(p);
    g_launches++;
    return scope.finish();
}
