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

}  // namespace hb

using namespace hb;

extern "C" int hb_harris(const hb_harris_desc *d, void *stream) {
    HB_REQUIRE(d, HB_ERR_INVALID, "hb_harris: null descriptor");
    hb_view in = norm_view(d->in), out = norm_view(d->out);
    HB_REQUIRE(view_ok(in) && view_ok(out) && in.dtype == HB_U8 && out.dtype == HB_U8, HB_ERR_INVALID, "hb_harris: needs valid u8 views");
    HB_REQUIRE(in.width == out.width && in.height == out.height, HB_ERR_INVALID, "hb_harris: input and output regions must have the same size");
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
    harris_fused_kernel<<<grid, dim3(HBX, HBY), 0, s>>>(p);
    g_launches++;
    return scope.finish();
}
