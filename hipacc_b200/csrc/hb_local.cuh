// hb_local.cuh -- parameter block and arithmetic helpers shared by the local-operator kernels
// (hb_local.cu: all-threads staged tiles, every type / mode; hb_local_tma.cu: TMA-pipelined float path).
#pragma once
#include "hb_common.cuh"
#include "hb_internal.h"

namespace hb {

constexpr int kMaxTaps = 169;  // up to 13 x 13

struct LocalParams {
    const void *in;
    void *out;
    int in_stride, in_iw, in_ih;
    Window win;
    int in_ox, in_oy;  // IS-relative (0,0) reads input pixel (in_ox, in_oy)  (dsl/image.hpp:412)
    int out_stride, out_ox, out_oy, is_w, is_h;
    int size_x, size_y;
    int reduce_mode, tap, acc_s16, epilogue;
    float epi_f[3];
    int epi_i[3];
    float cval_f;
    int cval_i;
    unsigned dom[6];  // bit k: tap k (row-major) is visited
    int first_tap;    // first visited tap: it initialises the accumulator (dsl/kernel.hpp:250,279)
    union {
        float f[kMaxTaps];
        int i[kMaxTaps];
    } coef;
    // packed-FP32 multiply (FMUL2, sm_100a) operands for the float SUM-of-products path of masks up to 7 x 7, rows
    // padded to 8 floats so that every pair is 8-byte aligned in the constant bank:
    //   cpair[0][dy][dx] = c[dy][dx]      -> pairs (c0,c1) (c2,c3) ...   for pixels whose first tap sits in an even register
    //   cpair[1][dy][dx] = c[dy][dx + 1]  -> pairs (c1,c2) (c3,c4) ...   for the others
    //   cdup[k]          = (c[k], c[k])   -> one coefficient for two adjacent channel elements (uchar4 images)
    alignas(16) float cpair[2][7][8];   // 16-byte aligned: a misaligned pair costs two uniform moves per use
    alignas(16) float cdup[49][2];
};

// ---- arithmetic in the accumulation type (float: separately rounded mul / add) ----
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ int mul_rn(int a, int b) { return a * b; }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ int add_rn(int a, int b) { return a + b; }

template <typename TS> __device__ __forceinline__ TS fold_identity(int mode);
template <> __device__ __forceinline__ float fold_identity<float>(int mode) {
    return mode == HB_REDUCE_SUM ? 0.0f : mode == HB_REDUCE_PROD ? 1.0f : mode == HB_REDUCE_MIN ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
}
template <> __device__ __forceinline__ int fold_identity<int>(int mode) {
    return mode == HB_REDUCE_SUM ? 0 : mode == HB_REDUCE_PROD ? 1 : mode == HB_REDUCE_MIN ? 2147483647 : (-2147483647 - 1);
}
template <typename TS>
__device__ __forceinline__ TS fold(TS acc, TS v, int mode) {
    switch (mode) {
    case HB_REDUCE_SUM: return add_rn(acc, v);
    case HB_REDUCE_MIN: return v < acc ? v : acc;  // hipacc::math::min(fun(), result), dsl/kernel.hpp:256
    case HB_REDUCE_MAX: return v > acc ? v : acc;
    default: return mul_rn(acc, v);
    }
}

// two separately rounded products in one instruction: (a.lo * b.lo, a.hi * b.hi).  FMUL2 issues at the scalar FMUL
// rate (profiles/r1h_fp32x2_probe.txt).  The additions stay scalar FADDs: ptxas contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 even with --fmad=false, which would break the separately-rounded contract.
__device__ __forceinline__ void mul2_rn(float a0, float a1, const float *b_pair, float &p0, float &p1) {
    unsigned long long a, b = *reinterpret_cast<const unsigned long long *>(b_pair), r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(r));
}

template <typename TS> __device__ __forceinline__ TS coef_of(const LocalParams &p, int k);
template <> __device__ __forceinline__ float coef_of<float>(const LocalParams &p, int k) { return p.coef.f[k]; }
template <> __device__ __forceinline__ int coef_of<int>(const LocalParams &p, int k) { return p.coef.i[k]; }

template <typename TO>
__device__ __forceinline__ TO epilogue(float acc, const LocalParams &p) {
    switch (p.epilogue) {
    case HB_EPI_ADD_CAST: return cast_out<TO, float>(__fadd_rn(acc, p.epi_f[0]));
    case HB_EPI_ADD_CLAMP_CAST: {
        float v = __fadd_rn(acc, p.epi_f[0]);
        v = v < p.epi_f[2] ? v : p.epi_f[2];
        v = v > p.epi_f[1] ? v : p.epi_f[1];
        return cast_out<TO, float>(v);
    }
    case HB_EPI_DIVI_CAST: return cast_out<TO, int>(__float2int_rz(acc) / p.epi_i[0]);
    case HB_EPI_DIVF_CAST: return cast_out<TO, float>(__fdiv_rn(acc, p.epi_f[0]));
    default: return cast_out<TO, float>(acc);
    }
}
template <typename TO>
__device__ __forceinline__ TO epilogue(int acc, const LocalParams &p) {
    if (p.acc_s16) acc = (int)(short)acc;
    switch (p.epilogue) {
    case HB_EPI_ADD_CAST: return cast_out<TO, int>(acc + p.epi_i[0]);
    case HB_EPI_ADD_CLAMP_CAST: {
        int v = acc + p.epi_i[0];
        v = min(v, p.epi_i[2]);
        v = max(v, p.epi_i[1]);
        return cast_out<TO, int>(v);
    }
    case HB_EPI_DIVI_CAST: return cast_out<TO, int>(acc / p.epi_i[0]);
    case HB_EPI_DIVF_CAST: return cast_out<TO, float>(__fdiv_rn((float)acc, p.epi_f[0]));
    default: return cast_out<TO, int>(acc);
    }
}

template <typename TS> __device__ __forceinline__ TS cval_of(const LocalParams &p);
template <> __device__ __forceinline__ float cval_of<float>(const LocalParams &p) { return p.cval_f; }
template <> __device__ __forceinline__ int cval_of<int>(const LocalParams &p) { return p.cval_i; }

// hb_local_tma.cu: persistent, TMA-pipelined kernel for float -> float SUM-of-products stencils (3x3, 5x5, 7x7).
// Returns HB_OK when it launched, HB_ERR_UNSUPPORTED when the operator / image is not eligible (caller falls
// through to the staged kernels).
int launch_local_tma_f32(const LocalParams &p, bool holes_allowed, cudaStream_t s);

// hb_local_pair.cu: float SUM of products over every tap with packed multiplies and packed additions (two pixels per
// FMUL2 / FADD2), 3x3 / 5x5 / 7x7.  dtypes after as_channels(); ch = 4 for uchar4 images.  HB_ERR_UNSUPPORTED = not taken.
int launch_local_pair(const LocalParams &p, int in_dtype, int out_dtype, int ch, cudaStream_t s);

}  // namespace hb
