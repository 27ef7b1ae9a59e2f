// fp32x2_probe.cu -- does packed FP32 (FMUL2 / FADD2 / FFMA2, sm_100a) issue at the scalar rate?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/probe/fp32x2_probe tools/probe/fp32x2_probe.cu
// Prints lane-operations per clock per SM for scalar FADD+FMUL, packed FADD2+FMUL2 and FFMA vs FFMA2.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 a) { float x, y; asm("mov.b64 {%0,%1},%2;" : "=f"(x), "=f"(y) : "l"(a)); return x + y; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
constexpr int NACC = 8;
__global__ void k_scalar(float *out, float a, float b, int iters) {
    float acc[2 * NACC];
    for (int i = 0; i < 2 * NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 2 * NACC; ++i) acc[i] = __fmul_rn(__fadd_rn(acc[i], b), a);
    float s = 0; for (int i = 0; i < 2 * NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float *out, float a, float b, int iters) {
    u64 acc[NACC]; const u64 A = pk(a, a), B = pk(b, b);
    for (int i = 0; i < NACC; ++i) acc[i] = pk(threadIdx.x * 1e-3f + i, i + 0.5f);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = mul2(add2(acc[i], B), A);
    float s = 0; for (int i = 0; i < NACC; ++i) s += lo(acc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float *out, float a, float b, int iters) {
    float acc[2 * NACC];
    for (int i = 0; i < 2 * NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 2 * NACC; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
    float s = 0; for (int i = 0; i < 2 * NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float a, float b, int iters) {
    u64 acc[NACC]; const u64 A = pk(a, a), B = pk(b, b);
    for (int i = 0; i < NACC; ++i) acc[i] = pk(threadIdx.x * 1e-3f + i, i + 0.5f);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma2(acc[i], A, B);
    float s = 0; for (int i = 0; i < NACC; ++i) s += lo(acc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// MUFU.EX2 next to FP32 work: ex2 + n_fp FP32 ops per element
template <int NFP>
__global__ void k_mufu(float *out, float a, float b, int iters) {
    float acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = acc[i];
#pragma unroll
            for (int k = 0; k < NFP; ++k) x = __fmaf_rn(x, a, b);
            float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
            acc[i] = r;
        }
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K> static float timeit(K kern, float *out, int blocks, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<blocks, 256>>>(out, 0.999f, 0.001f, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); kern<<<blocks, 256>>>(out, 0.999f, 0.001f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount, blocks = sms * 8, iters = 4096;
    float *out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    const double lanes = (double)blocks * 256 * 2 * NACC * iters;  // lane-results per kernel (each = 2 ops or 1 fma)
    const double clk = clk_khz * 1e3;
    struct { const char *name; float ms; double ops_per_lane; } r[] = {
        {"scalar FADD+FMUL", timeit(k_scalar, out, blocks, iters), 2}, {"packed FADD2+FMUL2", timeit(k_packed, out, blocks, iters), 2},
        {"scalar FFMA", timeit(k_ffma, out, blocks, iters), 1}, {"packed FFMA2", timeit(k_ffma2, out, blocks, iters), 1}};
    printf("%s, %d SMs, %.0f MHz nominal\n", pr.name, sms, clk / 1e6);
    for (auto &x : r) printf("%-22s %8.3f ms  %7.1f lane-instr/clk/SM (at nominal clock)\n", x.name, x.ms, lanes * x.ops_per_lane / (x.ms * 1e-3) / clk / sms);
    const double el = (double)blocks * 256 * 8 * iters;
    float m0 = timeit(k_mufu<0>, out, blocks, iters), m2 = timeit(k_mufu<2>, out, blocks, iters), m5 = timeit(k_mufu<5>, out, blocks, iters), m8 = timeit(k_mufu<8>, out, blocks, iters);
    printf("ex2 + {0,2,5,8} FFMA per element: %.1f %.1f %.1f %.1f elements/clk/SM\n", el / (m0 * 1e-3) / clk / sms, el / (m2 * 1e-3) / clk / sms,
           el / (m5 * 1e-3) / clk / sms, el / (m8 * 1e-3) / clk / sms);
    return 0;
}
