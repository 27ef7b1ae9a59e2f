// int_probe.cu -- issue rates of the integer instructions the fused Harris kernel is built from (sm_100a):
// IDP.4A (dp4a), IDP.2A (dp2a), IMAD, IADD3, PRMT, LOP3, SHF, I2F.  Prints lane-instructions per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/int_probe tools/probe/int_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define KERNEL(NAME, BODY)                                                        \
    __global__ void NAME(int *out, int a, int b, int iters) {                      \
        int x[8];                                                                  \
        for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;                        \
        for (int it = 0; it < iters; ++it) {                                       \
            _Pragma("unroll") for (int i = 0; i < 8; ++i) { BODY; }                \
        }                                                                          \
        int s = 0;                                                                 \
        for (int i = 0; i < 8; ++i) s += x[i];                                     \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                            \
    }
KERNEL(k_dp4a, asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL(k_dp2a, asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL(k_imad, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL(k_iadd3, asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(a)))
KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL(k_shf, asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)))
KERNEL(k_i2f, { float f; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(x[i])); x[i] = __float_as_int(f); })
KERNEL(k_vabs, asm volatile("abs.s32 %0, %0;" : "+r"(x[i])))
template <typename K> static double rate(K kern, int *out, int blocks, int iters, double clk, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<blocks, 256>>>(out, 0x01020304, 3, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); kern<<<blocks, 256>>>(out, 0x01020304, 3, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return (double)blocks * 256 * 8 * iters / (ms * 1e-3) / clk / sms;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount, blocks = sms * 8, iters = 8192;
    const double clk = clk_khz * 1e3;
    int *out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    printf("%s: lane-instructions / clk / SM (nominal %d MHz)\n", pr.name, clk_khz / 1000);
    printf("IDP.4A %.1f  IDP.2A %.1f  IMAD %.1f  IADD %.1f  PRMT %.1f  LOP3 %.1f  SHF %.1f  I2F %.1f  IABS %.1f\n", rate(k_dp4a, out, blocks, iters, clk, sms),
           rate(k_dp2a, out, blocks, iters, clk, sms), rate(k_imad, out, blocks, iters, clk, sms), rate(k_iadd3, out, blocks, iters, clk, sms),
           rate(k_prmt, out, blocks, iters, clk, sms), rate(k_lop3, out, blocks, iters, clk, sms), rate(k_shf, out, blocks, iters, clk, sms),
           rate(k_i2f, out, blocks, iters, clk, sms), rate(k_vabs, out, blocks, iters, clk, sms));
    return 0;
}
