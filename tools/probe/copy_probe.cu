// copy_probe.cu -- which streaming-copy shape reaches the HBM copy peak on B200 (256 MiB in, 256 MiB out)?
// Variants: persistent vs one-shot grid, loads in flight per thread (U), ld/st cache hints, CTA size.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/copy_probe tools/probe/copy_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int U, int HINT>  // HINT 0: plain, 1: ld.cs/st.cs, 2: ld.nc (ldg) + plain st, 3: ldg + st.cs
__global__ void copy_k(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < n; base += stride * U) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const size_t i = base + (size_t)k * blockDim.x;
            if (i < n) v[k] = HINT == 1 ? __ldcs(in + i) : (HINT >= 2 ? __ldg(in + i) : in[i]);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const size_t i = base + (size_t)k * blockDim.x;
            if (i < n) { if (HINT == 1 || HINT == 3) __stcs(out + i, v[k]); else out[i] = v[k]; }
        }
    }
}
template <typename K> static void run(const char *name, K kern, int blocks, int threads, const float4 *in, float4 *out, size_t n) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<blocks, threads>>>(in, out, n);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) kern<<<blocks, threads>>>(in, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
    printf("%-44s blocks %7d x %4d : %7.1f us  %7.1f GB/s\n", name, blocks, threads, ms * 1e3, 2.0 * n * 16 / (ms * 1e-3) / 1e9);
}
int main() {
    const size_t n = (size_t)8192 * 8192 / 4;  // float4 elements
    float4 *in, *out; cudaMalloc(&in, n * 16); cudaMalloc(&out, n * 16); cudaMemset(in, 1, n * 16);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); const int sms = pr.multiProcessorCount;
    { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaMemcpy(out, in, n * 16, cudaMemcpyDeviceToDevice);
      cudaEventRecord(e0); for (int i = 0; i < 20; ++i) cudaMemcpyAsync(out, in, n * 16, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20; printf("%-44s : %7.1f us  %7.1f GB/s\n", "cudaMemcpyAsync D2D", ms * 1e3, 2.0 * n * 16 / (ms * 1e-3) / 1e9); }
    const int one4_256 = (int)((n + 256 * 4 - 1) / (256 * 4)), one1_256 = (int)((n + 255) / 256), one4_512 = (int)((n + 512 * 4 - 1) / (512 * 4));
    run("persistent 8/SM, U=4, ld.cs/st.cs", copy_k<4, 1>, sms * 8, 256, in, out, n);
    run("persistent 8/SM, U=4, plain", copy_k<4, 0>, sms * 8, 256, in, out, n);
    run("persistent 8/SM, U=4, ldg + st", copy_k<4, 2>, sms * 8, 256, in, out, n);
    run("persistent 8/SM, U=8, plain", copy_k<8, 0>, sms * 8, 256, in, out, n);
    run("persistent 4/SM x512, U=4, plain", copy_k<4, 0>, sms * 4, 512, in, out, n);
    run("persistent 16/SM x128, U=4, plain", copy_k<4, 0>, sms * 16, 128, in, out, n);
    run("one-shot, U=4, plain", copy_k<4, 0>, one4_256, 256, in, out, n);
    run("one-shot, U=4, ld.cs/st.cs", copy_k<4, 1>, one4_256, 256, in, out, n);
    run("one-shot, U=4, ldg + st.cs", copy_k<4, 3>, one4_256, 256, in, out, n);
    run("one-shot, U=1, plain", copy_k<1, 0>, one1_256, 256, in, out, n);
    run("one-shot x512, U=4, plain", copy_k<4, 0>, one4_512, 512, in, out, n);
    run("one-shot, U=2, plain", copy_k<2, 0>, (int)((n + 511) / 512), 256, in, out, n);
    run("one-shot, U=8, plain", copy_k<8, 0>, (int)((n + 2047) / 2048), 256, in, out, n);
    return 0;
}
