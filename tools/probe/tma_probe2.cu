// TMA probe 2: the CUDA programming guide's own TMA example (libcu++ cuda::barrier + cde::cp_async_bulk_tensor_2d_global_to_shared)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW = 128, BH = 32;

__global__ void kern(const __grid_constant__ CUtensorMap tensor_map, int x, int y, float *out) {
    __shared__ alignas(128) float smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (variant == 1) cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    else cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int w = variant == 2 ? 512 : 517, h = variant == 2 ? 384 : 391, stride = 576;
    std::vector<float> host((size_t)stride * h);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (float)(i % 100003);
    float *img; cudaMalloc(&img, host.size() * 4); cudaMemcpy(img, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    cuuint64_t gstr[1] = {(cuuint64_t)stride * 4};
    cuuint32_t box[2] = {BW, BH};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, img, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    variant == 3 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode=%d ", variant, (int)r);
    const unsigned long long *mw = reinterpret_cast<const unsigned long long *>(&m);
    float *out; cudaMalloc(&out, BW * BH * 4);
    kern<<<1, 256>>>(m, 64, 16, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s\n", cudaGetErrorString(e));
    for (int i = 0; i < 16; ++i) printf("%016llx ", mw[i]);
    printf("\n");
    return 0;
}
