// TMA probe: which (box, descriptor placement) combinations run on this GPU.  nvcc -arch=sm_100a tma_probe.cu -o tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
#include "../../hipacc_b200/csrc/hb_tma.cuh"
using namespace hb;

template <int BW, int BH>
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int use_global, int x, int y, float *out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, BW * BH * 4);
        tma_load_2d(smem, use_global ? gmap : &tmap, x, y, &bar);
    }
    mbar_wait(&bar, 0);
    const float *t = reinterpret_cast<const float *>(smem);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = t[i];
}

static int g_x = 127;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BW, int BH>
void run(EncodeTiledFn fn, float *img, int w, int h, int stride, const std::vector<float> &host, int use_global) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    cuuint64_t gstr[1] = {(cuuint64_t)stride * 4};
    cuuint32_t box[2] = {BW, BH};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, img, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d global=%d encode=%d ", BW, BH, use_global, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return; }
    CUtensorMap *gm; cudaMalloc(&gm, sizeof(m)); cudaMemcpy(gm, &m, sizeof(m), cudaMemcpyHostToDevice);
    float *out; cudaMalloc(&out, BW * BH * 4);
    const int smem = BW * BH * 4 + 128;
    cudaFuncSetAttribute(probe_kernel<BW, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int x = g_x, y = 31;
    probe_kernel<BW, BH><<<1, 256, smem>>>(m, gm, use_global, x, y, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> o(BW * BH);
        cudaMemcpy(o.data(), out, BW * BH * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) {
            const int gx = x + c, gy = y + r2;
            const float want = (gx < w && gy < h) ? host[(size_t)gy * stride + gx] : 0.0f;
            bad += o[r2 * BW + c] != want;
        }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    cudaFree(out); cudaFree(gm);
}

int main(int argc, char **argv) {
    g_x = argc > 1 ? atoi(argv[1]) : 127;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int w = 517, h = 391, stride = 576;
    std::vector<float> host((size_t)stride * h);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (float)(i % 100003);
    float *img; cudaMalloc(&img, host.size() * 4); cudaMemcpy(img, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    for (int g = 0; g < 2; ++g) {
        run<128, 32>(fn, img, w, h, stride, host, g);
        run<132, 34>(fn, img, w, h, stride, host, g);
        run<64, 34>(fn, img, w, h, stride, host, g);
        run<136, 38>(fn, img, w, h, stride, host, g);
        run<256, 16>(fn, img, w, h, stride, host, g);
    }
    return 0;
}
