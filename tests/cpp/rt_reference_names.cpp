// Host code with ONLY the reference runtime's names, in the order and shape lib/Rewrite/CreateHostStrings.cpp:650-1190
// emits it for the CUDA target: hipacc_launch_info / dim3 block / hipaccCalcGridFromBlock / hipaccPrepareKernelLaunch /
// hipaccWriteSymbol / hipaccLaunchKernel(kernel, grid, block, ep, timing, smem, args...), hipaccApplyReductionShared,
// hipaccApplyBinningSegmented, HipaccPyramidTraversor.  The one thing a rewriter patch changes is what the kernel NAMES
// denote: hipacc_b200::OperatorKernel values (below, where the reference #includes the generated .cu files) instead of
// generated __global__ functions.  Checked against plain C loops.
#include "common.hpp"
#include "hipacc_b200/hipacc_rt.hpp"

using namespace hipacc_b200;

// ---- what stands where the generated "cuGaussianFilterKernel.cu" would be included -------------------------------------
static float _constmaskGaussianFilter[5][5];   // the reference: __device__ __constant__ float _constmask...[5][5]
static OperatorKernel make_cuGaussianFilterKernel() {
    OperatorKernel k;
    k.kind = OperatorKernel::LOCAL;
    k.local.kind = HB_LOCAL_CONVOLVE; k.local.reduce_mode = HB_REDUCE_SUM; k.local.tap = HB_TAP_MUL; k.local.acc_dtype = HB_F32;
    k.local.size_x = k.local.size_y = 5; k.local.coef_f32 = &_constmaskGaussianFilter[0][0];
    k.local.boundary = HB_BOUNDARY_CLAMP; k.local.epilogue = HB_EPI_ADD_CAST; k.local.epi_p[0] = 0.5;
    k.dtype[0] = k.dtype[1] = HB_U8;
    // (uchar *iter, int iter_width, int iter_height, int iter_stride, const uchar *input, int input_width, int input_height,
    //  int input_stride, int bh_start_left, int bh_start_right, int bh_start_top, int bh_start_bottom, int bh_fall_back)
    k.signature = {arg_ptr(0), arg_width(0), arg_height(0), arg_stride(0), arg_ptr(1), arg_width(1), arg_height(1), arg_stride(1),
                   arg_ignored(), arg_ignored(), arg_ignored(), arg_ignored(), arg_ignored()};
    return k;
}
static const OperatorKernel cuGaussianFilterKernel = make_cuGaussianFilterKernel();

// a point operator with a crop accessor and a scalar member: (out, w, h, stride, ox, oy, in1 ..., in2 ..., int norm, bh_start_right, bh_start_bottom)
static OperatorKernel make_cuSobelCombineKernel() {
    OperatorKernel k;
    k.kind = OperatorKernel::POINT;
    k.point_op = HB_POINT_SOBEL_COMBINE;
    k.dtype[0] = HB_U8; k.dtype[1] = k.dtype[2] = HB_S32;
    k.signature = {arg_ptr(0), arg_width(0), arg_height(0), arg_stride(0), arg_offset_x(0), arg_offset_y(0),
                   arg_ptr(1), arg_width(1), arg_height(1), arg_stride(1), arg_offset_x(1), arg_offset_y(1),
                   arg_ptr(2), arg_width(2), arg_height(2), arg_stride(2), arg_offset_x(2), arg_offset_y(2),
                   arg_scalar(0), arg_ignored(), arg_ignored()};
    return k;
}
static const OperatorKernel cuSobelCombineKernel = make_cuSobelCombineKernel();
static const ReductionKernel cuMaxReduce{HB_REDUCE_MAX};                       // hipacc_shared_reduction<int, cuMaxReduce>
static const BinningKernel cuHistBinning{HB_BIN_INDEX_PIXEL, HB_BIN_VALUE_ONE, 0.0};

int main() {
    hipaccInitCUDA();
    int rc = 0;
    const int width = 1500, height = 777;
    const float coef[5][5] = {{0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.026151f, 0.090339f, 0.136565f, 0.090339f, 0.026151f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f}};
    std::vector<uchar> host_in = tc::image_u8(width, height, 21);

    // ---------------- emitted for: GaussianFilter filter(iter, acc, mask); filter.execute(); -----------------------------
    HipaccImageCuda<uchar> in = hipaccCreateMemory<uchar>(NULL, width, height, 256);
    HipaccImageCuda<uchar> out = hipaccCreateMemory<uchar>(NULL, width, height, 256);
    hipaccWriteMemory(in, host_in.data());
    HipaccAccessor<uchar> acc = hipaccMakeAccessor<uchar>(in);
    HipaccAccessor<uchar> iter = hipaccMakeAccessor<uchar>(out);
    hipaccWriteSymbol<float>((const void *)&_constmaskGaussianFilter, (float *)coef, 5, 5);
    hipacc_launch_info filter_info0(2, 2, iter, 1, 1);
    dim3 block0(32, 4);
    dim3 grid0(hipaccCalcGridFromBlock(filter_info0, block0));
    hipaccPrepareKernelLaunch(filter_info0, block0);
    hipaccLaunchKernel(cuGaussianFilterKernel, grid0, block0, HipaccExecutionParameterCuda{}, false, 0, out->get_device_memory(), iter.width,
                       iter.height, out->get_stride(), in->get_device_memory(), acc.width, acc.height, in->get_stride(), filter_info0.bh_start_left,
                       filter_info0.bh_start_right, filter_info0.bh_start_top, filter_info0.bh_start_bottom, filter_info0.bh_fall_back);
    uchar *res = hipaccReadMemory(out);
    {
        long bad = 0, first = -1;
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x) {
                float sum = 0.0f;
                for (int j = 0; j < 5; ++j)
                    for (int i = 0; i < 5; ++i) {
                        const float v = coef[j][i] * (float)host_in[(size_t)tc::clampi(y + j - 2, 0, height - 1) * width + tc::clampi(x + i - 2, 0, width - 1)];
                        sum = (i == 0 && j == 0) ? v : sum + v;
                    }
                if (res[(size_t)y * width + x] != (uchar)(sum + 0.5f)) { if (!bad) first = (long)y * width + x; ++bad; }
            }
        rc |= tc::verdict("hipaccLaunchKernel [Gaussian 5x5, hipaccWriteSymbol mask]", bad, (size_t)width * height, first);
        // grid / launch-info arithmetic of hipacc_cu_standalone.hpp:66-110
        const bool info_ok = grid0.x == (unsigned)((width + 31) / 32) && grid0.y == (unsigned)((height + 3) / 4) && filter_info0.bh_start_left == 1 &&
                             filter_info0.bh_start_right == (width - 2) / 32 && filter_info0.bh_start_top == 1 && filter_info0.bh_fall_back == 0;
        std::printf("hipacc_launch_info: grid %ux%u, bh left %d right %d top %d bottom %d fall_back %d: Test %s\n", grid0.x, grid0.y, filter_info0.bh_start_left,
                    filter_info0.bh_start_right, filter_info0.bh_start_top, filter_info0.bh_start_bottom, filter_info0.bh_fall_back, info_ok ? "PASSED" : "FAILED");
        rc |= !info_ok;
    }

    // ---------------- a two-input point operator on a crop region, scalar member `norm` -------------------------------------
    std::vector<int> gx((size_t)width * height), gy(gx.size());
    for (size_t i = 0; i < gx.size(); ++i) { gx[i] = (int)(tc::splitmix64(i) % 2041) - 1020; gy[i] = (int)(tc::splitmix64(i + 77) % 2041) - 1020; }
    HipaccImageCuda<int> img_gx = hipaccCreateMemory<int>(gx.data(), width, height, 256), img_gy = hipaccCreateMemory<int>(gy.data(), width, height, 256);
    HipaccImageCuda<uchar> mag = hipaccCreateMemory<uchar>(NULL, width, height, 256);
    const int ox = 8, oy = 3, rw = 1200, rh = 700, norm = 6;
    HipaccAccessor<int> a1 = hipaccMakeAccessor<int>(img_gx, rw, rh, ox, oy), a2 = hipaccMakeAccessor<int>(img_gy, rw, rh, ox, oy);
    HipaccAccessor<uchar> is2 = hipaccMakeAccessor<uchar>(mag, rw, rh, ox, oy);
    hipacc_launch_info combine_info1(0, 0, is2, 1, 1);
    dim3 block1(128, 1);
    dim3 grid1(hipaccCalcGridFromBlock(combine_info1, block1));
    hipaccPrepareKernelLaunch(combine_info1, block1);
    hipaccLaunchKernel(cuSobelCombineKernel, grid1, block1, HipaccExecutionParameterCuda{}, false, 0, mag->get_device_memory(), is2.width, is2.height,
                       mag->get_stride(), is2.offset_x, is2.offset_y, img_gx->get_device_memory(), a1.width, a1.height, img_gx->get_stride(), a1.offset_x,
                       a1.offset_y, img_gy->get_device_memory(), a2.width, a2.height, img_gy->get_stride(), a2.offset_x, a2.offset_y, norm,
                       combine_info1.bh_start_right, combine_info1.bh_start_bottom);
    uchar *m = hipaccReadMemory(mag);
    {
        long bad = 0, first = -1;
        for (int y = 0; y < rh; ++y)
            for (int x = 0; x < rw; ++x) {
                const size_t i = (size_t)(y + oy) * width + x + ox;
                const int i1 = gx[i] / norm, i2 = gy[i] / norm;   // samples-public/3_Preprocessing/Sobel/src/main.cpp:88-96
                float r = sqrtf((float)(i1 * i1 + i2 * i2));
                r = r < 255.0f ? r : 255.0f;
                r = r > 0.0f ? r : 0.0f;
                if (m[i] != (uchar)r) { if (!bad) first = (long)i; ++bad; }
            }
        rc |= tc::verdict("hipaccLaunchKernel [SobelCombine, crop accessors, scalar member]", bad, (size_t)rw * rh, first);
    }

    // ---------------- reduction and binning with the reference's entry points ----------------------------------------------------
    int want_max = -100000;
    for (int v : gx) want_max = std::max(want_max, v);
    const int got_max = hipaccApplyReductionShared<int>(cuMaxReduce, hipaccMakeAccessor<int>(img_gx), 128, 16, HipaccExecutionParameterCuda{}, NULL);
    std::printf("hipaccApplyReductionShared: max %d (%d): Test %s\n", got_max, want_max, got_max == want_max ? "PASSED" : "FAILED");
    rc |= got_max != want_max;
    uint *bins = hipaccApplyBinningSegmented<uint, uchar>(cuHistBinning, acc, 16, 16, 256, HipaccExecutionParameterCuda{}, NULL, false);
    std::vector<uint> want_bins(256, 0);
    for (uchar v : host_in) ++want_bins[v];
    long bad_bins = 0;
    for (int b = 0; b < 256; ++b) bad_bins += bins[b] != want_bins[b];
    delete[] bins;
    rc |= tc::verdict("hipaccApplyBinningSegmented [256-bin histogram]", bad_bins, 256, -1);

    // ---------------- HipaccPyramidTraversor ------------------------------------------------------------------------------------------
    HipaccImageCuda<float> base = hipaccCreateMemory<float>(NULL, 64, 48);
    HipaccPyramidCuda<float> pyr = hipaccCreatePyramid<float>(base, 3);
    HipaccPyramidTraversor traversor;
    std::vector<int> order;
    traversor.hipaccTraverse(pyr, [&]() {
        order.push_back(pyr.level());
        traversor.hipaccTraverse(2, [&]() { order.push_back(-1); });
        order.push_back(10 + pyr.level());
    });
    // level 0 enters, level 1 runs twice (each time level 2 runs twice) with the in-between function, then everything unwinds
    const std::vector<int> want_order = {0, 1, 2, 12, -1, 2, 12, 11, -1, 1, 2, 12, -1, 2, 12, 11, 10};
    std::printf("HipaccPyramidTraversor: %zu steps: Test %s\n", order.size(), order == want_order ? "PASSED" : "FAILED");
    rc |= order != want_order;
    return rc;
}
