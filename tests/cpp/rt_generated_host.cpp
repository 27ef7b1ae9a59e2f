// Host code in the shape Hipacc's rewriter EMITS (lib/Rewrite/CreateHostStrings.cpp, lib/Rewrite/Rewrite.cpp:398-420):
// hipaccInitCUDA / hipaccCreateMemory / hipaccWriteMemory / hipaccMakeAccessor / launch / hipaccReadMemory /
// reduction / pyramid -- but against include/hipacc_b200/hipacc_rt.hpp, i.e. with descriptor launches in
// place of `hipaccLaunchKernel(generatedKernel, grid, block, ...)`.  Exercises the runtime surface of SURVEY.md 8b.
#include "common.hpp"
#include "hipacc_b200/hipacc_rt.hpp"

int main() {
    hipaccInitCUDA();
    const int width = 1021, height = 517;   // deliberately not multiples of any tile
    int rc = 0;

    // ---- Laplace 3x3 on uchar with a crop accessor and CONSTANT boundary handling ------------------------
    std::vector<uchar> host_in = tc::image_u8(width, height, 7);
    HipaccImageCuda<uchar> in = hipaccCreateMemory<uchar>(nullptr, width, height);
    HipaccImageCuda<uchar> out = hipaccCreateMemory<uchar>(nullptr, width, height, 64);
    hipaccWriteMemory(in, host_in.data());
    const int ox = 5, oy = 9, rw = 900, rh = 400;
    HipaccAccessor<uchar> acc_in = hipaccMakeAccessor<uchar>(in, rw, rh, ox, oy);
    HipaccAccessor<uchar> is_out = hipaccMakeAccessor<uchar>(out, rw, rh, ox, oy);
    const int lap[9] = {0, 1, 0, 1, -4, 1, 0, 1, 0};
    hb_local_desc d;
    std::memset(&d, 0, sizeof(d));
    d.kind = HB_LOCAL_REDUCE_DOMAIN; d.reduce_mode = HB_REDUCE_SUM; d.tap = HB_TAP_MUL; d.acc_dtype = HB_S32;
    d.size_x = d.size_y = 3; d.coef_s32 = lap; d.boundary = HB_BOUNDARY_CONSTANT; d.boundary_const = 7;
    d.epilogue = HB_EPI_ADD_CLAMP_CAST; d.epi_p[0] = 128; d.epi_p[1] = 0; d.epi_p[2] = 255;
    hipaccLaunchLocalOperator(acc_in, is_out, d, nullptr, true);
    uchar *res = hipaccReadMemory(out);
    {
        long bad = 0, first = -1;
        for (int y = 0; y < rh; ++y)
            for (int x = 0; x < rw; ++x) {
                int sum = 0;
                for (int j = -1; j <= 1; ++j)
                    for (int i = -1; i <= 1; ++i) {
                        const int c = lap[(j + 1) * 3 + i + 1];
                        if (!c) continue;
                        const int xx = x + i, yy = y + j;   // region-relative; outside the accessor's window -> constant
                        const int v = (xx < 0 || xx >= rw || yy < 0 || yy >= rh) ? 7 : host_in[(size_t)(yy + oy) * width + xx + ox];
                        sum += c * v;
                    }
                sum += 128; sum = sum > 255 ? 255 : sum; sum = sum < 0 ? 0 : sum;
                if (res[(size_t)(y + oy) * width + x + ox] != (uchar)sum) { if (!bad) first = (long)y * rw + x; ++bad; }
            }
        rc |= tc::verdict("rt laplace crop CONSTANT", bad, (size_t)rw * rh, first);
    }

    // ---- copy, region copy, reductions ----------------------------------------------------------------------
    HipaccImageCuda<uchar> copy = hipaccCreateMemory<uchar>(nullptr, width, height);
    hipaccCopyMemory(in, copy);
    hipaccCopyMemoryRegion(hipaccMakeAccessor<uchar>(out, rw, rh, ox, oy), hipaccMakeAccessor<uchar>(copy, rw, rh, 0, 0));
    uchar *cp = hipaccReadMemory(copy);
    {
        long bad = 0;
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x) {
                const uchar want = (x < rw && y < rh) ? res[(size_t)(y + oy) * width + x + ox] : host_in[(size_t)y * width + x];
                bad += cp[(size_t)y * width + x] != want;
            }
        rc |= tc::verdict("rt copy + region copy", bad, (size_t)width * height, -1);
    }
    std::vector<int> host_i((size_t)width * height);
    long long want_sum = 0;
    int want_max = -1000000;
    for (size_t i = 0; i < host_i.size(); ++i) { host_i[i] = (int)host_in[i] - 100; want_sum += host_i[i]; want_max = std::max(want_max, host_i[i]); }
    HipaccImageCuda<int> img_i = hipaccCreateMemory<int>(host_i.data(), width, height);
    const int got_sum = hipaccApplyReduction<int>(img_i, HB_REDUCE_SUM), got_max = hipaccApplyReduction<int>(hipaccMakeAccessor<int>(img_i), HB_REDUCE_MAX);
    std::printf("rt reduce int: sum %d (%lld) max %d (%d): Test %s\n", got_sum, want_sum, got_max, want_max,
                (got_sum == (int)want_sum && got_max == want_max) ? "PASSED" : "FAILED");
    rc |= !(got_sum == (int)want_sum && got_max == want_max);

    // ---- pyramid creation: level 0 aliases the image, level sizes truncate ---------------------------------------
    HipaccImageCuda<float> base = hipaccCreateMemory<float>(nullptr, 101, 67);
    HipaccPyramidCuda<float> pyr = hipaccCreatePyramid<float>(base, 4);
    int visited = 0;
    bool sizes_ok = pyr.at(0).get() == base.get();
    hipaccTraverse(pyr, [&]() {
        sizes_ok = sizes_ok && pyr(0)->get_width() == (101 >> pyr.level()) && pyr(0)->get_height() == (67 >> pyr.level());
        ++visited;
        hipaccTraverse();
    });
    std::printf("rt pyramid: %d levels visited, sizes %s: Test %s\n", visited, sizes_ok ? "ok" : "wrong", (visited == 4 && sizes_ok) ? "PASSED" : "FAILED");
    rc |= !(visited == 4 && sizes_ok);

    // ---- error convention: log and continue, no fallback -----------------------------------------------------------
    d.size_x = d.size_y = 4;   // even mask: unsupported
    hipaccLaunchLocalOperator(acc_in, is_out, d);
    std::printf("rt error path: last error = \"%s\": Test %s\n", hb_last_error(), hb_last_error()[0] ? "PASSED" : "FAILED");
    rc |= !hb_last_error()[0];
    return rc;
}
