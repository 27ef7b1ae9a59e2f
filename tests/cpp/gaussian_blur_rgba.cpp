// Vector pixel types (samples-public/1_Local_Operators/Gaussian_Blur_RGBA/src/main.cpp): Gaussian blur 5x5 on a
// uchar4 image, float4 accumulate, convert_uchar4(sum + 0.5f), MIRROR boundary.  The DSL's vector arithmetic is
// element-wise, so the check is the plain C loop per channel over ALL pixels.   usage: gaussian_blur_rgba [width height]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;
using namespace hipacc::math;

class GaussianBlur : public Kernel<uchar4> {
    Accessor<uchar4> &input;
    Mask<float> &mask;

  public:
    GaussianBlur(IterationSpace<uchar4> &iter, Accessor<uchar4> &input, Mask<float> &mask) : Kernel(iter), input(input), mask(mask) {
        add_accessor(&input);
    }
    void kernel() override {}   // float4 sum = convolve(mask, Reduce::SUM, [&]{ return mask() * convert_float4(input(mask)); }); output() = convert_uchar4(sum + 0.5f);
    b200::Lowering lower() override { return b200::convolve(input, mask, Reduce::SUM, b200::add_cast(0.5), HB_F32); }
};

int main(int argc, char **argv) {
    const int width = argc > 2 ? std::atoi(argv[1]) : 4032, height = argc > 2 ? std::atoi(argv[2]) : 3024;
    const float coef[5][5] = {{0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.026151f, 0.090339f, 0.136565f, 0.090339f, 0.026151f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f}};
    std::vector<unsigned char> bytes = tc::image_u8(width * 4, height, 31);   // interleaved RGBA

    Mask<float> mask(coef);
    Image<uchar4> in(width, height, reinterpret_cast<uchar4 *>(bytes.data()));
    Image<uchar4> out(width, height);
    BoundaryCondition<uchar4> bound(in, mask, Boundary::MIRROR);
    Accessor<uchar4> acc(bound);
    IterationSpace<uchar4> iter(out);
    GaussianBlur filter(iter, acc, mask);
    filter.execute();
    std::printf("Gaussian 5x5 uchar4 %dx%d MIRROR: %.4f ms\n", width, height, hipacc_last_kernel_timing());
    const unsigned char *result = reinterpret_cast<const unsigned char *>(out.data());

    std::vector<unsigned char> ref((size_t)width * height * 4);
#pragma omp parallel for
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x)
            for (int c = 0; c < 4; ++c) {
                float sum = 0.0f;
                for (int j = 0; j < 5; ++j)
                    for (int i = 0; i < 5; ++i) {
                        const float v = (float)bytes[((size_t)tc::mirrori(y + j - 2, height) * width + tc::mirrori(x + i - 2, width)) * 4 + c];
                        const float t = coef[j][i] * v;
                        sum = (j == 0 && i == 0) ? t : sum + t;
                    }
                ref[((size_t)y * width + x) * 4 + c] = (unsigned char)(sum + 0.5f);
            }
    long first = -1;
    return tc::verdict("gaussian_rgba", tc::count_diff(result, ref.data(), ref.size(), 0, &first), ref.size(), first);
}
