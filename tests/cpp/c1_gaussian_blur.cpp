// C1 (BASELINE.json configs[0]): Gaussian blur 5x5, uchar, CLAMP -- a Hipacc DSL program in the shape of
// samples-public/1_Local_Operators/Gaussian_Blur/src/main.cpp, run on the B200 through the DSL front and
// checked on ALL pixels (borders included) against a plain C loop.   usage: c1_gaussian_blur [width height]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;
using namespace hipacc::math;

class GaussianFilter : public Kernel<uchar> {
    Accessor<uchar> &input;
    Mask<float> &mask;

  public:
    GaussianFilter(IterationSpace<uchar> &iter, Accessor<uchar> &input, Mask<float> &mask) : Kernel(iter), input(input), mask(mask) {
        add_accessor(&input);
    }
    void kernel() override {
        output() = (uchar)(convolve(mask, Reduce::SUM, [&]() -> float { return mask() * input(mask); }) + 0.5f);
    }
    b200::Lowering lower() override { return b200::convolve(input, mask, Reduce::SUM, b200::add_cast(0.5)); }
};

static void gaussian_reference(const uchar *in, uchar *out, const float *m, int s, int w, int h) {
    const int r = s / 2;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float sum = 0.0f;
            for (int j = 0; j < s; ++j)
                for (int i = 0; i < s; ++i)
                    sum += m[j * s + i] * (float)in[(size_t)tc::clampi(y + j - r, 0, h - 1) * w + tc::clampi(x + i - r, 0, w - 1)];
            out[(size_t)y * w + x] = (uchar)(sum + 0.5f);
        }
}

int main(int argc, char **argv) {
    const int width = argc > 2 ? std::atoi(argv[1]) : 4096, height = argc > 2 ? std::atoi(argv[2]) : 4096;
    const float coef[5][5] = {{0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.026151f, 0.090339f, 0.136565f, 0.090339f, 0.026151f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f}};
    std::vector<uchar> input = tc::image_u8(width, height, 1);

    Mask<float> mask(coef);
    Image<uchar> in(width, height, input.data());
    Image<uchar> out(width, height);
    BoundaryCondition<uchar> bound(in, mask, Boundary::CLAMP);
    Accessor<uchar> acc(bound);
    IterationSpace<uchar> iter(out);
    GaussianFilter filter(iter, acc, mask);
    filter.execute();
    const float ms = hipacc_last_kernel_timing();
    filter.execute();  // second call is a no-op (dsl/kernel.hpp:95)
    uchar *result = out.data();
    std::printf("Hipacc-B200 Gaussian 5x5 uchar %dx%d CLAMP: %.4f ms, %.1f Mpixel/s\n", width, height, ms, width * (double)height / ms / 1000.0);

    std::vector<uchar> ref((size_t)width * height);
    gaussian_reference(input.data(), ref.data(), &coef[0][0], 5, width, height);
    long first = -1;
    const long bad = tc::count_diff(result, ref.data(), ref.size(), 0, &first);
    return tc::verdict("c1_gaussian_blur", bad, ref.size(), first);
}
