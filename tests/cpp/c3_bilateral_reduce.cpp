// C3 (BASELINE.json configs[2]): bilateral filter 13x13 on a float image (the iterate() body of
// samples-public/3_Preprocessing/Bilateral_Filter/src/main.cpp:63-77 with `output() = p/d`), followed by
// three global reductions min / max / sum expressed as Kernel::reduce(left, right) and read with
// reduced_data() (samples-public/2_Global_Operators/Reduction_Sum/src/main.cpp).  Float contract: 1e-5
// relative.                                               usage: c3_bilateral_reduce [width height]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;
using namespace hipacc::math;

class BilateralFilter : public Kernel<float> {
    Accessor<float> &in;
    Mask<float> &mask;
    Domain &dom;
    int sigma_r;

  public:
    BilateralFilter(IterationSpace<float> &iter, Accessor<float> &in, Mask<float> &mask, Domain &dom, int sigma_r)
        : Kernel(iter), in(in), mask(mask), dom(dom), sigma_r(sigma_r) {
        add_accessor(&in);
    }
    void kernel() override {
        float c_r = 0.5f / (sigma_r * sigma_r);
        float d = 0.0f, p = 0.0f;
        iterate(dom, [&]() {
            float diff = in(dom) - in();
            float s = expf(-c_r * diff * diff) * mask(dom);
            d += s;
            p += s * in(dom);
        });
        output() = p / d;
    }
    b200::Lowering lower() override { return b200::bilateral(in, mask, sigma_r); }
};

template <int MODE> class Reduction : public Kernel<float> {
    Accessor<float> &in;

  public:
    Reduction(IterationSpace<float> &iter, Accessor<float> &in) : Kernel(iter), in(in) { add_accessor(&in); }
    void kernel() override { output() = in(); }
    b200::Lowering lower() override { return b200::point(HB_POINT_COPY, {&in}); }
    float reduce(float left, float right) const override { return MODE == 0 ? min(left, right) : MODE == 1 ? max(left, right) : left + right; }
};

int main(int argc, char **argv) {
    const int width = argc > 2 ? std::atoi(argv[1]) : 1024, height = argc > 2 ? std::atoi(argv[2]) : 768;
    const int S = 13, sigma_r = 16;
    // spatial Gaussian, sigma_s = S: exp(-(x^2 + y^2) / (2 * (S/3)^2))-shaped table like the sample's 13x13 mask
    float coef[13][13];
    for (int j = 0; j < S; ++j)
        for (int i = 0; i < S; ++i) {
            const float dx = (float)(i - S / 2), dy = (float)(j - S / 2);
            coef[j][i] = std::exp(-(dx * dx + dy * dy) / (2.0f * 4.0f * 4.0f));
        }
    std::vector<float> input = tc::image_f32(width, height, 3, 255.0f);

    Mask<float> mask(coef);
    Domain dom(mask);
    Image<float> in(width, height, input.data());
    Image<float> out(width, height);
    BoundaryCondition<float> bound(in, mask, Boundary::MIRROR);
    Accessor<float> acc(bound);
    IterationSpace<float> iter(out);
    BilateralFilter bf(iter, acc, mask, dom, sigma_r);
    bf.execute();
    std::printf("bilateral 13x13 float %dx%d MIRROR: %.4f ms\n", width, height, hipacc_last_kernel_timing());
    float *result = out.data();

    std::vector<float> ref((size_t)width * height);
    const float c_r = 0.5f / (sigma_r * sigma_r);
#pragma omp parallel for
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            float d = 0.0f, p = 0.0f;
            const float c = input[(size_t)y * width + x];
            for (int j = 0; j < S; ++j)
                for (int i = 0; i < S; ++i) {
                    const float v = input[(size_t)tc::mirrori(y + j - S / 2, height) * width + tc::mirrori(x + i - S / 2, width)];
                    const float diff = v - c;
                    const float s = std::exp(-c_r * diff * diff) * coef[j][i];
                    d += s;
                    p += s * v;
                }
            ref[(size_t)y * width + x] = p / d;
        }
    long first = -1;
    int rc = tc::verdict("bilateral", tc::count_diff_rel(result, ref.data(), ref.size(), 1e-5, 0.0, &first), ref.size(), first);

    // global reductions over the filtered image
    Image<float> copy(width, height);
    Accessor<float> acc_out(out);
    float got[3];
    {
        IterationSpace<float> it(copy);
        Reduction<0> rmin(it, acc_out);
        got[0] = rmin.reduced_data();
        Reduction<1> rmax(it, acc_out);
        got[1] = rmax.reduced_data();
        Reduction<2> rsum(it, acc_out);
        got[2] = rsum.reduced_data();
    }
    float mn = result[0], mx = result[0];
    double sum = 0.0;
    for (size_t i = 0; i < ref.size(); ++i) { mn = std::min(mn, result[i]); mx = std::max(mx, result[i]); sum += result[i]; }
    const bool ok = got[0] == mn && got[1] == mx && std::fabs(got[2] - sum) <= 1e-5 * std::fabs(sum);
    std::printf("reduce: min %g (%g)  max %g (%g)  sum %.9g (float64 %.9g): Test %s\n", got[0], mn, got[1], mx, got[2], sum, ok ? "PASSED" : "FAILED");
    return rc | (ok ? 0 : 1);
}
