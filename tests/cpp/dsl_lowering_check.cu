// lower() versus kernel(): with HIPACC_B200_CHECK_LOWERING=1 execute() runs the library kernel that lower() names AND
// the device-compiled body, and aborts when they differ.
//   dsl_lowering_check good    Gaussian 5x5 uchar CLAMP with the matching lowering: passes, bit-identical
//   dsl_lowering_check wrong   the same body with a lowering that forgot the + 0.5f: must abort ("DISAGREE")
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;

class GaussianFilter : public Kernel<uchar> {
    Accessor<uchar> &input;
    Mask<float> &mask;
    const bool wrong;

  public:
    GaussianFilter(IterationSpace<uchar> &iter, Accessor<uchar> &input, Mask<float> &mask, bool wrong)
        : Kernel(iter), input(input), mask(mask), wrong(wrong) { add_accessor(&input); }
    void kernel() override {
        output() = (uchar)(convolve(mask, Reduce::SUM, [&]() -> float { return mask() * input(mask); }) + 0.5f);
    }
    b200::Lowering lower() override { return b200::convolve(input, mask, Reduce::SUM, wrong ? b200::cast() : b200::add_cast(0.5)); }
};

int main(int argc, char **argv) {
    const bool wrong = argc > 1 && !std::strcmp(argv[1], "wrong");
    setenv("HIPACC_B200_CHECK_LOWERING", "1", 1);
    const int w = 777, h = 333;
    const float coef[5][5] = {{0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.026151f, 0.090339f, 0.136565f, 0.090339f, 0.026151f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f}};
    std::vector<uchar> input = tc::image_u8(w, h, 3);
    Mask<float> mask(coef);
    Image<uchar> in(w, h, input.data()), out(w, h);
    BoundaryCondition<uchar> bound(in, mask, Boundary::CLAMP);
    Accessor<uchar> acc(bound);
    IterationSpace<uchar> iter(out);
    GaussianFilter filter(iter, acc, mask, wrong);
    filter.execute();   // aborts here when the lowering does not describe the body
    std::printf("dsl_lowering_check: lower() and kernel() agree on %d pixels: Test PASSED\n", w * h);
    return 0;
}
