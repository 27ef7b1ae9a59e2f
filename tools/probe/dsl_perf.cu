// compiled-body path vs lowered path on the same operator (Gaussian 5x5 uchar CLAMP 4096^2, Sobel-X 3x3 float MIRROR 8192^2):
// each Kernel object executes once (dsl/kernel.hpp:95), so the timed launch is the second object's -- the first one pays
// CUDA's lazy module load.   build: nvcc -x cu -std=c++17 -O2 -fmad=false -gencode arch=compute_100a,code=sm_100a -I include ...
#include <cstdio>
#include <vector>
#include "hipacc_b200/hipacc.hpp"
using namespace hipacc;

class GaussBody : public Kernel<uchar> {
    Accessor<uchar> &input; Mask<float> &mask;
  public:
    GaussBody(IterationSpace<uchar> &iter, Accessor<uchar> &input, Mask<float> &mask) : Kernel(iter), input(input), mask(mask) { add_accessor(&input); }
    void kernel() { output() = (uchar)(convolve(mask, Reduce::SUM, [&]() -> float { return mask() * input(mask); }) + 0.5f); }
};
class GaussLowered : public Kernel<uchar> {
    Accessor<uchar> &input; Mask<float> &mask;
  public:
    GaussLowered(IterationSpace<uchar> &iter, Accessor<uchar> &input, Mask<float> &mask) : Kernel(iter), input(input), mask(mask) { add_accessor(&input); }
    void kernel() { output() = (uchar)(convolve(mask, Reduce::SUM, [&]() -> float { return mask() * input(mask); }) + 0.5f); }
    b200::Lowering lower() override { return b200::convolve(input, mask, Reduce::SUM, b200::add_cast(0.5)); }
};
class SobelBody : public Kernel<float> {
    Accessor<float> &input; Domain &dom; Mask<float> &mask;
  public:
    SobelBody(IterationSpace<float> &iter, Accessor<float> &input, Domain &dom, Mask<float> &mask) : Kernel(iter), input(input), dom(dom), mask(mask) { add_accessor(&input); }
    void kernel() { output() = reduce(dom, Reduce::SUM, [&]() -> float { return mask(dom) * input(dom); }); }
};

int main() {
    const float coef[5][5] = {{0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f}, {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.026151f, 0.090339f, 0.136565f, 0.090339f, 0.026151f}, {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f}};
    const float sob[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
    {
        const int n = 4096;
        std::vector<uchar> h((size_t)n * n, 7);
        Mask<float> mask(coef);
        Image<uchar> in(n, n, h.data()), out(n, n);
        BoundaryCondition<uchar> bc(in, mask, Boundary::CLAMP);
        Accessor<uchar> acc(bc);
        IterationSpace<uchar> is(out);
        float ms[2][2];
        for (int rep = 0; rep < 2; ++rep) {
            GaussBody a(is, acc, mask); a.execute(); ms[0][rep] = hipacc_last_kernel_timing();
            GaussLowered b(is, acc, mask); b.execute(); ms[1][rep] = hipacc_last_kernel_timing();
        }
        std::printf("Gaussian 5x5 uchar CLAMP %d^2: compiled body %.3f ms = %.1f Gpx/s, lowered %.3f ms = %.1f Gpx/s\n", n, ms[0][1], n * (double)n / ms[0][1] / 1e6,
                    ms[1][1], n * (double)n / ms[1][1] / 1e6);
    }
    {
        const int n = 8192;
        std::vector<float> h((size_t)n * n, 1.5f);
        Mask<float> mask(sob);
        Domain dom(mask);
        Image<float> in(n, n, h.data()), out(n, n);
        BoundaryCondition<float> bc(in, mask, Boundary::MIRROR);
        Accessor<float> acc(bc);
        IterationSpace<float> is(out);
        float ms[2];
        for (int rep = 0; rep < 2; ++rep) { SobelBody a(is, acc, dom, mask); a.execute(); ms[rep] = hipacc_last_kernel_timing(); }
        std::printf("Sobel-X 3x3 float MIRROR %d^2 (Domain holes): compiled body %.3f ms = %.1f Gpx/s\n", n, ms[1], n * (double)n / ms[1] / 1e6);
    }
    return 0;
}
