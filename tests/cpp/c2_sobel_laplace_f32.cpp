// C2 (BASELINE.json configs[1]): Sobel-X, Sobel-Y and Laplace 3x3 on a float image with MIRROR boundary
// handling -- DSL kernels of the form `output() = reduce(dom, SUM, mask(dom) * in(dom))` (zero taps are
// Domain holes and are not visited), cf. samples-public/3_Preprocessing/Sobel/src/main.cpp:55-73 and
// 1_Local_Operators/Laplace/src/main.cpp:50-72.  All pixels are compared BIT-EXACTLY against plain C loops
// (compile with -ffp-contract=off).                       usage: c2_sobel_laplace_f32 [width height]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;

class LocalFloat : public Kernel<float> {
    Accessor<float> &input;
    Domain &dom;
    Mask<float> &mask;

  public:
    LocalFloat(IterationSpace<float> &iter, Accessor<float> &input, Domain &dom, Mask<float> &mask)
        : Kernel(iter), input(input), dom(dom), mask(mask) {
        add_accessor(&input);
    }
    void kernel() override {
        output() = reduce(dom, Reduce::SUM, [&]() -> float { return mask(dom) * input(dom); });
    }
    b200::Lowering lower() override { return b200::reduce(input, dom, mask, Reduce::SUM); }
};

static void reference(const float *in, float *out, const float m[3][3], int w, int h) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float sum = 0.0f;
            bool first = true;
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) {
                    if (m[j][i] == 0.0f) continue;
                    const float v = m[j][i] * in[(size_t)tc::mirrori(y + j - 1, h) * w + tc::mirrori(x + i - 1, w)];
                    sum = first ? v : sum + v;
                    first = false;
                }
            out[(size_t)y * w + x] = sum;
        }
}

int main(int argc, char **argv) {
    const int width = argc > 2 ? std::atoi(argv[1]) : 2048, height = argc > 2 ? std::atoi(argv[2]) : 2048;
    const float sobel_x[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
    const float sobel_y[3][3] = {{-1, -2, -1}, {0, 0, 0}, {1, 2, 1}};
    const float laplace[3][3] = {{0, 1, 0}, {1, -4, 1}, {0, 1, 0}};
    const struct { const char *name; const float (*m)[3]; } ops[3] = {{"sobel_x", sobel_x}, {"sobel_y", sobel_y}, {"laplace", laplace}};
    std::vector<float> input = tc::image_f32(width, height, 2);
    Image<float> in(width, height, input.data());
    int rc = 0;
    for (const auto &op : ops) {
        float m[3][3];
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) m[j][i] = op.m[j][i];
        Mask<float> mask(m);
        Domain dom(mask);
        Image<float> out(width, height);
        BoundaryCondition<float> bound(in, mask, Boundary::MIRROR);
        Accessor<float> acc(bound);
        IterationSpace<float> iter(out);
        LocalFloat k(iter, acc, dom, mask);
        k.execute();
        std::printf("%s float %dx%d MIRROR: %.4f ms\n", op.name, width, height, hipacc_last_kernel_timing());
        float *result = out.data();
        std::vector<float> ref((size_t)width * height);
        reference(input.data(), ref.data(), m, width, height);
        long bad = 0, first = -1;
        for (size_t i = 0; i < ref.size(); ++i)
            if (std::memcmp(&result[i], &ref[i], 4) != 0 && !(result[i] == 0.0f && ref[i] == 0.0f)) { if (!bad) first = (long)i; ++bad; }
        rc |= tc::verdict(op.name, bad, ref.size(), first);
    }
    return rc;
}
