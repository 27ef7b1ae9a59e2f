// C5 (BASELINE.json configs[4]): Gaussian / Laplacian pyramid traverse on a float image -- the program of
// samples-public/5_Other/Gaussian_Laplacian_Pyramid/src/main.cpp:199-248 with float pixels: on the way down
// Gaussian (CLAMP) -> NN subsample -> difference with the LF-upsampled coarse level; on the way up Restore
// and Blend.  Checked against a plain C restatement of the same traversal, and with the sample's own
// property (restored input == input).                      usage: c5_pyramid [width height depth]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;

class Gaussian : public Kernel<float> {
    Accessor<float> &input; Mask<float> &mask;
  public:
    Gaussian(IterationSpace<float> &it, Accessor<float> &in, Mask<float> &mask) : Kernel(it), input(in), mask(mask) { add_accessor(&input); }
    void kernel() override { output() = convolve(mask, Reduce::SUM, [&]() { return input(mask) * mask(); }); }
    b200::Lowering lower() override { return b200::convolve(input, mask, Reduce::SUM); }
};
class Subsample : public Kernel<float> {
    Accessor<float> &input;
  public:
    Subsample(IterationSpace<float> &it, Accessor<float> &in) : Kernel(it), input(in) { add_accessor(&input); }
    void kernel() override { output() = input(); }
    b200::Lowering lower() override { return b200::point(HB_POINT_COPY, {&input}); }
};
template <int OP> class Combine : public Kernel<float> {   // DifferenceOfGaussian (SUB), Restore (ADD), Blend
    Accessor<float> &a, &b;
  public:
    Combine(IterationSpace<float> &it, Accessor<float> &a, Accessor<float> &b) : Kernel(it), a(a), b(b) { add_accessor(&a); add_accessor(&b); }
    void kernel() override { output() = OP == HB_POINT_SUB ? a() - b() : OP == HB_POINT_ADD ? a() + b() : a() + b() / 2; }
    b200::Lowering lower() override { return b200::point(OP, {&a, &b}); }
};

// ---- plain C restatement -------------------------------------------------------------------------
struct Plane { int w, h; std::vector<float> p; float at(int x, int y) const { return p[(size_t)tc::clampi(y, 0, h - 1) * w + tc::clampi(x, 0, w - 1)]; } };
static float lf(const Plane &c, int gx, int gy, int is_w, int is_h) {   // dsl/image.hpp:390-422 with the default CLAMP
    const float sx = (float)c.w / (float)is_w, sy = (float)c.h / (float)is_h;
    float xb = (0.0f + sx / 2.0f + sx * (float)gx) - 0.5f, yb = (0.0f + sy / 2.0f + sy * (float)gy) - 0.5f;
    if (xb < 0.0f) xb = 0.0f;
    if (yb < 0.0f) yb = 0.0f;
    const int xi = (int)xb, yi = (int)yb;
    const float xf = xb - (float)xi, yf = yb - (float)yi;
    return (1.0f - xf) * (1.0f - yf) * c.at(xi, yi) + xf * (1.0f - yf) * c.at(xi + 1, yi) + (1.0f - xf) * yf * c.at(xi, yi + 1) + xf * yf * c.at(xi + 1, yi + 1);
}
static void reference(std::vector<Plane> &g, std::vector<Plane> &l, const float *m, int s) {
    const int depth = (int)g.size();
    for (int lv = 1; lv < depth; ++lv) {
        const Plane &f = g[lv - 1];
        Plane tmp{f.w, f.h, std::vector<float>(f.p.size())};
        for (int y = 0; y < f.h; ++y)
            for (int x = 0; x < f.w; ++x) {
                float sum = 0.0f;
                for (int j = 0; j < s; ++j)
                    for (int i = 0; i < s; ++i) sum += f.at(x + i - s / 2, y + j - s / 2) * m[j * s + i];
                tmp.p[(size_t)y * f.w + x] = sum;
            }
        Plane &c = g[lv];
        const float sx = (float)f.w / (float)c.w, sy = (float)f.h / (float)c.h;
        for (int y = 0; y < c.h; ++y)
            for (int x = 0; x < c.w; ++x) c.p[(size_t)y * c.w + x] = tmp.at((int)(sx / 2.0f + sx * (float)x), (int)(sy / 2.0f + sy * (float)y));
        for (int y = 0; y < f.h; ++y)
            for (int x = 0; x < f.w; ++x) l[lv - 1].p[(size_t)y * f.w + x] = f.p[(size_t)y * f.w + x] - lf(c, x, y, f.w, f.h);
    }
    for (int lv = depth - 2; lv >= 0; --lv) {
        Plane &f = g[lv];
        for (int y = 0; y < f.h; ++y)
            for (int x = 0; x < f.w; ++x) {
                const size_t i = (size_t)y * f.w + x;
                f.p[i] = lf(g[lv + 1], x, y, f.w, f.h) + l[lv].p[i];
                l[lv].p[i] = lf(l[lv + 1], x, y, f.w, f.h) + l[lv].p[i] / 2;
            }
    }
}

int main(int argc, char **argv) {
    const int width = argc > 3 ? std::atoi(argv[1]) : 1000, height = argc > 3 ? std::atoi(argv[2]) : 744, depth = argc > 3 ? std::atoi(argv[3]) : 5;
    const float coef[5][5] = {{0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.026151f, 0.090339f, 0.136565f, 0.090339f, 0.026151f},
                              {0.017300f, 0.059761f, 0.090339f, 0.059761f, 0.017300f},
                              {0.005008f, 0.017300f, 0.026151f, 0.017300f, 0.005008f}};
    std::vector<float> input = tc::image_f32(width, height, 5);

    Image<float> gaus(width, height, input.data()), tmp(width, height), lap(width, height);
    Mask<float> mask(coef);
    Pyramid<float> pgaus(gaus, depth), ptmp(tmp, depth), plap(lap, depth);
    float timing = 0.0f;
    traverse(pgaus, ptmp, plap, [&]() {
        if (!pgaus.is_top_level()) {
            BoundaryCondition<float> bound(pgaus(-1), mask, Boundary::CLAMP);
            Accessor<float> acc1(bound);
            IterationSpace<float> iter1(ptmp(-1));
            Gaussian blur(iter1, acc1, mask);
            blur.execute(); timing += hipacc_last_kernel_timing();
            Accessor<float> acc2(ptmp(-1), Interpolate::NN);
            IterationSpace<float> iter2(pgaus(0));
            Subsample sub(iter2, acc2);
            sub.execute(); timing += hipacc_last_kernel_timing();
            Accessor<float> acc3(pgaus(-1)), acc4(pgaus(0), Interpolate::LF);
            IterationSpace<float> iter3(plap(-1));
            Combine<HB_POINT_SUB> dog(iter3, acc3, acc4);
            dog.execute(); timing += hipacc_last_kernel_timing();
        }
        traverse();
        if (!pgaus.is_bottom_level()) {
            Accessor<float> acc1(pgaus(1), Interpolate::LF), acc2(plap(0));
            IterationSpace<float> iter1(pgaus(0));
            Combine<HB_POINT_ADD> res(iter1, acc1, acc2);
            res.execute(); timing += hipacc_last_kernel_timing();
            Accessor<float> acc3(plap(1), Interpolate::LF), acc4(plap(0));
            IterationSpace<float> iter2(plap(0));
            Combine<HB_POINT_BLEND> blend(iter2, acc3, acc4);
            blend.execute(); timing += hipacc_last_kernel_timing();
        }
    });
    std::printf("pyramid traverse float %dx%d depth %d: %.4f ms (sum of kernel timings)\n", width, height, depth, timing);
    float *restored = gaus.data(), *output = lap.data();

    std::vector<Plane> g, l;
    for (int lv = 0, w = width, h = height; lv < depth; ++lv, w /= 2, h /= 2) {
        g.push_back(Plane{w, h, std::vector<float>((size_t)w * h)});
        l.push_back(Plane{w, h, std::vector<float>((size_t)w * h)});
    }
    g[0].p = input;
    reference(g, l, &coef[0][0], 5);
    long first = -1;
    int rc = tc::verdict("restored gaus(0) vs C", tc::count_diff_rel(restored, g[0].p.data(), input.size(), 1e-5, 1e-6, &first), input.size(), first);
    rc |= tc::verdict("blended lap(0) vs C", tc::count_diff_rel(output, l[0].p.data(), input.size(), 1e-5, 1e-6, &first), input.size(), first);
    rc |= tc::verdict("restored == input", tc::count_diff_rel(restored, input.data(), input.size(), 1e-5, 1e-6, &first), input.size(), first);
    return rc;
}
