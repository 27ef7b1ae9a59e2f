// Compiled-body path: an unsharp-masking pipeline of FIVE different kernel() bodies in the shape of
// samples-public/1_Local_Operators/Unsharp/src/main.cpp -- uchar4 -> float luminance (powf), float Gaussian via
// convolve() under a CLAMP boundary, two two-input point operators, and a mixed uchar4 / float operator that scales
// channels in place.  None of the classes has a lower(); every body is compiled for the device by nvcc.
// Checked against plain C loops: float stages within 1e-5 relative (device powf vs libm), the uchar4 result within 1 LSB
// (the tolerance of the sample's own comparison, samples-public/common/hipacc_helper.hpp:180-207).
//   usage: dsl_unsharp [width height] [--io in.raw out.raw]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;
using namespace hipacc::math;

class RGB2Gray : public Kernel<float> {
    Accessor<uchar4> &input;

  public:
    RGB2Gray(IterationSpace<float> &iter, Accessor<uchar4> &input) : Kernel(iter), input(input) { add_accessor(&input); }
    void kernel() {
        const uchar4 pixel = input();
        const float c = 1.0f / 2.2f;
        output() = 0.2126f * powf((float)pixel.x, c) + 0.7152f * powf((float)pixel.y, c) + 0.0722f * powf((float)pixel.z, c);
    }
};

class GaussianBlur : public Kernel<float> {
    Accessor<float> &input;
    Mask<float> &mask;

  public:
    GaussianBlur(IterationSpace<float> &iter, Accessor<float> &input, Mask<float> &mask) : Kernel(iter), input(input), mask(mask) { add_accessor(&input); }
    void kernel() {
        output() = convolve(mask, Reduce::SUM, [&]() -> float { return mask() * input(mask); });
    }
};

class Sharpen : public Kernel<float> {
    Accessor<float> &gray, &blurred;

  public:
    Sharpen(IterationSpace<float> &is, Accessor<float> &gray, Accessor<float> &blurred) : Kernel(is), gray(gray), blurred(blurred) {
        add_accessor(&gray);
        add_accessor(&blurred);
    }
    void kernel() { output() = 2 * gray() - blurred(); }
};

class Ratio : public Kernel<float> {
    Accessor<float> &gray, &sharp;

  public:
    Ratio(IterationSpace<float> &is, Accessor<float> &gray, Accessor<float> &sharp) : Kernel(is), gray(gray), sharp(sharp) {
        add_accessor(&gray);
        add_accessor(&sharp);
    }
    void kernel() {
        float pixel = gray();
        pixel = max(pixel, 0.01f);
        output() = sharp() / pixel;
    }
};

class Unsharp : public Kernel<uchar4> {
    Accessor<uchar4> &input;
    Accessor<float> &ratio;

  public:
    Unsharp(IterationSpace<uchar4> &is, Accessor<uchar4> &input, Accessor<float> &ratio) : Kernel(is), input(input), ratio(ratio) {
        add_accessor(&input);
        add_accessor(&ratio);
    }
    void kernel() {
        uchar4 in = input();
        const float r = ratio();
        in.x *= r;
        in.y *= r;
        in.z *= r;
        output() = in;
    }
};

int main(int argc, char **argv) {
    const tc::Args a(argc, argv, 1030, 517);
    const int w = a.w, h = a.h;
    std::vector<uchar4> input((size_t)w * h);
    if (a.in) tc::read_raw(a.in, input);
    else {
        // channel values in [50, 100]: every `in.x *= r` stays inside uchar.  An out-of-range float -> uchar conversion is
        // undefined in C++ and differs between x86 (wraps through int) and the GPU (saturates) in the reference as well.
        const std::vector<unsigned char> raw = tc::image_u8(w * 4, h, 12);
        for (size_t i = 0; i < input.size(); ++i)
            input[i] = make_uchar4((uchar)(50 + raw[4 * i] % 51), (uchar)(50 + raw[4 * i + 1] % 51), (uchar)(50 + raw[4 * i + 2] % 51), raw[4 * i + 3]);
    }
    const float coef[3][3] = {{0.057118f, 0.124758f, 0.057118f}, {0.124758f, 0.272496f, 0.124758f}, {0.057118f, 0.124758f, 0.057118f}};
    Mask<float> mask(coef);

    Image<uchar4> in(w, h, input.data());
    Image<uchar4> out(w, h);
    Image<float> gray(w, h), blur(w, h), sharp(w, h), ratio(w, h);

    Accessor<uchar4> acc_in(in);
    IterationSpace<float> is_gray(gray);
    RGB2Gray k_gray(is_gray, acc_in);
    k_gray.execute();

    BoundaryCondition<float> bound(gray, mask, Boundary::CLAMP);
    Accessor<float> acc_gray_bc(bound);
    IterationSpace<float> is_blur(blur);
    GaussianBlur k_blur(is_blur, acc_gray_bc, mask);
    k_blur.execute();

    Accessor<float> acc_gray(gray), acc_blur(blur);
    IterationSpace<float> is_sharp(sharp);
    Sharpen k_sharp(is_sharp, acc_gray, acc_blur);
    k_sharp.execute();

    Accessor<float> acc_sharp(sharp);
    IterationSpace<float> is_ratio(ratio);
    Ratio k_ratio(is_ratio, acc_gray, acc_sharp);
    k_ratio.execute();

    Accessor<float> acc_ratio(ratio);
    IterationSpace<uchar4> is_out(out);
    Unsharp k_out(is_out, acc_in, acc_ratio);
    k_out.execute();

    uchar4 *result = out.data();
    if (a.out) tc::write_raw(a.out, result, (size_t)w * h);
    std::vector<float> dev_ratio(ratio.data(), ratio.data() + (size_t)w * h);

    // plain C reference of the same pipeline
    std::vector<float> g((size_t)w * h), b(g.size()), rt(g.size());
    const float c = 1.0f / 2.2f;
    for (size_t i = 0; i < g.size(); ++i)
        g[i] = 0.2126f * powf((float)input[i].x, c) + 0.7152f * powf((float)input[i].y, c) + 0.0722f * powf((float)input[i].z, c);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float s = 0.0f;
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) {
                    const float v = coef[j][i] * g[(size_t)tc::clampi(y + j - 1, 0, h - 1) * w + tc::clampi(x + i - 1, 0, w - 1)];
                    s = (i == 0 && j == 0) ? v : s + v;
                }
            b[(size_t)y * w + x] = s;
        }
    std::vector<uchar4> ref(g.size());
    for (size_t i = 0; i < g.size(); ++i) {
        const float sh = 2 * g[i] - b[i];
        float p = g[i];
        p = p > 0.01f ? p : 0.01f;
        rt[i] = sh / p;
        uchar4 o = input[i];
        o.x *= rt[i]; o.y *= rt[i]; o.z *= rt[i];
        ref[i] = o;
    }
    long first = -1;
    long bad = tc::count_diff_rel(dev_ratio.data(), rt.data(), rt.size(), 1e-5, 1e-6, &first);
    int rc = tc::verdict("dsl_unsharp (ratio image, float)", bad, rt.size(), first);
    bad = tc::count_diff(reinterpret_cast<const uchar *>(result), reinterpret_cast<const uchar *>(ref.data()), ref.size() * 4, 1, &first);
    rc |= tc::verdict("dsl_unsharp (uchar4 result)", bad, ref.size() * 4, first);
    return rc;
}
