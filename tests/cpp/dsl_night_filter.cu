// Compiled-body path: an edge-avoiding a-trous filter in the shape of samples-public/4_Postprocessing/Night_Filter --
// iterate() over a Domain WITH HOLES (a 3 x 3 stencil dilated to 5 x 5), mask(dom) and input(dom) inside the lambda,
// packed-RGBA uint pixels, locals captured by reference, a MIRROR boundary, and an exponential evaluated by repeated
// squaring (pure arithmetic, so the result is bit-identical to the plain C loop when built with -fmad=false).  A second
// kernel uses x() / y() and output_at() / pixel_at(): it writes a vertically flipped, position-tinted copy.
//   usage: dsl_night_filter [width height] [--io in.raw out.raw]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;

#define PACK(a, b, c, d) (uint)((uint)(a) | (uint)(b) << 8 | (uint)(c) << 16 | (uint)(d) << 24)

static inline __host__ __device__ float exp_by_squaring(float v) {
    float t = 1.0f + v / 256.0f;
    for (int i = 0; i < 8; ++i) t *= t;
    return t;
}

class Atrous : public Kernel<uint> {
    Accessor<uint> &input;
    Domain &dom;
    Mask<float> &mask;

  public:
    Atrous(IterationSpace<uint> &iter, Accessor<uint> &input, Domain &dom, Mask<float> &mask) : Kernel(iter), input(input), dom(dom), mask(mask) {
        add_accessor(&input);
    }
    void kernel() {
        uint in = input();
        const float rin = (in & 0xff) / 255.0f, gin = ((in >> 8) & 0xff) / 255.0f, bin = ((in >> 16) & 0xff) / 255.0f;
        float sum_w = 0.0f, sum_r = 0.0f, sum_g = 0.0f, sum_b = 0.0f;
        iterate(dom, [&]() {
            const uint px = input(dom);
            const float r = (px & 0xff) / 255.0f, g = ((px >> 8) & 0xff) / 255.0f, b = ((px >> 16) & 0xff) / 255.0f;
            const float rd = r - rin, gd = g - gin, bd = b - bin;
            float weight = exp_by_squaring(-(rd * rd + gd * gd + bd * bd));
            if (weight > 1.0f) weight = 1.0f;
            weight *= mask(dom);
            sum_w += weight;
            sum_r += r * weight;
            sum_g += g * weight;
            sum_b += b * weight;
        });
        const float ro = sum_r / sum_w * 255.0f, go = sum_g / sum_w * 255.0f, bo = sum_b / sum_w * 255.0f;
        output() = PACK(ro, go, bo, 255);
    }
};

class FlipTint : public Kernel<uint> {
    Accessor<uint> &input;
    int height;

  public:
    FlipTint(IterationSpace<uint> &iter, Accessor<uint> &input, int height) : Kernel(iter), input(input), height(height) { add_accessor(&input); }
    void kernel() {
        const uint px = input.pixel_at(x(), height - 1 - y());
        output_at(x(), y()) = px ^ (uint)((x() & 0xff) << 8);
    }
};

static inline int mirror(int v, int n) { return tc::mirrori(v, n); }

int main(int argc, char **argv) {
    const tc::Args a(argc, argv, 700, 413);
    const int w = a.w, h = a.h;
    std::vector<uint> input((size_t)w * h);
    if (a.in) tc::read_raw(a.in, input);
    else
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) input[(size_t)y * w + x] = (uint)(tc::pixel_bits(x, y, 13) & 0xffffffffu);
    const float coef[5][5] = {{0.0625f, 0, 0.125f, 0, 0.0625f}, {0, 0, 0, 0, 0}, {0.125f, 0, 0.25f, 0, 0.125f}, {0, 0, 0, 0, 0}, {0.0625f, 0, 0.125f, 0, 0.0625f}};
    Mask<float> mask(coef);
    Domain dom(mask);   // zero coefficients are holes

    Image<uint> in(w, h, input.data()), mid(w, h), out(w, h);
    BoundaryCondition<uint> bound(in, mask, Boundary::MIRROR);
    Accessor<uint> acc(bound);
    IterationSpace<uint> is_mid(mid);
    Atrous k1(is_mid, acc, dom, mask);
    k1.execute();
    Accessor<uint> acc_mid(mid);
    IterationSpace<uint> is_out(out);
    FlipTint k2(is_out, acc_mid, h);
    k2.execute();
    uint *result = out.data();
    if (a.out) tc::write_raw(a.out, result, (size_t)w * h);

    std::vector<uint> m((size_t)w * h), ref(m.size());
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint c0 = input[(size_t)y * w + x];
            const float rin = (c0 & 0xff) / 255.0f, gin = ((c0 >> 8) & 0xff) / 255.0f, bin = ((c0 >> 16) & 0xff) / 255.0f;
            float sum_w = 0.0f, sum_r = 0.0f, sum_g = 0.0f, sum_b = 0.0f;
            for (int j = 0; j < 5; ++j)
                for (int i = 0; i < 5; ++i) {
                    if (coef[j][i] == 0.0f) continue;
                    const uint px = input[(size_t)mirror(y + j - 2, h) * w + mirror(x + i - 2, w)];
                    const float r = (px & 0xff) / 255.0f, g = ((px >> 8) & 0xff) / 255.0f, b = ((px >> 16) & 0xff) / 255.0f;
                    const float rd = r - rin, gd = g - gin, bd = b - bin;
                    float weight = exp_by_squaring(-(rd * rd + gd * gd + bd * bd));
                    if (weight > 1.0f) weight = 1.0f;
                    weight *= coef[j][i];
                    sum_w += weight; sum_r += r * weight; sum_g += g * weight; sum_b += b * weight;
                }
            const float ro = sum_r / sum_w * 255.0f, go = sum_g / sum_w * 255.0f, bo = sum_b / sum_w * 255.0f;
            m[(size_t)y * w + x] = PACK(ro, go, bo, 255);
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) ref[(size_t)y * w + x] = m[(size_t)(h - 1 - y) * w + x] ^ (uint)((x & 0xff) << 8);
    long first = -1, bad = 0;
    for (size_t i = 0; i < ref.size(); ++i)
        if (ref[i] != result[i]) { if (!bad) first = (long)i; ++bad; }
    return tc::verdict("dsl_night_filter", bad, ref.size(), first);
}
