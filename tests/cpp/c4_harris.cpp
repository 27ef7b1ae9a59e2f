// C4 (BASELINE.json configs[3]): Harris corner detector on a uchar image -- the nine-kernel DSL pipeline of
// samples-public/3_Preprocessing/Harris_Corner/src/main.cpp:230-305 (Sobel dx/dy -> squares -> 3x3 binomial ->
// response, CLAMP everywhere) run kernel by kernel through the DSL front, the fused single-kernel operator
// (b200::harris), and a plain C restatement.  All three must agree bit for bit.   usage: c4_harris [width height]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;

class Deriv : public Kernel<short> {   // uchar -> short: (short)sum / 6 over the non-zero taps
    Accessor<uchar> &input; Domain &dom; Mask<int> &mask;
  public:
    Deriv(IterationSpace<short> &it, Accessor<uchar> &in, Domain &dom, Mask<int> &mask) : Kernel(it), input(in), dom(dom), mask(mask) { add_accessor(&input); }
    void kernel() override {
        short sum = 0;
        sum += reduce(dom, Reduce::SUM, [&]() -> short { return input(dom) * mask(dom); });
        output() = sum / 6;
    }
    // the generic local operator reads uchar and writes short through hb_local_op's (u8 -> s16) instantiation
    b200::Lowering lower() override {
        hb_local_desc d;
        std::memset(&d, 0, sizeof(d));
        auto ci = std::make_shared<std::vector<int>>(mask.coefficients());
        d.in = input.rt().view();
        d.kind = HB_LOCAL_REDUCE_DOMAIN; d.reduce_mode = HB_REDUCE_SUM; d.tap = HB_TAP_MUL; d.acc_dtype = HB_S16;
        d.size_x = mask.size_x(); d.size_y = mask.size_y(); d.boundary = (int)input.bmode; d.epilogue = HB_EPI_DIVI_CAST; d.epi_p[0] = 6;
        b200::Lowering L;
        L.kind = b200::Lowering::LOCAL;
        L.launch = [d, ci](const hb_view &out, void *stream) mutable {
            d.out = out; d.coef_s32 = ci->data();
            hipacc_b200::check(hb_local_op(&d, stream), "Deriv");
        };
        return L;
    }
};
class Square1 : public Kernel<short> {
    Accessor<short> &in;
  public:
    Square1(IterationSpace<short> &it, Accessor<short> &in) : Kernel(it), in(in) { add_accessor(&in); }
    void kernel() override { short v = in(); output() = v * v; }
    b200::Lowering lower() override { return b200::point(HB_POINT_SQUARE, {&in}); }
};
class Square2 : public Kernel<short> {
    Accessor<short> &a, &b;
  public:
    Square2(IterationSpace<short> &it, Accessor<short> &a, Accessor<short> &b) : Kernel(it), a(a), b(b) { add_accessor(&a); add_accessor(&b); }
    void kernel() override { output() = a() * b(); }
    b200::Lowering lower() override { return b200::point(HB_POINT_MUL, {&a, &b}); }
};
class Gauss : public Kernel<short> {   // short -> short, int accumulate over all taps, / 16
    Accessor<short> &input; Mask<int> &mask;
  public:
    Gauss(IterationSpace<short> &it, Accessor<short> &in, Mask<int> &mask) : Kernel(it), input(in), mask(mask) { add_accessor(&input); }
    void kernel() override {
        int sum = convolve(mask, Reduce::SUM, [&]() -> int { return input(mask) * mask(); });
        output() = sum / 16;
    }
    b200::Lowering lower() override { return b200::convolve(input, mask, Reduce::SUM, b200::div_int_cast(16), HB_S32); }
};
class Response : public Kernel<uchar> {
    Accessor<short> &dx, &dy, &dxy; float k, threshold;
  public:
    Response(IterationSpace<uchar> &it, Accessor<short> &dx, Accessor<short> &dy, Accessor<short> &dxy, float k, float threshold)
        : Kernel(it), dx(dx), dy(dy), dxy(dxy), k(k), threshold(threshold) { add_accessor(&dx); add_accessor(&dy); add_accessor(&dxy); }
    void kernel() override {
        int x = dx(), y = dy(), xy = dxy();
        float R = ((x * y) - (xy * xy)) - (k * (x + y) * (x + y));
        output() = R > threshold ? 1 : 0;
    }
    b200::Lowering lower() override { return b200::point(HB_POINT_HARRIS, {&dx, &dy, &dxy}, k, threshold); }
};
class HarrisFused : public Kernel<uchar> {
    Accessor<uchar> &in; float k, threshold;
  public:
    HarrisFused(IterationSpace<uchar> &it, Accessor<uchar> &in, float k, float threshold) : Kernel(it), in(in), k(k), threshold(threshold) { add_accessor(&in); }
    void kernel() override { output() = in(); /* stands for the whole pipeline */ }
    b200::Lowering lower() override { return b200::harris(in, k, threshold); }
};

// plain C restatement of the pipeline with CLAMP on every (intermediate) image
static void harris_reference(const uchar *in, uchar *out, int w, int h, float k, float threshold) {
    std::vector<short> dx((size_t)w * h), dy(dx.size()), sx(dx.size()), sy(dx.size()), sxy(dx.size()), gx(dx.size()), gy(dx.size()), gxy(dx.size());
    const int mx[3][3] = {{-1, 0, 1}, {-1, 0, 1}, {-1, 0, 1}}, my[3][3] = {{-1, -1, -1}, {0, 0, 0}, {1, 1, 1}}, mg[3][3] = {{1, 2, 1}, {2, 4, 2}, {1, 2, 1}};
    auto at = [&](const auto &img, int x, int y) { return (int)img[(size_t)tc::clampi(y, 0, h - 1) * w + tc::clampi(x, 0, w - 1)]; };
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            short a = 0, b = 0;
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) { a += (short)(at(in, x + i - 1, y + j - 1) * mx[j][i]); b += (short)(at(in, x + i - 1, y + j - 1) * my[j][i]); }
            dx[(size_t)y * w + x] = a / 6; dy[(size_t)y * w + x] = b / 6;
        }
    for (size_t i = 0; i < dx.size(); ++i) { sx[i] = dx[i] * dx[i]; sy[i] = dy[i] * dy[i]; sxy[i] = dx[i] * dy[i]; }
    auto gauss = [&](const std::vector<short> &src, std::vector<short> &dst) {
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                int s = 0;
                for (int j = 0; j < 3; ++j)
                    for (int i = 0; i < 3; ++i) s += at(src, x + i - 1, y + j - 1) * mg[j][i];
                dst[(size_t)y * w + x] = s / 16;
            }
    };
    gauss(sx, gx); gauss(sy, gy); gauss(sxy, gxy);
    for (size_t i = 0; i < dx.size(); ++i) {
        const int x = gx[i], y = gy[i], xy = gxy[i];
        const float R = ((x * y) - (xy * xy)) - (k * (x + y) * (x + y));
        out[i] = R > threshold ? 1 : 0;
    }
}

int main(int argc, char **argv) {
    const int width = argc > 2 ? std::atoi(argv[1]) : 2048, height = argc > 2 ? std::atoi(argv[2]) : 1024;
    const float k = 0.04f, threshold = 20000.0f;
    const int mx[3][3] = {{-1, 0, 1}, {-1, 0, 1}, {-1, 0, 1}}, my[3][3] = {{-1, -1, -1}, {0, 0, 0}, {1, 1, 1}}, mg[3][3] = {{1, 2, 1}, {2, 4, 2}, {1, 2, 1}};
    std::vector<uchar> input = tc::image_blocks(width, height, 4);

    Image<uchar> in(width, height, input.data()), out(width, height), out_fused(width, height);
    Image<short> dx(width, height), dy(width, height), sx(width, height), sy(width, height), sxy(width, height), gx(width, height), gy(width, height), gxy(width, height);
    Mask<int> maskx(mx), masky(my), maskg(mg);
    Domain domx(maskx), domy(masky);
    float total = 0.0f;
    {
        BoundaryCondition<uchar> bx(in, maskx, Boundary::CLAMP), by(in, masky, Boundary::CLAMP);
        Accessor<uchar> ax(bx), ay(by);
        IterationSpace<short> ix(dx), iy(dy);
        Deriv kx(ix, ax, domx, maskx), ky(iy, ay, domy, masky);
        kx.execute(); total += hipacc_last_kernel_timing();
        ky.execute(); total += hipacc_last_kernel_timing();
        Accessor<short> adx(dx), ady(dy);
        IterationSpace<short> isx(sx), isy(sy), isxy(sxy);
        Square1 s1(isx, adx), s2(isy, ady);
        Square2 s3(isxy, adx, ady);
        s1.execute(); total += hipacc_last_kernel_timing();
        s2.execute(); total += hipacc_last_kernel_timing();
        s3.execute(); total += hipacc_last_kernel_timing();
        BoundaryCondition<short> bsx(sx, maskg, Boundary::CLAMP), bsy(sy, maskg, Boundary::CLAMP), bsxy(sxy, maskg, Boundary::CLAMP);
        Accessor<short> asx(bsx), asy(bsy), asxy(bsxy);
        IterationSpace<short> igx(gx), igy(gy), igxy(gxy);
        Gauss g1(igx, asx, maskg), g2(igy, asy, maskg), g3(igxy, asxy, maskg);
        g1.execute(); total += hipacc_last_kernel_timing();
        g2.execute(); total += hipacc_last_kernel_timing();
        g3.execute(); total += hipacc_last_kernel_timing();
        Accessor<short> agx(gx), agy(gy), agxy(gxy);
        IterationSpace<uchar> iout(out);
        Response r(iout, agx, agy, agxy, k, threshold);
        r.execute(); total += hipacc_last_kernel_timing();
    }
    std::printf("Harris 9-kernel pipeline uchar %dx%d: %.4f ms\n", width, height, total);
    {
        BoundaryCondition<uchar> b(in, 5, 5, Boundary::CLAMP);
        Accessor<uchar> a(b);
        IterationSpace<uchar> it(out_fused);
        HarrisFused f(it, a, k, threshold);
        f.execute();
        std::printf("Harris fused: %.4f ms\n", hipacc_last_kernel_timing());
    }
    std::vector<uchar> ref((size_t)width * height);
    harris_reference(input.data(), ref.data(), width, height, k, threshold);
    long corners = 0;
    for (uchar v : ref) corners += v;
    std::printf("corners in the reference: %ld\n", corners);
    long first = -1;
    int rc = tc::verdict("harris unfused", tc::count_diff(out.data(), ref.data(), ref.size(), 0, &first), ref.size(), first);
    rc |= tc::verdict("harris fused", tc::count_diff(out_fused.data(), ref.data(), ref.size(), 0, &first), ref.size(), first);
    return rc | (corners > 0 ? 0 : 1);
}
