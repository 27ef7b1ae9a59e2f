// Compiled-body path of the DSL front (include/hipacc_b200/hipacc.hpp under nvcc): an RGBA -> gray point operator in the
// shape of samples-public/0_Point_Operators/Color_Conversion/src/main.cpp.  The Kernel subclass has NO lower(): its
// kernel() body is compiled for the device and runs one thread per pixel; checked bit for bit against a plain C loop
// (build with -fmad=false: the C loop is compiled without FMA contraction).
//   usage: dsl_color_conversion [width height] [--io in.raw out.raw]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;

class ColorConversion : public Kernel<uchar> {
  private:
    Accessor<uchar4> &in;

  public:
    ColorConversion(IterationSpace<uchar> &iter, Accessor<uchar4> &acc) : Kernel(iter), in(acc) { add_accessor(&in); }

    void kernel() {
        uchar4 pixel = in();
        output() = .3f * pixel.x + .59f * pixel.y + .11f * pixel.z + .5f;
    }
};

int main(int argc, char **argv) {
    const tc::Args a(argc, argv, 1030, 517);
    std::vector<uchar4> input((size_t)a.w * a.h);
    if (a.in) tc::read_raw(a.in, input);
    else {
        const std::vector<unsigned char> raw = tc::image_u8(a.w * 4, a.h, 11);
        std::memcpy(input.data(), raw.data(), raw.size());
    }
    Image<uchar4> in(a.w, a.h, input.data());
    Image<uchar> out(a.w, a.h);
    Accessor<uchar4> acc(in);
    IterationSpace<uchar> iter(out);
    ColorConversion filter(iter, acc);
    filter.execute();
    const float ms = hipacc_last_kernel_timing();
    uchar *result = out.data();
    std::printf("compiled-body color conversion %dx%d: %.4f ms\n", a.w, a.h, ms);
    if (a.out) tc::write_raw(a.out, result, (size_t)a.w * a.h);

    std::vector<uchar> ref((size_t)a.w * a.h);
    for (size_t i = 0; i < ref.size(); ++i) ref[i] = .3f * input[i].x + .59f * input[i].y + .11f * input[i].z + .5f;
    long first = -1;
    const long bad = tc::count_diff(result, ref.data(), ref.size(), 0, &first);
    return tc::verdict("dsl_color_conversion", bad, ref.size(), first);
}
