// Global operator with binning (samples-public/2_Global_Operators/Histogram/src/main.cpp): a float image is
// binned into num_bins counters with `bin(pixel/255.0f*num_bins()) = 1` and `reduce(l, r) = l + r`, read with
// binned_data().  Compared bin for bin with the plain C loop of the sample.     usage: histogram [width height bins]
#include "common.hpp"
#include "hipacc_b200/hipacc.hpp"

using namespace hipacc;
using namespace hipacc::math;

class Histogram : public Kernel<float, uint> {
    Accessor<float> &in;

  public:
    Histogram(IterationSpace<float> &iter, Accessor<float> &in) : Kernel(iter), in(in) { add_accessor(&in); }
    void kernel() override { output() = in(); }
    void binning(unsigned int, unsigned int, float pixel) override { bin(pixel / 255.0f * num_bins()) = 1; }
    uint reduce(uint left, uint right) const override { return left + right; }
    b200::Lowering lower() override { return b200::point(HB_POINT_COPY, {&in}); }
    b200::Binning lower_binning() override { return b200::bin_scaled_count(255.0); }
};

int main(int argc, char **argv) {
    const int width = argc > 2 ? std::atoi(argv[1]) : 4096, height = argc > 2 ? std::atoi(argv[2]) : 4096;
    const unsigned num_bins = argc > 3 ? (unsigned)std::atoi(argv[3]) : 256;
    std::vector<float> input = tc::image_f32(width, height, 21, 254.99f);

    Image<float> in(width, height, input.data());
    Image<float> out(width, height);
    Accessor<float> acc(in);
    IterationSpace<float> iter(out);
    Histogram filter(iter, acc);
    filter.execute();
    uint *output = filter.binned_data(num_bins);
    std::printf("histogram %u bins float %dx%d: %.4f ms\n", num_bins, width, height, hipacc_last_kernel_timing());

    std::vector<uint> ref(num_bins, 0u);
    for (size_t p = 0; p < input.size(); ++p) ref[(uint)(input[p] / 255.0f * num_bins)] += 1;
    long first = -1;
    const int rc = tc::verdict("histogram", tc::count_diff(output, ref.data(), ref.size(), 0, &first), ref.size(), first);
    delete[] output;
    return rc;
}
