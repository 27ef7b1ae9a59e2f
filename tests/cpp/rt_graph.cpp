// rt_graph.cpp -- the -use-graph wrappers of include/hipacc_b200/hipacc_rt.hpp (hipaccGraphBegin / End / Launch,
// reference: runtime/hipacc_cu_standalone.hpp:331-356) with the front's DEFAULT configuration, i.e. with kernel timing
// switched on as hipaccInitCUDA leaves it: a captured launch must not be timed (an event synchronise inside a capture
// invalidates it), the replays must compute, and a blocking reduction recorded with the async form must stay valid
// after a larger reduction ran in between.
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.hpp"
#include "hipacc_b200/hipacc_rt.hpp"

int main() {
    const int w = 1000, h = 600;
    hipaccInitCUDA();   // timing enabled by default
    void *stream = nullptr;
    hipacc_b200::check(hb_stream_create(&stream), "hb_stream_create");
    HipaccExecutionParameterCuda ep = std::make_shared<HipaccStreamParameter>(stream);

    std::vector<float> a = tc::image_f32(w, h, 11), b = tc::image_f32(w, h, 12);
    auto in = hipaccCreateMemory<float>(a.data(), w, h);
    auto out = hipaccCreateMemory<float>(nullptr, w, h);
    const float lap[9] = {0, 1, 0, 1, -4, 1, 0, 1, 0};
    hb_local_desc d;
    std::memset(&d, 0, sizeof(d));
    d.kind = HB_LOCAL_REDUCE_DOMAIN; d.reduce_mode = HB_REDUCE_SUM; d.tap = HB_TAP_MUL; d.acc_dtype = HB_F32;
    d.size_x = d.size_y = 3; d.coef_f32 = lap; d.boundary = HB_BOUNDARY_MIRROR; d.epilogue = HB_EPI_CAST;

    HipaccGraph graph;
    hipaccGraphBegin(ep);
    hipaccLaunchLocalOperator(hipaccMakeAccessor<float>(in), hipaccMakeAccessor<float>(out), d, ep);
    hipaccGraphEnd(ep, graph);
    if (!graph.get()) { std::printf("rt_graph: Test FAILED, capture produced no graph\n"); return 1; }

    int rc = 0;
    for (int frame = 0; frame < 2; ++frame) {
        std::vector<float> &src = frame ? b : a;
        hipaccWriteMemory(in, src.data());
        hipaccGraphLaunch(graph, ep);
        hb_stream_synchronize(stream);
        const float *got = hipaccReadMemory(out);
        std::vector<float> want((size_t)w * h);
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                auto px = [&](int dx, int dy) { return src[(size_t)tc::mirrori(y + dy, h) * w + tc::mirrori(x + dx, w)]; };
                float acc = 1.0f * px(0, -1);   // zero taps are Domain holes; row-major order, first visited tap initialises
                acc = acc + 1.0f * px(-1, 0);
                acc = acc + -4.0f * px(0, 0);
                acc = acc + 1.0f * px(1, 0);
                acc = acc + 1.0f * px(0, 1);
                want[(size_t)y * w + x] = acc;
            }
        long first = -1;
        const long bad = tc::count_diff_rel(got, want.data(), want.size(), 0.0, 0.0, &first);
        rc |= tc::verdict(frame ? "rt_graph replay 2" : "rt_graph replay 1", bad, want.size(), first);
    }
    hb_stream_destroy(stream);
    return rc;
}
