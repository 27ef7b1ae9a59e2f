// Compile-only translation unit (tests/test_cpp_front.py::test_runtime_graph_wrappers_compile): the -use-graph wrappers
// of include/hipacc_b200/hipacc_rt.hpp in the shape generated host code would use them.  Never executed by the tests.
#include "hipacc_b200/hipacc_rt.hpp"

void record_and_replay(const HipaccAccessor<float> &in, const HipaccAccessor<float> &out, hb_local_desc desc,
                       HipaccExecutionParameterCuda const &ep, int frames) {
    HipaccGraph graph;
    hipaccGraphBegin(ep);
    hipaccLaunchLocalOperator(in, out, desc, ep);
    hipaccGraphEnd(ep, graph);
    for (int f = 0; f < frames; ++f) hipaccGraphLaunch(graph, ep);
}

int main() { return 0; }
