"""The Hipacc-compatible C++ front (include/hipacc_b200/hipacc.hpp = DSL surface, hipacc_rt.hpp = runtime
surface) on top of the C ABI.  tests/cpp/*.cpp are Hipacc programs in the shape of the reference's own samples
(samples-public/*/src/main.cpp): DSL kernels, a plain C reference loop in the same file, "Test PASSED".

not gpu : every program compiles and links against libhipacc_b200.so with the host compiler alone
gpu     : every program runs on the B200 and its embedded comparison passes
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "bin")
LIBDIR = os.path.join(ROOT, "hipacc_b200", "lib")
PROGRAMS = ["c1_gaussian_blur", "c2_sobel_laplace_f32", "c3_bilateral_reduce", "c4_harris", "c5_pyramid", "histogram", "gaussian_blur_rgba", "rt_generated_host", "rt_graph"]
# small sizes keep the embedded plain C loops to about a second each; the full BASELINE sizes are covered by
# tests/test_gpu_parity.py::test_full_size_*
ARGS = {"c1_gaussian_blur": ["1531", "1027"], "c2_sobel_laplace_f32": ["2048", "1100"], "c3_bilateral_reduce": ["640", "333"],
        "c4_harris": ["1500", "700"], "c5_pyramid": ["1000", "744", "5"], "histogram": ["2050", "1033", "256"], "gaussian_blur_rgba": ["1031", "517"], "rt_generated_host": [], "rt_graph": []}


def compile_program(name):
    from hipacc_b200 import build as hb_build
    hb_build.build()
    os.makedirs(BIN, exist_ok=True)
    exe = os.path.join(BIN, name)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fopenmp", "-ffp-contract=off", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(CPP, name + ".cpp"), "-L", LIBDIR, "-lhipacc_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


@pytest.mark.parametrize("name", PROGRAMS)
def test_program_compiles_and_links(name):
    exe = compile_program(name)
    assert os.path.exists(exe)


def test_front_headers_need_no_cuda_toolkit():
    """The including translation unit sees only the C ABI: no cuda_runtime.h, no torch."""
    for h in ("hipacc.hpp", "hipacc_rt.hpp"):
        src = open(os.path.join(ROOT, "include", "hipacc_b200", h)).read()
        assert "cuda_runtime" not in src and "torch" not in src


@pytest.mark.gpu
@pytest.mark.parametrize("name", PROGRAMS)
def test_program_runs_on_device(name):
    exe = compile_program(name)
    r = subprocess.run([exe] + ARGS[name], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "FAILED" not in out and "PASSED" in out, out
