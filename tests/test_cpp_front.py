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
PROGRAMS = ["c1_gaussian_blur", "c2_sobel_laplace_f32", "c3_bilateral_reduce", "c4_harris", "c5_pyramid", "histogram", "gaussian_blur_rgba", "rt_generated_host", "rt_graph",
            "rt_reference_names"]
# translation units compiled by nvcc: their kernel() bodies are compiled for the device (hipacc.hpp, "compiled-body path")
CU_PROGRAMS = ["dsl_color_conversion", "dsl_unsharp", "dsl_night_filter", "dsl_lowering_check"]
# small sizes keep the embedded plain C loops to about a second each; the full BASELINE sizes are covered by
# tests/test_gpu_parity.py::test_full_size_*
ARGS = {"c1_gaussian_blur": ["1531", "1027"], "c2_sobel_laplace_f32": ["2048", "1100"], "c3_bilateral_reduce": ["640", "333"],
        "c4_harris": ["1500", "700"], "c5_pyramid": ["1000", "744", "5"], "histogram": ["2050", "1033", "256"], "gaussian_blur_rgba": ["1031", "517"], "rt_generated_host": [], "rt_graph": [],
        "rt_reference_names": [], "dsl_color_conversion": ["1030", "517"], "dsl_unsharp": ["1030", "517"], "dsl_night_filter": ["700", "413"], "dsl_lowering_check": ["good"]}


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


NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def compile_cu_program(name, src=None, out=None):
    """nvcc -x cu: -fmad=false / -ffp-contract=off so that the device body and the plain C loop of the same file round alike"""
    from hipacc_b200 import build as hb_build
    hb_build.build()
    os.makedirs(BIN, exist_ok=True)
    exe = os.path.join(BIN, out or name)
    src = src or os.path.join(CPP, name + ".cu")
    if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(p) for p in
            [src, os.path.join(CPP, "common.hpp")] + [os.path.join(ROOT, "include", "hipacc_b200", h) for h in ("hipacc.hpp", "hipacc_rt.hpp", "hipacc_types.hpp")]):
        return exe
    cmd = [NVCC, "-x", "cu", "-std=c++17", "-O2", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
           "-Xcompiler", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), src, "-L", LIBDIR, "-lhipacc_b200",
           "-Xlinker", f"-rpath={LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


@pytest.mark.parametrize("name", PROGRAMS)
def test_program_compiles_and_links(name):
    exe = compile_program(name)
    assert os.path.exists(exe)


@pytest.mark.parametrize("name", CU_PROGRAMS)
def test_compiled_body_program_builds(name):
    """the kernel() bodies compile for sm_100a (nvcc cross-compiles without a GPU) and a dsl_kernel<...> is in the binary"""
    exe = compile_cu_program(name)
    r = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", exe], capture_output=True, text=True)
    assert "dsl_kernel" in r.stdout


def test_unmodified_dsl_source_compiles_both_ways():
    """the same DSL source (kernel() + lower()) builds with the host compiler alone AND with nvcc, where the body becomes
    device code next to the lowering"""
    exe = compile_cu_program("c1_gaussian_blur", src=os.path.join(CPP, "c1_gaussian_blur.cpp"), out="c1_gaussian_blur_nvcc")
    assert os.path.exists(exe)


def test_front_headers_need_no_cuda_toolkit():
    """Under the host compiler the including translation unit sees only the C ABI: no CUDA header, no torch (the
    compiled-body path of hipacc.hpp includes cuda_runtime.h only under nvcc)."""
    for h in ("hipacc.hpp", "hipacc_rt.hpp", "hipacc_types.hpp"):
        src = open(os.path.join(ROOT, "include", "hipacc_b200", h)).read()
        assert "torch" not in src
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-M", "-I", os.path.join(ROOT, "include"), "-x", "c++",
                        os.path.join(ROOT, "include", "hipacc_b200", "hipacc.hpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "cuda" not in r.stdout.lower(), r.stdout


# ------------------------------------------------------------------ the reference's own sample programs, unmodified
import ref_samples  # noqa: E402  (tests/ref_samples.py)


@pytest.mark.skipif(not ref_samples.have_reference(), reason="the reference tree exists in the build container only")
def test_reference_samples_build_unmodified():
    """every sample of samples-public/ in ref_samples.SAMPLES compiles against the B200 front without touching its source"""
    failed = ref_samples.build_all()
    assert not failed, "\n".join(f"{k}:\n{v[-1500:]}" for k, v in failed.items())


@pytest.mark.gpu
@pytest.mark.parametrize("sample", sorted(ref_samples.SAMPLES))
def test_reference_sample_runs_on_device(sample):
    """the sample's own main(): its DSL kernels run on the B200, its embedded plain C reference decides PASSED / FAILED"""
    exe = os.path.join(ref_samples.OUT, ref_samples.name_of(sample))
    if not os.path.exists(exe):
        pytest.skip("binary not built (tests/ref_samples.py builds it where /root/reference exists)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, cwd=ref_samples.OUT)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "FAILED" not in out and "ERROR" not in out, out[-3000:]
    if ref_samples.SAMPLES[sample] == "PASSED":
        assert "Test PASSED" in out, out[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CU_PROGRAMS)
def test_compiled_body_program_runs_on_device(name):
    exe = compile_cu_program(name)
    r = subprocess.run([exe] + ARGS[name], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "FAILED" not in out and "PASSED" in out, out


@pytest.mark.gpu
def test_wrong_lowering_is_caught():
    """HIPACC_B200_CHECK_LOWERING=1: a lower() that does not describe the kernel() body aborts the program"""
    exe = compile_cu_program("dsl_lowering_check")
    r = subprocess.run([exe, "wrong"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "DISAGREE" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_lowering_check_of_the_c1_program():
    """c1_gaussian_blur.cpp built by nvcc: the library kernel lower() names and the compiled body agree on every pixel"""
    exe = compile_cu_program("c1_gaussian_blur", src=os.path.join(CPP, "c1_gaussian_blur.cpp"), out="c1_gaussian_blur_nvcc")
    r = subprocess.run([exe, "1531", "1027"], capture_output=True, text=True, timeout=600, env={**os.environ, "HIPACC_B200_CHECK_LOWERING": "1"})
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "PASSED" in out and "DISAGREE" not in out, out


@pytest.mark.gpu
@pytest.mark.parametrize("name", PROGRAMS)
def test_program_runs_on_device(name):
    exe = compile_program(name)
    r = subprocess.run([exe] + ARGS[name], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "FAILED" not in out and "PASSED" in out, out
