"""The reference's OWN sample programs (samples-public/*/src/main.cpp), UNMODIFIED, built against the B200 front.

Each sample is a complete Hipacc DSL program: kernel classes, a main() that runs them and -- for most -- a plain C
reference of the same operator with a comparison that prints "Test PASSED" / "Test FAILED".  Here the file is compiled
where it lies under /root/reference by nvcc with

    -I include/hipacc_b200/compat      its `#include "hipacc.hpp"` finds the B200 front instead of the reference DSL
    -I <reference>/samples-public/common   hipacc_helper.hpp (timing / comparison helpers of the samples)

so every kernel() body becomes a device kernel (include/hipacc_b200/hipacc.hpp, compiled-body path) and every Image lives in
HBM.  Nothing of the reference is copied into the repository: the sources are read at build time in the build container
(like oracle/_ref), the binaries land in tests/cpp/bin/ref_samples/ (git-ignored, travels to the GPU box) and the GPU test
runs whichever binaries are there.  This is the drop-in claim at the outermost boundary: a user's existing DSL source.
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HIPACC_REFERENCE", "/root/reference")
SAMPLES_DIR = os.path.join(REF, "samples-public")
OUT = os.path.join(ROOT, "tests", "cpp", "bin", "ref_samples")
LIBDIR = os.path.join(ROOT, "hipacc_b200", "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# sample directory -> what its own main() must print.  "PASSED": the sample compares against its embedded C reference;
# "RUNS": the sample has no comparison (it only times the pipeline), exit status 0 is what can be checked.
SAMPLES = {
    "0_Point_Operators/Color_Conversion": "PASSED",
    "0_Point_Operators/Scaling": "PASSED",              # interpolating Accessor (Interpolate::LF)
    "0_Point_Operators/Windowing": "PASSED",            # crop Accessors / IterationSpace offsets
    "1_Local_Operators/Box_Blur": "PASSED",
    "1_Local_Operators/Box_Blur_RGBA": "PASSED",
    "1_Local_Operators/Dilate": "PASSED",
    "1_Local_Operators/Dilate_RGBA": "PASSED",
    "1_Local_Operators/Erode": "PASSED",
    "1_Local_Operators/Erode_RGBA": "PASSED",
    "1_Local_Operators/Gaussian_Blur": "PASSED",
    "1_Local_Operators/Gaussian_Blur_RGBA": "PASSED",
    "1_Local_Operators/Laplace": "PASSED",
    "1_Local_Operators/Laplace_RGBA": "PASSED",
    "1_Local_Operators/Unsharp": "RUNS",
    "2_Global_Operators/Histogram": "PASSED",           # binning() + reduce() bodies compiled for the device
    "2_Global_Operators/Reduction_Max": "PASSED",
    "2_Global_Operators/Reduction_Sum": "PASSED",
    "3_Preprocessing/Bilateral_Filter": "PASSED",
    "3_Preprocessing/Bilateral_Filter_RGBA": "PASSED",
    "3_Preprocessing/Harris_Corner": "RUNS",
    "3_Preprocessing/ShiTomasi_Corner": "RUNS",
    "3_Preprocessing/Sobel": "PASSED",
    "3_Preprocessing/Sobel_RGBA": "PASSED",
    "4_Postprocessing/Night_Filter": "RUNS",
    "5_Other/Gaussian_Laplacian_Pyramid": "PASSED",     # Pyramid / traverse, NN + LF accessors
    "6_Test/Kernel_Fusion_L2L": "PASSED",
    "6_Test/Kernel_Fusion_L2P": "PASSED",
    "6_Test/Kernel_Fusion_Mixed": "PASSED",
    "6_Test/Kernel_Fusion_P2L": "PASSED",
    "6_Test/Kernel_Fusion_P2P": "PASSED",
}

# Samples that are only BUILT (unmodified, like the ones above): they compile against the front for the device, which is
# what says the DSL surface they use exists; they are not part of the GPU run list -- Optical_Flow and Motion_Interpolation
# abort on a DSL assert in the reference itself (an accessor that is never registered), the other
# five carry no comparison against a C reference and were added after the round's GPU time was spent.
BUILD_ONLY = {
    "3_Preprocessing/Optical_Flow": "BUILDS",
    "4_Postprocessing/Bokeh_Effect": "BUILDS",
    "4_Postprocessing/Chromatic_Abberation": "BUILDS",
    "5_Other/Game_of_Life": "BUILDS",
    "5_Other/Mandelbrot": "BUILDS",
    "5_Other/Motion_Interpolation": "BUILDS",
    "5_Other/WCE_Enhance": "BUILDS",
}


def name_of(sample):
    return os.path.basename(sample)


def have_reference():
    return os.path.isdir(SAMPLES_DIR)


def build_one(sample):
    src = os.path.join(SAMPLES_DIR, sample, "src", "main.cpp")
    exe = os.path.join(OUT, name_of(sample))
    hdrs = [os.path.join(ROOT, "include", "hipacc_b200", h) for h in ("hipacc.hpp", "hipacc_rt.hpp", "hipacc_types.hpp")] + [os.path.join(ROOT, "include", "hipacc_b200.h")]
    if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(p) for p in [src] + hdrs):
        return exe, ""
    os.makedirs(OUT, exist_ok=True)
    cmd = [NVCC, "-x", "cu", "-std=c++17", "-O2", "-fmad=false", "-DHIPACC_B200_DEVICE_GLOBAL_OPS", "-gencode", "arch=compute_100a,code=sm_100a",
           "-ccbin", "/usr/bin/g++", "-Xcompiler", "-ffp-contract=off", "-w",
           "-I", os.path.join(ROOT, "include", "hipacc_b200", "compat"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(SAMPLES_DIR, "common"),
           src, "-L", LIBDIR, "-lhipacc_b200", "-Xlinker", f"-rpath={LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return (exe if r.returncode == 0 else None), r.stdout + r.stderr


def build_all(jobs=8):
    """build every sample (parallel); returns {sample: error text} for the ones that failed"""
    import concurrent.futures as cf
    from hipacc_b200 import build as hb_build
    hb_build.build()
    failed = {}
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        todo = list(SAMPLES) + list(BUILD_ONLY)
        for sample, (exe, log) in zip(todo, ex.map(build_one, todo)):
            if exe is None:
                failed[sample] = log
    return failed


if __name__ == "__main__":
    import sys
    sys.path.insert(0, ROOT)
    bad = build_all()
    for s, log in bad.items():
        print("FAILED to build", s, "\n", log[-2000:])
    n = len(SAMPLES) + len(BUILD_ONLY)
    print(f"built {n - len(bad)} of {n} reference samples into {OUT}")
