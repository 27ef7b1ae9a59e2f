"""Build libhipacc_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m hipacc_b200.build [--force] [--verbose]

Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false
  -fmad=false is part of the numeric contract (DESIGN.md "Numerics"): the DSL's float multiply and
  add are separately rounded; FMAs are only used where written explicitly.
The shared library is git-ignored but travels with the repo snapshot to the GPU box.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libhipacc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-std=c++17", "-O3", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC",
          "-ccbin", "/usr/bin/g++"] + os.environ.get("HB_EXTRA_NVCC_FLAGS", "").split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "hipacc_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, obj, verbose):
    cmd = [NVCC] + ARCH + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    hdrs = headers()
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(OUT_DIR, src[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, src)] + hdrs):
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, log in ex.map(lambda j: _compile(j[0], j[1], verbose), jobs):
                if verbose or rc != 0:
                    sys.stderr.write(f"--- nvcc {src}\n{log}\n")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    if jobs or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ccbin", "/usr/bin/g++"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
