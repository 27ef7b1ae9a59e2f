"""CPU tests: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/hipacc_b200.h declares; host-side logic (views, specs, pyramid sizes).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hipacc_b200 import _abi as A, specs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_and_export_list_agree():
    hdr = open(os.path.join(ROOT, "include", "hipacc_b200.h")).read()
    declared = set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", hdr)) - {"hb_log_fn"}
    assert declared == set(A.EXPORTS), declared ^ set(A.EXPORTS)


def test_library_loads_and_exports_all_symbols(hb):
    L = hb.lib()
    for sym in A.EXPORTS:
        assert hasattr(L, sym), f"libhipacc_b200.so does not export {sym}"


def test_no_device_is_an_error_not_a_fallback(hb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    L = hb.lib()
    assert L.hb_device_count() == 0
    assert L.hb_init(0) == A.HB_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.hb_last_error()


def test_struct_layouts_match_the_header():
    # sizes computed by hand from include/hipacc_b200.h on LP64
    assert C.sizeof(A.hb_view) == 48
    assert C.sizeof(A.hb_local_desc) == 2 * 48 + 6 * 4 + 3 * 8 + 4 + 4 + 8 + 4 + 4 + 24
    assert C.sizeof(A.hb_point_desc) == 3 * 48 + 12 + 4 + 48 + 4 + 4 + 16
    assert C.sizeof(A.hb_harris_desc) == 2 * 48 + 8
    assert C.sizeof(A.hb_pyr_up_desc) == 4 * 48


def test_struct_layouts_match_the_compiler(tmp_path):
    """Compile a tiny C program against the real header and compare sizeof / offsetof."""
    import subprocess
    src = tmp_path / "lay.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hipacc_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(hb_view), sizeof(hb_local_desc),'
                   'sizeof(hb_bilateral_desc), sizeof(hb_point_desc), sizeof(hb_harris_desc), sizeof(hb_pyr_down_desc),'
                   'sizeof(hb_pyr_up_desc), offsetof(hb_local_desc, epi_p), offsetof(hb_point_desc, p));return 0;}\n')
    exe = tmp_path / "lay"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(A.hb_view), C.sizeof(A.hb_local_desc), C.sizeof(A.hb_bilateral_desc), C.sizeof(A.hb_point_desc),
            C.sizeof(A.hb_harris_desc), C.sizeof(A.hb_pyr_down_desc), C.sizeof(A.hb_pyr_up_desc),
            A.hb_local_desc.epi_p.offset, A.hb_point_desc.p.offset]
    assert got == want


def test_new_struct_layouts_match_the_compiler(tmp_path):
    """binning / halo / IPC descriptors: ctypes mirror vs the real header"""
    import subprocess
    src = tmp_path / "lay2.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hipacc_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(hb_binning_desc), offsetof(hb_binning_desc, p0),'
                   'sizeof(hb_halo_desc), offsetof(hb_halo_desc, ctrl), offsetof(hb_halo_desc, down_ghost_top), sizeof(hb_ipc_mem),'
                   '(size_t)HB_U8X4);return 0;}\n')
    exe = tmp_path / "lay2"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(A.hb_binning_desc), A.hb_binning_desc.p0.offset, C.sizeof(A.hb_halo_desc), A.hb_halo_desc.ctrl.offset,
            A.hb_halo_desc.down_ghost_top.offset, C.sizeof(A.hb_ipc_mem), A.U8X4]
    assert got == want


def test_gather_desc_layout_matches_the_compiler(tmp_path):
    import subprocess
    src = tmp_path / "lay3.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hipacc_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %d\\n", sizeof(hb_gather_desc), offsetof(hb_gather_desc, ctrl),'
                   'offsetof(hb_gather_desc, peer_ctrl), offsetof(hb_gather_desc, peer_slot), HB_MAX_PEERS);return 0;}\n')
    exe = tmp_path / "lay3"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(A.hb_gather_desc), A.hb_gather_desc.ctrl.offset, A.hb_gather_desc.peer_ctrl.offset,
                   A.hb_gather_desc.peer_slot.offset, A.HB_MAX_PEERS]


def test_pyramid_sizes_truncate_like_the_reference():
    assert S.pyramid_sizes(16384, 16384, 8)[-1] == (128, 128)
    assert S.pyramid_sizes(101, 67, 3) == [(101, 67), (50, 33), (25, 16)]
    with pytest.raises(AssertionError):
        S.pyramid_sizes(8, 8, 5)


def test_spec_fill_keeps_host_arrays_alive():
    spec = S.gaussian_blur(np.ones((3, 3), np.float32) / 9)
    d = A.hb_local_desc()
    spec.fill(d)
    assert d.size_x == 3 and d.epilogue == A.EPI_ADD_CAST and d.epi_p[0] == 0.5
    assert abs(d.coef_f32[4] - 1 / 9) < 1e-7 and not d.coef_s32


def test_pair_kernel_sass_keeps_multiply_and_add_separately_rounded():
    """hb_local_pair.cu's contract: FMUL2 + FADD2, never FFMA2 (ptxas contracts mul.rn.f32x2 -> add.rn.f32x2 unless the
    product reaches the addition with its halves exchanged) -- a contracted product would round once instead of twice."""
    import shutil
    import subprocess
    from hipacc_b200 import build as hb_build
    hb_build.build()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    obj = os.path.join(os.path.dirname(hb_build.LIB), "hb_local_pair.o")
    sass = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True, check=True).stdout
    assert "FMUL2" in sass and "FADD2" in sass
    assert "FFMA2" not in sass   # (scalar FFMA does appear: the correctly rounded division of the DIVF epilogue)


def _sass_of(obj_name, kernel_substr):
    """SASS text of the kernels of one object whose mangled name contains `kernel_substr` (cuobjdump; no GPU needed)"""
    import shutil
    import subprocess
    from hipacc_b200 import build as hb_build
    hb_build.build()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", os.path.join(os.path.dirname(hb_build.LIB), obj_name)], capture_output=True, text=True, check=True).stdout
    keep, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = kernel_substr in line
        if on:
            keep.append(line)
    assert keep, f"no kernel matching {kernel_substr} in {obj_name}"
    return "\n".join(keep)


def test_harris_v3_sass_is_tma_staged_and_dot_product_based():
    """hb_harris.cu version 3: the byte tile arrives by TMA on an mbarrier, the Sobel sums and the binomial run on the integer
    dot-product instructions, the default instantiation (5 CTAs per SM) has no local-memory spills, and stage C realigns
    nothing with byte permutes (the 21 PRMT left are the byte windows of stage B, three per staged input row)."""
    sass = _sass_of("hb_harris.o", "harris_fused3_kernelILi5E")
    assert "UTMALDG.2D" in sass and "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass
    assert sass.count("IDP.4A") >= 56 and sass.count("IDP.2A") >= 96
    assert "STL" not in sass and "LDL" not in sass
    assert sass.count("PRMT") <= 30


def test_pyr_down_and_local_tma_sass_use_tma():
    for obj, kern in (("hb_pyramid.o", "pyr_down_fused_kernelILi5ELb1ELi6E"), ("hb_local_tma.o", "local_tma_f32_kernel")):
        sass = _sass_of(obj, kern)
        assert "UTMALDG.2D" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass, (obj, kern)
    # the TMA-staged instantiation of the fused down step is the 6-CTA one: 40 registers, and what spills sits in the
    # border tiles' loader only (the 5-CTA instantiation, used when TMA cannot address the image, has none)
    assert "STL" not in _sass_of("hb_pyramid.o", "pyr_down_fused_kernelILi5ELb1ELi5E")


def test_binning_sass_is_branch_free_shared_atomics():
    """hb_binning.cu, shared-memory path: native shared-memory atomics (ATOMS), no generic-address ATOM, and the four pixels
    of a 16-byte load are binned without a branch between their atomics."""
    sass = _sass_of("hb_binning.o", "binning_kernelIfLi0ELi0ELb1ELb1E")
    assert "ATOMS" in sass
    lines = [l for l in sass.splitlines() if "/*0" in l or "/*1" in l or "/*2" in l or "/*3" in l]
    idx = [i for i, l in enumerate(lines) if "ATOMS" in l]
    runs = [b - a for a, b in zip(idx, idx[1:])]
    assert runs and min(runs) <= 2          # consecutive atomics of one vector sit next to each other ...
    first = idx[0]
    assert not any("BRA" in l for l in lines[first:first + 6])   # ... with no branch in between
    assert "ATOM.E" not in sass.replace("ATOMS", "")
