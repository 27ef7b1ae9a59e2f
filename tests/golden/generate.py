"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref: the reference's own DSL
headers + sample kernel classes executed as C++).  Run in the build container, where
/root/reference exists:

    python tests/golden/generate.py

Inputs are not stored: they are regenerated from hipacc_b200.synth (pure functions of the seed).
Only the reference's outputs are committed, so the oracle and the CUDA path can be pinned
against the reference even where neither /root/reference nor oracle/_ref is available.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hipacc_b200 import _abi as A, masks as M, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BMODES = [A.CLAMP, A.REPEAT, A.MIRROR, A.CONSTANT]

# shared with tests/test_golden.py
U8_SHAPE = (45, 70)     # (h, w), deliberately not multiples of any tile size
F32_SHAPE = (39, 66)
HARRIS_SHAPE = (96, 128)
PYR_CASES = [(64, 96, 4, 5), (50, 77, 3, 3)]  # h, w, depth, mask size


def main():
    out = {}
    u8 = synth.image_np("uint8", U8_SHAPE[1], U8_SHAPE[0], seed=1)
    f32 = synth.image_np("float32", F32_SHAPE[1], F32_SHAPE[0], seed=2)
    f255 = synth.image_np("float32", F32_SHAPE[1], F32_SHAPE[0], seed=3, scale=255.0)
    for b in BMODES:
        for sz in (3, 5, 7):
            out[f"gauss_u8_{sz}_{b}"] = O.ref_gaussian_u8(u8, M.GAUSS[sz], b)
        for name, m in (("sobel3x", M.SOBEL3_X), ("sobel3y", M.SOBEL3_Y), ("sobel5x", M.SOBEL5_X)):
            out[f"{name}_u8_{b}"] = O.ref_sobel_u8(u8, m, b)
            out[f"{name}_f32_{b}"] = O.ref_local_f32(f32, m.astype(np.float32), 1, A.SUM, b)
        for name, m in (("lap3", M.LAPLACE3), ("lap3n4", M.LAPLACE3_4N), ("lap5", M.LAPLACE5)):
            out[f"{name}_u8_{b}"] = O.ref_laplace_u8(u8, m, b)
            out[f"{name}_f32_{b}"] = O.ref_local_f32(f32, m.astype(np.float32), 1, A.SUM, b)
        out[f"gauss5_f32_{b}"] = O.ref_local_f32(f32, M.GAUSS5, 0, A.SUM, b)
        out[f"dilate_u8_{b}"] = O.ref_minmax_u8(u8, 5, 3, 1, b)
        out[f"erode_u8_{b}"] = O.ref_minmax_u8(u8, 3, 5, 0, b)
        out[f"box_u8_{b}"] = O.ref_box_u8(u8, 5, 5, b)
    a = O.ref_sobel_u8(u8, M.SOBEL3_X, A.CLAMP)
    bb = O.ref_sobel_u8(u8, M.SOBEL3_Y, A.CLAMP)
    out["sobel_combine"] = O.ref_sobel_combine(a, bb, 4)
    for sz in (3, 5, 13):
        out[f"bilateral_u8_{sz}"] = O.ref_bilateral(u8, sz, M.bilateral_mask(sz), 16, A.CLAMP)
        out[f"bilateral_f32_{sz}"] = O.ref_bilateral(f255, sz, M.bilateral_mask(sz), 16, A.MIRROR)
    himg = synth.blocks_np(HARRIS_SHAPE[1], HARRIS_SHAPE[0], seed=5)
    c, gx, gy, gxy = O.ref_harris_u8(himg)
    out["harris_out"], out["harris_gx"], out["harris_gy"], out["harris_gxy"] = c, gx, gy, gxy
    out["interp_nn"] = O.ref_interp_f32(f32, F32_SHAPE[1] // 2, F32_SHAPE[0] // 2, A.INTERP_NN)
    out["interp_lf"] = O.ref_interp_f32(out["interp_nn"], F32_SHAPE[1], F32_SHAPE[0], A.INTERP_LF)
    for i, (h, w, d, sz) in enumerate(PYR_CASES):
        img = synth.image_np("float32", w, h, seed=7 + i)
        g, l = O.ref_pyramid_f32(img, d, M.GAUSS[sz])
        for lv in range(d):
            out[f"pyr{i}_g{lv}"], out[f"pyr{i}_l{lv}"] = g[lv], l[lv]
    for op, nm in ((0, "sum"), (1, "min"), (2, "max")):
        out[f"reduce_{nm}"] = np.array([O.ref_global_reduce_f32(f32, op)], np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_dsl.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_dsl.npz")), "bytes")


HIST_SHAPE = (61, 83)
HIST_BINS = (256, 64, 1000)


def hist():
    """reference_hist.npz: the Histogram sample's kernel (binning() + binned_data()) executed by the reference DSL."""
    out = {}
    img = synth.image_np("float32", HIST_SHAPE[1], HIST_SHAPE[0], seed=9, scale=254.99)
    for nb in HIST_BINS:
        out[f"hist_{nb}"] = O.ref_sample_histogram_f32(img, nb)
    np.savez_compressed(os.path.join(HERE, "reference_hist.npz"), **out)
    print("wrote", len(out), "arrays to reference_hist.npz")


RGBA_SHAPE = (37, 53)


def rgba():
    """reference_rgba.npz: the RGBA samples' kernels (Kernel<uchar4>) executed by the reference DSL."""
    out = {}
    img = np.ascontiguousarray(synth.image_np("uint8", RGBA_SHAPE[1] * 4, RGBA_SHAPE[0], seed=11).reshape(RGBA_SHAPE[0], RGBA_SHAPE[1], 4))
    for b in (A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT):
        for sz in (3, 5):
            out[f"gauss_rgba_{sz}_{b}"] = O.ref_gaussian_rgba(img, M.GAUSS[sz], b)
        out[f"laplace_rgba_3_{b}"] = O.ref_laplace_rgba(img, M.LAPLACE3, b)
        out[f"laplace_rgba_5_{b}"] = O.ref_laplace_rgba(img, M.LAPLACE5, b)
        out[f"dilate_rgba_3_{b}"] = O.ref_dilate_rgba(img, 3, 3, b)
        out[f"box_rgba_5_{b}"] = O.ref_box_rgba(img, 5, 5, b)
    np.savez_compressed(os.path.join(HERE, "reference_rgba.npz"), **out)
    print("wrote", len(out), "arrays to reference_rgba.npz")


INTERP_SHAPE = (29, 37)
INTERP_TARGETS = [(58, 74), (20, 19), (41, 50)]   # 2x up, a ragged down-scale, a non-integer up-scale


def interp():
    """reference_interp.npz: a copy kernel through an interpolating Accessor (B5 / CF / L3; dsl/image.hpp:424-528) executed
    by the reference DSL -- the Scaling sample's operator with the three wide modes."""
    out = {}
    img = (synth.image_np("float32", INTERP_SHAPE[1], INTERP_SHAPE[0], seed=5) * 255).astype(np.float32)
    for mode, name in ((A.INTERP_B5, "b5"), (A.INTERP_CF, "cf"), (A.INTERP_L3, "l3")):
        for oh, ow in INTERP_TARGETS:
            out[f"{name}_{oh}x{ow}"] = O.ref_interp_f32(img, ow, oh, mode)
    np.savez_compressed(os.path.join(HERE, "reference_interp.npz"), **out)
    print("wrote", len(out), "arrays to reference_interp.npz")


if __name__ == "__main__":
    if "--only-interp" in sys.argv:
        interp()
    elif "--only-hist" in sys.argv:
        hist()
    elif "--only-rgba" in sys.argv:
        rgba()
    else:
        main()
        hist()
        rgba()
        interp()
