"""Operator cases shared by the CPU (oracle vs reference / golden) and GPU (product vs oracle) tests.

Each local-operator case is (golden key, input name, LocalSpec); the golden arrays were produced by
the compiled reference DSL (tests/golden/generate.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hipacc_b200 import _abi as A, masks as M, specs as S, synth  # noqa: E402

GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_dsl.npz")
BMODES = [A.CLAMP, A.REPEAT, A.MIRROR, A.CONSTANT]
U8_SHAPE = (45, 70)
F32_SHAPE = (39, 66)
HARRIS_SHAPE = (96, 128)
PYR_CASES = [(64, 96, 4, 5), (50, 77, 3, 3)]
RGBA_SHAPE = (37, 53)


def rgba_image(h, w, seed=11):
    """uchar4 image as uint8 [h, w, 4] (channels interleaved)"""
    return np.ascontiguousarray(synth.image_np("uint8", w * 4, h, seed=seed).reshape(h, w, 4))


HIST_SHAPE = (61, 83)
HIST_BINS = (256, 64, 1000)

_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = dict(np.load(GOLDEN_PATH))
    return _golden


def inputs():
    return {
        "u8": synth.image_np("uint8", U8_SHAPE[1], U8_SHAPE[0], seed=1),
        "f32": synth.image_np("float32", F32_SHAPE[1], F32_SHAPE[0], seed=2),
        "f255": synth.image_np("float32", F32_SHAPE[1], F32_SHAPE[0], seed=3, scale=255.0),
        "harris": synth.blocks_np(HARRIS_SHAPE[1], HARRIS_SHAPE[0], seed=5),
    }


def local_cases():
    """-> list of (golden_key, input_name, LocalSpec)"""
    out = []
    for b in BMODES:
        for sz in (3, 5, 7):
            out.append((f"gauss_u8_{sz}_{b}", "u8", S.gaussian_blur(M.GAUSS[sz], b)))
        for name, m in (("sobel3x", M.SOBEL3_X), ("sobel3y", M.SOBEL3_Y), ("sobel5x", M.SOBEL5_X)):
            out.append((f"{name}_u8_{b}", "u8", S.sobel_u8(m, b)))
            out.append((f"{name}_f32_{b}", "f32", S.domain_reduce_f32(m.astype(np.float32), b)))
        for name, m in (("lap3", M.LAPLACE3), ("lap3n4", M.LAPLACE3_4N), ("lap5", M.LAPLACE5)):
            out.append((f"{name}_u8_{b}", "u8", S.laplace_u8(m, b)))
            out.append((f"{name}_f32_{b}", "f32", S.domain_reduce_f32(m.astype(np.float32), b)))
        out.append((f"gauss5_f32_{b}", "f32", S.convolve_f32(M.GAUSS5, b)))
        out.append((f"dilate_u8_{b}", "u8", S.minmax_u8(5, 3, True, b)))
        out.append((f"erode_u8_{b}", "u8", S.minmax_u8(3, 5, False, b)))
        out.append((f"box_u8_{b}", "u8", S.box_blur_u8(5, 5, b)))
    return out


def case_id(c):
    return c[0]


# known-answer vectors produced from the reference DSL (W=6, H=4, v = 10*y+x+1)
KAT_IMG = np.array([[10 * y + x + 1 for x in range(6)] for y in range(4)], dtype=np.uint8)
KAT_SUM3 = {
    A.CLAMP: [42, 48, 57, 66, 75, 81, 102, 108, 117, 126, 135, 141, 192, 198, 207, 216, 225, 231, 252, 258, 267, 276, 285, 291],
    A.REPEAT: [147, 138, 147, 156, 165, 156, 117, 108, 117, 126, 135, 126, 207, 198, 207, 216, 225, 216, 177, 168, 177, 186, 195, 186],
    A.CONSTANT: [26, 42, 48, 54, 60, 42, 69, 108, 117, 126, 135, 93, 129, 198, 207, 216, 225, 153, 106, 162, 168, 174, 180, 122],
}
KAT_SUM3[A.MIRROR] = KAT_SUM3[A.CLAMP]
KAT_SUM5 = {
    A.CLAMP: [190, 205, 225, 250, 270, 285, 340, 355, 375, 400, 420, 435, 490, 505, 525, 550, 570, 585, 640, 655, 675, 700, 720, 735],
    A.MIRROR: [245, 255, 275, 300, 320, 330, 345, 355, 375, 400, 420, 430, 495, 505, 525, 550, 570, 580, 595, 605, 625, 650, 670, 680],
    A.REPEAT: [485, 480, 475, 500, 495, 490, 535, 530, 525, 550, 545, 540, 385, 380, 375, 400, 395, 390, 435, 430, 425, 450, 445, 440],
}
KAT_TAP_M2P2 = {  # single tap in(-2,+2)
    A.CLAMP: [21, 21, 21, 22, 23, 24, 31, 31, 31, 32, 33, 34, 31, 31, 31, 32, 33, 34, 31, 31, 31, 32, 33, 34],
    A.MIRROR: [22, 21, 21, 22, 23, 24, 32, 31, 31, 32, 33, 34, 32, 31, 31, 32, 33, 34, 22, 21, 21, 22, 23, 24],
    A.REPEAT: [25, 26, 21, 22, 23, 24, 35, 36, 31, 32, 33, 34, 5, 6, 1, 2, 3, 4, 15, 16, 11, 12, 13, 14],
    A.CONSTANT: [0, 0, 21, 22, 23, 24, 0, 0, 31, 32, 33, 34] + [0] * 12,
}
# A.3 interpolation: float image v=i (row-major index)
KAT_NN_8x4 = [9, 11, 13, 15, 25, 27, 29, 31]
KAT_LF_4x2_to_8x4 = [9, 9.5, 10.5, 11.5, 12.5, 13.5, 14.5, 15, 13, 13.5, 14.5, 15.5, 16.5, 17.5, 18.5, 19,
                     21, 21.5, 22.5, 23.5, 24.5, 25.5, 26.5, 27, 25, 25.5, 26.5, 27.5, 28.5, 29.5, 30.5, 31]
KAT_NN_7x5 = [8, 10, 12, 22, 24, 26]


def sum_domain_spec(size, boundary, const=0.0):
    """reduce(Domain size x size all ones, SUM, in(dom)) -> int (appendix A.1)"""
    return S.LocalSpec(size, size, A.REDUCE_DOMAIN, A.SUM, A.TAP_IN, A.S32, None,
                       np.ones((size, size), np.uint8), boundary, const, A.EPI_CAST, (0, 0, 0), A.S32)


def single_tap_spec(dx, dy, size, boundary, const=0.0):
    """in(dx,dy) expressed as a Domain with a single non-zero tap"""
    dom = np.zeros((size, size), np.uint8)
    dom[size // 2 + dy, size // 2 + dx] = 1
    return S.LocalSpec(size, size, A.REDUCE_DOMAIN, A.SUM, A.TAP_IN, A.S32, None, dom, boundary, const,
                       A.EPI_CAST, (0, 0, 0), A.U8)


# tests/golden/reference_interp.npz (generate.py: INTERP_SHAPE / INTERP_TARGETS, kept equal)
INTERP_SHAPE = (29, 37)
INTERP_TARGETS = [(58, 74), (20, 19), (41, 50)]
