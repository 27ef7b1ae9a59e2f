"""A numpy model of the PACKED integer arithmetic of harris_fused3_kernel (hipacc_b200/csrc/hb_harris.cu, version 3),
checked against the oracle's nine-kernel pipeline on the CPU.  It restates exactly what the kernel does with its registers --
16-bit pairs in 32-bit words, wrap-around uint32 adds, the 8192 bias of the dx*dy plane, the two dp2a coefficient words --
so a carry between the halves of a pair, a wrong coefficient byte or a bias that does not cancel shows up here without a GPU.
Interior pixels only (>= 3 px from the image edge): the CLAMP fix-up of the border tiles is index logic, not arithmetic."""
import numpy as np
import pytest

from hipacc_b200 import masks as M, synth

U32 = np.uint32
XY_BIAS = 8192


def _dp2a(a, coef_lo, coef_hi, acc):
    """dp2a.{lo,hi}.u32.u32: acc + a.lo16 * coef_lo + a.hi16 * coef_hi on uint32 words (wrap-around like the hardware)"""
    lo, hi = (a & U32(0xFFFF)).astype(np.uint64), (a >> U32(16)).astype(np.uint64)
    return ((acc.astype(np.uint64) + lo * coef_lo + hi * coef_hi) & 0xFFFFFFFF).astype(U32)


def _model(img, k, threshold):
    a = img.astype(np.int32)
    h, w = a.shape
    # ---- stage B: per input row D = a[i+1] - a[i-1], S = a[i-1] + a[i] + a[i+1] (two IDP.4A per position and row)
    D = np.zeros_like(a); S = np.zeros_like(a)
    D[:, 1:-1] = a[:, 2:] - a[:, :-2]
    S[:, 1:-1] = a[:, :-2] + a[:, 1:-1] + a[:, 2:]
    dx6 = np.zeros_like(a); dy6 = np.zeros_like(a)
    dx6[1:-1] = D[:-2] + D[1:-1] + D[2:]          # one three-input add
    dy6[1:-1] = S[2:] - S[:-2]                     # one subtraction
    qx = np.trunc(dx6 / 6).astype(np.int32)        # C truncating division
    qy = np.trunc(dy6 / 6).astype(np.int32)
    # packed products: position pairs (even x, odd x); the odd position's quotients are scaled by 256
    assert w % 2 == 0
    x0, x1, y0, y1 = qx[:, 0::2], qx[:, 1::2] << 8, qy[:, 0::2], qy[:, 1::2] << 8
    to_u32 = lambda v: (v.astype(np.int64) & 0xFFFFFFFF).astype(U32)   # noqa: E731  (int32 register contents as uint32)
    pxx = to_u32(x1 * x1 + x0 * x0)
    pyy = to_u32(y1 * y1 + y0 * y0)
    pxy = to_u32(x1.astype(np.int64) * y1 + (x0.astype(np.int64) * y0 + XY_BIAS * 65537))
    # every half is the unsigned 14-bit plane value the kernel stores
    for pk, lo_want, hi_want in ((pxx, qx[:, 0::2] ** 2, qx[:, 1::2] ** 2), (pyy, qy[:, 0::2] ** 2, qy[:, 1::2] ** 2),
                                 (pxy, qx[:, 0::2] * qy[:, 0::2] + XY_BIAS, qx[:, 1::2] * qy[:, 1::2] + XY_BIAS)):
        assert np.array_equal(pk & U32(0xFFFF), lo_want.astype(U32)) and np.array_equal(pk >> U32(16), hi_want.astype(U32))
        assert int((pk & U32(0xFFFF)).max()) < 1 << 14 and int((pk >> U32(16)).max()) < 1 << 14
    # ---- stage C: vertical a + 2b + c on the packed words (uint32 wrap-around adds), then the horizontal [1 2 1] by dp2a
    G = []
    for pl, pk in enumerate((pxx, pyy, pxy)):
        V = np.zeros_like(pk)
        V[1:-1] = (pk[1:-1] * U32(2) + pk[:-2]) + pk[2:]
        assert int((V & U32(0xFFFF)).max()) < 1 << 16        # trivially true for a mask; the point is the next line:
        want_lo = (pk[:-2] & U32(0xFFFF)).astype(np.int64) + 2 * (pk[1:-1] & U32(0xFFFF)).astype(np.int64) + (pk[2:] & U32(0xFFFF)).astype(np.int64)
        assert np.array_equal((V[1:-1] & U32(0xFFFF)).astype(np.int64), want_lo), "carry out of the low half of a packed pair"
        init = U32((-16 * XY_BIAS) & 0xFFFFFFFF) if pl == 2 else U32(0)
        acc0 = np.full(V[:, 1:-1].shape, init, U32)
        # word j holds pixels (2j, 2j+1).  Even pixel 2j: v[2j-1] + 2 v[2j] + v[2j+1] = word j-1 with (0, 1) | word j with (2, 1)
        # (coefficient word ce = 0x01020100); odd pixel 2j+1: word j with (1, 2) | word j+1 with (1, 0) (cw = 0x00010201)
        even = _dp2a(V[:, 1:-1], 2, 1, _dp2a(V[:, :-2], 0, 1, acc0))
        odd = _dp2a(V[:, 2:], 1, 0, _dp2a(V[:, 1:-1], 1, 2, acc0))
        g = np.zeros((h, w), np.int64)
        g[:, 2:-2:2] = even.astype(np.int32)
        g[:, 3:-2:2] = odd.astype(np.int32)
        G.append(g)
    # ---- response: >> 4 on the non-negative planes, |.| >> 4 on dx*dy (only its square is used), separately rounded floats
    x, y, xy = G[0] >> 4, G[1] >> 4, np.abs(G[2]) >> 4
    det = (x * y - xy * xy).astype(np.int32).astype(np.float32)
    s = (x + y).astype(np.float32)
    tr = (np.float32(k) * s) * s
    return ((det - tr) > np.float32(threshold)).astype(np.uint8), G


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_packed_arithmetic_model_equals_the_oracle_pipeline(oracle, seed):
    from test_oracle import _harris_extreme_images
    imgs = [synth.image_np("uint8", 96, 64, seed=seed), synth.blocks_np(96, 64, seed=seed)] + (_harris_extreme_images() if seed == 1 else [])
    for n, img in enumerate(imgs):
        want, gx, gy, gxy = oracle.harris(img, return_intermediates=True)
        got, G = _model(img, M.HARRIS_K, M.HARRIS_THRESHOLD)
        c = (slice(3, -3), slice(4, -4))
        # the smoothed planes: trunc(sum / 16) of the reference == the model's shifts
        np.testing.assert_array_equal((G[0] >> 4)[c], gx[c], err_msg=f"image {n}: xx plane")
        np.testing.assert_array_equal((G[1] >> 4)[c], gy[c], err_msg=f"image {n}: yy plane")
        np.testing.assert_array_equal((np.abs(G[2]) >> 4)[c], np.abs(gxy[c].astype(np.int64)), err_msg=f"image {n}: |xy| plane")
        np.testing.assert_array_equal(got[c], want[c], err_msg=f"image {n}: corners")
