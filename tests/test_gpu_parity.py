"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
  * the committed golden outputs of the compiled reference DSL (tests/golden/),
  * the CPU oracle (oracle/emit_cpu.cpp) on the same seeded inputs,
  * the compiled reference itself (oracle/_ref) when the prebuilt library travelled with the repo,
  * size-independent properties at BASELINE.json's full sizes.
Bar: bit-exact for uchar / int / index work AND for float stencils (separately rounded mul/add);
1e-5 relative for kernels with expf / division chains; float SUM within 1e-5 of the float64 sum.
"""
import os

import numpy as np
import pytest

import cases
from hipacc_b200 import _abi as A, masks as M, specs as S, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(hb):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    hb.init(0)
    return torch.device("cuda:0")


def to_dev(hb, a, dev, padded=True):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if not padded or a.ndim == 3:   # uchar4 images [H, W, 4]: dense
        return t.to(dev)
    img = hb.empty_image(A.NUMPY_DTYPE[a.dtype.name], a.shape[1], a.shape[0], device=dev)
    img.copy_(t)
    return img


def to_np(t):
    return t.cpu().numpy()


# ------------------------------------------------------------------ local operators
@pytest.mark.parametrize("case", cases.local_cases(), ids=cases.case_id)
def test_local_vs_golden_and_oracle(hb, oracle, dev, case):
    key, inp, spec = case
    img = cases.inputs()[inp]
    got = to_np(hb.local_op(spec, to_dev(hb, img, dev)))
    np.testing.assert_array_equal(got, cases.golden()[key])
    np.testing.assert_array_equal(got, oracle.local_op(spec, img))


BIG = (391, 517)  # (h, w): several tiles, ragged right / bottom edges


def _big_specs():
    out = []
    for b in cases.BMODES + [A.UNDEFINED]:
        out += [("gauss5_u8", "uint8", S.gaussian_blur(M.GAUSS5, b)), ("gauss3_u8", "uint8", S.gaussian_blur(M.GAUSS3, b)),
                ("gauss7_u8", "uint8", S.gaussian_blur(M.GAUSS7, b)),
                ("sobel3x_f32", "float32", S.domain_reduce_f32(M.SOBEL3_X.astype(np.float32), b)),
                ("sobel3y_f32", "float32", S.domain_reduce_f32(M.SOBEL3_Y.astype(np.float32), b)),
                ("laplace3_f32", "float32", S.domain_reduce_f32(M.LAPLACE3.astype(np.float32), b)),
                ("gauss5_f32", "float32", S.convolve_f32(M.GAUSS5, b)), ("gauss7_f32", "float32", S.convolve_f32(M.GAUSS7, b)),
                ("sobel5_u8_s32", "uint8", S.sobel_u8(M.SOBEL5_X, b)), ("laplace5_u8", "uint8", S.laplace_u8(M.LAPLACE5, b)),
                ("dilate_u8", "uint8", S.minmax_u8(3, 3, True, b)), ("erode5x3_u8", "uint8", S.minmax_u8(5, 3, False, b)),
                ("box7_u8", "uint8", S.box_blur_u8(7, 7, b)), ("harris_dx", "uint8", S.harris_deriv(M.HARRIS_DX))]
    return [(f"{n}_{A.BOUNDARY_NAMES[s.boundary]}", dt, s) for n, dt, s in out]


@pytest.mark.parametrize("case", _big_specs(), ids=lambda c: c[0])
@pytest.mark.parametrize("padded", [True, False], ids=["pitch256", "dense"])
def test_local_multi_tile_vs_oracle(hb, oracle, dev, case, padded):
    _, dt, spec = case
    if spec.boundary == A.UNDEFINED:
        pytest.skip("UNDEFINED reads are unspecified at the border; covered by the interior test below")
    img = synth.image_np(dt, BIG[1], BIG[0], seed=21)
    if spec.boundary == A.CONSTANT:
        spec.boundary_const = 7
    got = to_np(hb.local_op(spec, to_dev(hb, img, dev, padded)))
    np.testing.assert_array_equal(got, oracle.local_op(spec, img))


def test_undefined_boundary_interior_matches(hb, oracle, dev):
    img = synth.image_np("float32", BIG[1], BIG[0], seed=22)
    spec = S.convolve_f32(M.GAUSS5, A.UNDEFINED)
    got = to_np(hb.local_op(spec, to_dev(hb, img, dev)))
    np.testing.assert_array_equal(got[2:-2, 2:-2], oracle.local_op(S.convolve_f32(M.GAUSS5, A.CLAMP), img)[2:-2, 2:-2])


@pytest.mark.parametrize("size", [(1, 3), (3, 1), (5, 3), (9, 9), (13, 13), (1, 1)])
def test_generic_mask_sizes(hb, oracle, dev, size):
    sx, sy = size
    rng = np.random.default_rng(3)
    m = rng.random((sy, sx), dtype=np.float32)
    m[rng.random((sy, sx)) < 0.2] = 0.0
    img = synth.image_np("float32", 150, 97, seed=23)
    for spec in (S.convolve_f32(m, A.MIRROR), S.domain_reduce_f32(m, A.REPEAT), S.domain_reduce_f32(m, A.CLAMP, A.MAX),
                 S.convolve_f32(m, A.CLAMP, A.MIN), S.domain_reduce_f32(m, A.CONSTANT, A.PROD, const=0.5)):
        if max(m.shape) // 2 >= min(img.shape) and spec.boundary == A.MIRROR:
            continue
        got = to_np(hb.local_op(spec, to_dev(hb, img, dev)))
        np.testing.assert_array_equal(got, oracle.local_op(spec, img))


@pytest.mark.parametrize("mode", [A.MIN, A.MAX, A.PROD])
@pytest.mark.parametrize("size", [3, 5, 7])
def test_tiled_general_variant_modes(hb, oracle, dev, mode, size):
    rng = np.random.default_rng(size)
    m = (rng.random((size, size), dtype=np.float32) + 0.5).astype(np.float32)
    m[0, 0] = 0.0
    img = synth.image_np("float32", 300, 200, seed=24)
    for spec in (S.domain_reduce_f32(m, A.MIRROR, mode), S.convolve_f32(m, A.CLAMP, mode)):
        got = to_np(hb.local_op(spec, to_dev(hb, img, dev)))
        np.testing.assert_array_equal(got, oracle.local_op(spec, img))


ROIS = [((30, 20, 5, 7), (30, 20, 11, 3)), ((200, 150, 100, 60), (200, 150, 0, 0)), ((129, 33, 3, 1), (129, 33, 250, 200)),
        ((517, 1, 0, 390), (517, 1, 0, 0)), ((1, 391, 516, 0), (1, 391, 3, 0))]


@pytest.mark.parametrize("roi", ROIS)
@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT])
def test_roi_and_crop_accessors(hb, oracle, dev, roi, b):
    ris, racc = roi
    if b in (A.MIRROR,) and min(racc[0], racc[1]) < 2:
        pytest.skip("single reflection with halo > window is undefined in the reference")
    img = synth.image_np("uint8", BIG[1], BIG[0], seed=25)
    base = synth.image_np("uint8", BIG[1], BIG[0], seed=26)
    spec = S.gaussian_blur(M.GAUSS5, b)
    spec.boundary_const = 9
    want = oracle.local_op(spec, img, out=base.copy(), roi_in=racc, roi_out=ris)
    got = to_np(hb.local_op(spec, to_dev(hb, img, dev), dst=to_dev(hb, base, dev), roi_in=racc, roi_out=ris))
    np.testing.assert_array_equal(got, want)  # pixels outside the iteration space untouched as well


def test_ghost_rows_make_strips_equal_the_whole(hb, oracle, dev):
    """Row-strip sharding: a strip with R ghost rows gives exactly the rows of the full result."""
    img = synth.image_np("float32", 300, 240, seed=27)
    spec = S.domain_reduce_f32(M.LAPLACE5.astype(np.float32), A.MIRROR)
    full = oracle.local_op(spec, img)
    R = 2
    for (y0, y1) in [(0, 80), (80, 160), (160, 240)]:
        g0, g1 = min(R, y0), min(R, 240 - y1)
        strip = np.ascontiguousarray(img[y0 - g0:y1 + g1])
        roi = (300, y1 - y0, 0, g0)
        out = hb.local_op(spec, to_dev(hb, strip, dev), roi_in=roi, roi_out=roi, ghost=(g0, g1))
        np.testing.assert_array_equal(to_np(out)[g0:g0 + y1 - y0], full[y0:y1])


def test_tiny_and_degenerate_images(hb, oracle, dev):
    for shape in [(1, 1), (1, 9), (7, 1), (2, 3), (5, 5), (33, 129)]:
        img = synth.image_np("uint8", shape[1], shape[0], seed=28)
        for b in (A.CLAMP, A.CONSTANT, A.REPEAT):
            spec = S.gaussian_blur(M.GAUSS3, b)
            np.testing.assert_array_equal(to_np(hb.local_op(spec, to_dev(hb, img, dev, padded=False))), oracle.local_op(spec, img))


def test_unsupported_combinations_fail_loudly(hb, dev):
    import torch
    img = torch.zeros((16, 16), dtype=torch.int32, device=dev)
    with pytest.raises(hb.HbError) as e:
        hb.local_op(S.convolve_f32(M.GAUSS3), img)
    assert e.value.status == A.HB_ERR_UNSUPPORTED and "no CPU fallback" in str(e.value)
    f = torch.zeros((16, 16), dtype=torch.float32, device=dev)
    with pytest.raises(hb.HbError):
        hb.local_op(S.convolve_f32(np.ones((15, 15), np.float32)), f)
    with pytest.raises(hb.HbError):
        hb.bilateral(f, 9, np.ones((9, 9), np.float32), 16)


# ------------------------------------------------------------------ bilateral
@pytest.mark.parametrize("size", [3, 5, 7, 13])
def test_bilateral(hb, oracle, dev, size):
    cm = M.bilateral_mask(size)
    u8 = synth.image_np("uint8", 300, 170, seed=31)
    got = to_np(hb.bilateral(to_dev(hb, u8, dev), size, cm, 16, A.CLAMP)).astype(np.int32)
    want = oracle.bilateral(u8, size, cm, 16, A.CLAMP).astype(np.int32)
    diff = np.abs(got - want)
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3   # rounding ties of (uchar)(p/d+0.5f) only
    f = synth.image_np("float32", 300, 170, seed=32, scale=255.0)
    got = to_np(hb.bilateral(to_dev(hb, f, dev), size, cm, 16, A.MIRROR))
    want = oracle.bilateral(f, size, cm, 16, A.MIRROR)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)


def test_bilateral_vs_golden(hb, dev):
    inp, g = cases.inputs(), cases.golden()
    for size in (3, 5, 13):
        got = to_np(hb.bilateral(to_dev(hb, inp["f255"], dev), size, M.bilateral_mask(size), 16, A.MIRROR))
        np.testing.assert_allclose(got, g[f"bilateral_f32_{size}"], rtol=1e-5, atol=0)
        gotu = to_np(hb.bilateral(to_dev(hb, inp["u8"], dev), size, M.bilateral_mask(size), 16, A.CLAMP)).astype(int)
        assert np.abs(gotu - g[f"bilateral_u8_{size}"].astype(int)).max() <= 1


# ------------------------------------------------------------------ point operators + interpolation
def test_point_ops(hb, oracle, dev):
    u8 = cases.inputs()["u8"]
    a = oracle.local_op(S.sobel_u8(M.SOBEL3_X), u8)
    b = oracle.local_op(S.sobel_u8(M.SOBEL3_Y), u8)
    got = to_np(hb.point_op(A.POINT_SOBEL_COMBINE, [to_dev(hb, a, dev), to_dev(hb, b, dev)], A.U8, p=(4, 0)))
    np.testing.assert_array_equal(got, cases.golden()["sobel_combine"])
    rng = np.random.default_rng(5)
    s1 = rng.integers(-127, 128, (77, 131)).astype(np.int16)
    s2 = rng.integers(-127, 128, (77, 131)).astype(np.int16)
    s3 = rng.integers(-16129, 16130, (77, 131)).astype(np.int16)
    for padded in (True, False):
        d1, d2, d3 = (to_dev(hb, x, dev, padded) for x in (s1, s2, s3))
        np.testing.assert_array_equal(to_np(hb.point_op(A.POINT_SQUARE, [d1], A.S16)), oracle.point_op(A.POINT_SQUARE, [s1], A.S16))
        np.testing.assert_array_equal(to_np(hb.point_op(A.POINT_MUL, [d1, d2], A.S16)), oracle.point_op(A.POINT_MUL, [s1, s2], A.S16))
        sq1, sq2 = np.abs(s3), np.abs(s3[::-1]).copy()
        np.testing.assert_array_equal(
            to_np(hb.point_op(A.POINT_HARRIS, [to_dev(hb, sq1, dev, padded), to_dev(hb, sq2, dev, padded), d3], A.U8, p=(0.04, 20000.0))),
            oracle.point_op(A.POINT_HARRIS, [sq1, sq2, s3], A.U8, p=(0.04, 20000.0)))
    f1 = synth.image_np("float32", 131, 77, seed=33)
    f2 = synth.image_np("float32", 131, 77, seed=34)
    for op in (A.POINT_SUB, A.POINT_ADD, A.POINT_BLEND, A.POINT_MUL):
        np.testing.assert_array_equal(to_np(hb.point_op(op, [to_dev(hb, f1, dev), to_dev(hb, f2, dev)], A.F32)),
                                      oracle.point_op(op, [f1, f2], A.F32))


@pytest.mark.parametrize("shape", [((64, 96), (32, 48)), ((39, 66), (19, 33)), ((50, 77), (25, 38)), ((128, 128), (64, 64))])
def test_interpolation(hb, oracle, dev, shape):
    (h, w), (ch, cw) = shape
    f = synth.image_np("float32", w, h, seed=35)
    nn = to_np(hb.point_op(A.POINT_COPY, [to_dev(hb, f, dev)], A.F32, (ch, cw), [A.INTERP_NN]))
    np.testing.assert_array_equal(nn, oracle.point_op(A.POINT_COPY, [f], A.F32, (ch, cw), [A.INTERP_NN]))
    lf = to_np(hb.point_op(A.POINT_COPY, [to_dev(hb, nn, dev)], A.F32, (h, w), [A.INTERP_LF]))
    np.testing.assert_array_equal(lf, oracle.point_op(A.POINT_COPY, [nn], A.F32, (h, w), [A.INTERP_LF]))


@pytest.mark.parametrize("mode", [A.INTERP_B5, A.INTERP_CF, A.INTERP_L3])
def test_wide_interpolation(hb, oracle, dev, mode):
    """B5 / CF / L3 at arbitrary scale factors: golden from the reference DSL, then float and integer images against the
    oracle.  B5 / CF are pure float arithmetic (bit-exact); L3's weights go through a double-precision sin (1e-5)."""
    G = cases
    g = np.load(os.path.join(os.path.dirname(cases.GOLDEN_PATH), "reference_interp.npz"))
    name = {A.INTERP_B5: "b5", A.INTERP_CF: "cf", A.INTERP_L3: "l3"}[mode]
    img = (synth.image_np("float32", G.INTERP_SHAPE[1], G.INTERP_SHAPE[0], seed=5) * 255).astype(np.float32)
    def close(got, want, what):
        if mode == A.INTERP_L3:
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-4, err_msg=what)
        else:
            np.testing.assert_array_equal(got, want, err_msg=what)
    for oh, ow in G.INTERP_TARGETS:
        close(to_np(hb.point_op(A.POINT_COPY, [to_dev(hb, img, dev)], A.F32, (oh, ow), [mode])), g[f"{name}_{oh}x{ow}"], f"golden {oh}x{ow}")
    f = synth.image_np("float32", 333, 211, seed=36)
    for oh, ow in ((422, 666), (97, 150), (211, 333), (500, 123)):
        close(to_np(hb.point_op(A.POINT_COPY, [to_dev(hb, f, dev)], A.F32, (oh, ow), [mode])),
              oracle.point_op(A.POINT_COPY, [f], A.F32, (oh, ow), [mode]), f"f32 {oh}x{ow}")
    u8 = synth.image_np("uint8", 200, 120, seed=37)
    got = to_np(hb.point_op(A.POINT_COPY, [to_dev(hb, u8, dev)], A.U8, (171, 333), [mode])).astype(np.int32)
    want = oracle.point_op(A.POINT_COPY, [u8], A.U8, (171, 333), [mode]).astype(np.int32)
    if mode == A.INTERP_L3:
        # a 1-ulp weight difference can move a value across an integer: at most 1 LSB, and only rarely
        assert np.abs(got - want).max() <= 1 and (got != want).mean() < 1e-3
    else:
        np.testing.assert_array_equal(got, want)


def test_interpolation_kat(hb, dev):
    img = np.arange(32, dtype=np.float32).reshape(4, 8)
    nn = to_np(hb.point_op(A.POINT_COPY, [to_dev(hb, img, dev)], A.F32, (2, 4), [A.INTERP_NN]))
    assert nn.ravel().tolist() == cases.KAT_NN_8x4
    lf = to_np(hb.point_op(A.POINT_COPY, [to_dev(hb, nn, dev)], A.F32, (4, 8), [A.INTERP_LF]))
    assert lf.ravel().tolist() == cases.KAT_LF_4x2_to_8x4


@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT])
def test_kat_appendix_a(hb, dev, b):
    k = to_dev(hb, cases.KAT_IMG, dev, padded=False)
    assert to_np(hb.local_op(cases.sum_domain_spec(3, b), k)).ravel().tolist() == cases.KAT_SUM3[b]
    if b in cases.KAT_SUM5:
        assert to_np(hb.local_op(cases.sum_domain_spec(5, b), k)).ravel().tolist() == cases.KAT_SUM5[b]
    assert to_np(hb.local_op(cases.single_tap_spec(-2, 2, 5, b), k)).ravel().tolist() == cases.KAT_TAP_M2P2[b]


# ------------------------------------------------------------------ global reductions
@pytest.mark.parametrize("shape", [(39, 66), (1, 1), (3, 1025), (1024, 1024), (517, 391)])
def test_reduce_minmaxsum(hb, oracle, dev, shape):
    f = synth.image_np("float32", shape[1], shape[0], seed=41)
    for padded in (True, False):
        mn, mx, sm = hb.reduce_minmaxsum(to_dev(hb, f, dev, padded))
        s64 = f.astype(np.float64).sum()
        assert np.float32(mn) == f.min() and np.float32(mx) == f.max()   # order independent: bit-exact
        assert abs(sm - s64) <= 1e-5 * abs(s64)                           # float SUM contract
    # determinism: same grid, same order -> same bits
    d = to_dev(hb, f, dev)
    assert hb.reduce_minmaxsum(d) == hb.reduce_minmaxsum(d)


def test_reduce_roi_and_generic_entry(hb, oracle, dev):
    f = synth.image_np("float32", 130, 100, seed=42)
    d = to_dev(hb, f, dev)
    roi = (50, 40, 7, 9)
    mn, mx, sm = hb.reduce_minmaxsum(d, roi)
    sub = f[9:49, 7:57]
    assert np.float32(mn) == sub.min() and np.float32(mx) == sub.max()
    assert abs(sm - sub.astype(np.float64).sum()) <= 1e-5 * sub.sum()
    assert hb.reduce(d, A.MIN) == f.min() and hb.reduce(d, A.MAX) == f.max()
    u8 = synth.image_np("uint8", 130, 100, seed=43)
    du = to_dev(hb, u8, dev)
    assert hb.reduce(du, A.MAX) == u8.max() and hb.reduce(du, A.MIN) == u8.min()
    assert hb.reduce(du, A.SUM) == np.uint8(u8.astype(np.uint64).sum() & 0xFF)   # uchar reduce(uchar,uchar) wraps
    s32 = synth.image_np("uint8", 130, 100, seed=44).astype(np.int32) - 100
    assert hb.reduce(to_dev(hb, s32, dev), A.SUM) == s32.sum()


def test_reduce_vs_golden(hb, dev):
    g = cases.golden()
    mn, mx, sm = hb.reduce_minmaxsum(to_dev(hb, cases.inputs()["f32"], dev))
    assert np.float32(mn) == g["reduce_min"][0] and np.float32(mx) == g["reduce_max"][0]
    assert abs(sm - float(g["reduce_sum"][0])) <= 1e-5 * abs(sm)


@pytest.mark.parametrize("dt", ["uint8", "int8", "int16", "int32"])
@pytest.mark.parametrize("shape,roi", [((39, 66), None), ((3, 1025), None), ((517, 391), (300, 200, 33, 17)), ((1024, 2048), None)])
def test_reduce_integer_types_all_modes(hb, dev, dt, shape, roi):
    """Integer images fold in the pixel type's modular arithmetic (data_t reduce(data_t, data_t), dsl/kernel.hpp:121-151):
    SUM and PROD wrap, so the vectorised kernel is bit-exact whatever its order."""
    raw = synth.image_np("uint8", shape[1], shape[0], seed=47)
    img = raw.astype(dt) if dt in ("uint8", "int8") else (raw.astype(np.int32) * 37 - 4000).astype(dt)
    d = to_dev(hb, img, dev)
    sub = img if roi is None else img[roi[3]:roi[3] + roi[1], roi[2]:roi[2] + roi[0]]
    bits = 8 * img.dtype.itemsize
    want_sum = np.array([int(sub.astype(np.int64).sum()) & ((1 << bits) - 1)], dtype=np.uint64).astype(f"uint{bits}").view(dt)[0]
    assert hb.reduce(d, A.SUM, roi) == want_sum
    assert hb.reduce(d, A.MIN, roi) == sub.min() and hb.reduce(d, A.MAX, roi) == sub.max()
    # PROD modulo 2^bits: odd values keep the product non-zero
    odd = (img | 1).astype(dt)
    prod = 1
    sub_odd = odd if roi is None else odd[roi[3]:roi[3] + roi[1], roi[2]:roi[2] + roi[0]]
    for v in sub_odd.ravel().tolist()[:20000]:
        prod = (prod * v) & ((1 << bits) - 1)
    small = np.ascontiguousarray(sub_odd.ravel()[:20000].reshape(1, -1))
    assert hb.reduce(to_dev(hb, small, dev), A.PROD) == np.array([prod], dtype=np.uint64).astype(f"uint{bits}").view(dt)[0]


def test_reduce_sum_sample_int_4096(hb, dev):
    """the Reduction_Sum sample's configuration: int 4096 x 4096 (samples-public/2_Global_Operators/Reduction_Sum)"""
    img = (synth.image_np("uint8", 4096, 4096, seed=48).astype(np.int32) - 90)
    want = np.array([int(img.astype(np.int64).sum()) & 0xFFFFFFFF], dtype=np.uint64).astype(np.uint32).view(np.int32)[0]
    assert hb.reduce(to_dev(hb, img, dev), A.SUM) == want


def test_reduce_float_prod(hb, dev):
    f = np.exp((synth.image_np("float32", 300, 211, seed=49).astype(np.float64) - 0.5) * 0.02).astype(np.float32)   # around 1: the product stays finite
    got = hb.reduce(to_dev(hb, f, dev), A.PROD)
    want = np.exp(np.log(f.astype(np.float64)).sum())
    assert abs(got - want) <= 1e-5 * abs(want)


def test_reduce_nan(hb, dev):
    """MIN / MAX ignore NaN pixels (fminf / fmaxf).  The reference has no thread-count independent result here: its DSL
    fold restarts after a NaN (min(a,b) = a < b ? a : b, dsl/math_functions.hpp:349-351), its OpenMP runtime does so per
    thread chunk -- the documented contract of this library is 'NaNs are skipped'."""
    f = synth.image_np("float32", 257, 131, seed=50)
    g = f.copy()
    g[0, 0] = np.nan
    g[77, 100] = np.nan
    g[130, 256] = np.nan
    mn, mx, sm = hb.reduce_minmaxsum(to_dev(hb, g, dev))
    assert np.float32(mn) == np.nanmin(g) and np.float32(mx) == np.nanmax(g) and np.isnan(sm)
    allnan = np.full((5, 9), np.nan, np.float32)
    mn, mx, _ = hb.reduce_minmaxsum(to_dev(hb, allnan, dev))
    assert mn == np.inf and mx == -np.inf


def test_reductions_on_two_streams_do_not_share_scratch(hb, dev):
    """per-(device, stream) scratch: reductions in flight on two streams each fold their own partials"""
    import torch
    a = synth.image_np("float32", 4096, 2048, seed=51)
    b = synth.image_np("float32", 4096, 2048, seed=52) * 3.0
    da, db = to_dev(hb, a, dev), to_dev(hb, b, dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    pa = torch.zeros((64, 4), dtype=torch.float32, device=dev)
    pb = torch.zeros((64, 4), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    for i in range(64):   # interleaved launches: with one shared ticket the last-CTA folds would mix
        hb.reduce_minmaxsum_async(da, pa[i], stream=s1)
        hb.reduce_minmaxsum_async(db, pb[i], stream=s2)
    torch.cuda.synchronize()
    for part, img in ((pa, a), (pb, b)):
        h = part.cpu().numpy()
        assert (h[:, 0] == img.min()).all() and (h[:, 1] == img.max()).all()
        sums = np.ascontiguousarray(h[:, 2:4]).view(np.float64).ravel()
        assert (sums == sums[0]).all() and abs(sums[0] - img.astype(np.float64).sum()) <= 1e-5 * img.astype(np.float64).sum()


def test_graph_capture_with_timing_enabled_and_captured_reduction(hb, oracle, dev):
    """hb_set_timing(1) must not break a capture (captured operators are not timed), and a captured reduction keeps
    its scratch: a larger reduction between capture and replay does not invalidate the graph."""
    import torch
    stream = torch.cuda.Stream(device=dev)
    f0, f1 = synth.image_np("float32", 640, 333, seed=53), synth.image_np("float32", 640, 333, seed=54)
    src = to_dev(hb, f0, dev)
    dst = torch.zeros_like(src)
    part = torch.zeros(4, dtype=torch.float32, device=dev)
    spec = S.domain_reduce_f32(M.LAPLACE3.astype(np.float32), A.MIRROR)
    hb.set_timing(True)
    try:
        with torch.cuda.stream(stream):
            with hb.Graph(stream) as g:   # first use of the reduction on this stream happens INSIDE the capture
                hb.local_op(spec, src, dst=dst, stream=stream)
                hb.reduce_minmaxsum_async(dst, part, stream=stream)
            big = to_dev(hb, synth.image_np("float32", 8192, 2048, seed=55), dev)
            hb.reduce_minmaxsum(big, stream=stream)   # a larger grid in between
            for img in (f1, f0):
                src.copy_(torch.from_numpy(img).to(dev))
                stream.synchronize()
                g.launch()
                stream.synchronize()
                want = oracle.local_op(spec, img)
                np.testing.assert_array_equal(to_np(dst), want)
                h = part.cpu().numpy()
                assert h[0] == want.min() and h[1] == want.max()
            g.destroy()
            hb.local_op(spec, src, dst=dst, stream=stream)   # outside a capture the operator is timed again
            assert hb.last_kernel_ms() > 0.0
    finally:
        hb.set_timing(False)


def test_first_tap_initialises_and_holes_skip_non_finite_pixels(hb, oracle, dev):
    """dsl/kernel.hpp:250,279: the first visited tap initialises the accumulator and Domain holes are never read --
    an inf / NaN pixel under a hole must not reach the result (dense and 256-byte pitched rows: TMA and tiled paths)."""
    f = synth.image_np("float32", 300, 200, seed=56)
    f[50, 60] = np.inf
    f[120, 7] = np.nan
    f[0, 0] = -np.inf
    for m in (M.SOBEL3_X, M.LAPLACE3, np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]])):
        spec = S.domain_reduce_f32(m.astype(np.float32), A.CLAMP)
        want = oracle.local_op(spec, f)
        for padded in (True, False):
            got = to_np(hb.local_op(spec, to_dev(hb, f, dev, padded)))
            np.testing.assert_array_equal(got, want)


def test_reduce_async_leaves_the_scalar_on_the_device_and_replays_in_a_graph(hb, oracle, dev):
    """hb_reduce_async (integer images, every mode; float PROD): same scalar as the blocking call, capturable."""
    import torch
    rng = np.random.default_rng(77)
    out = torch.zeros(2, dtype=torch.int32, device=dev)
    for dt, lo, hi in (("int32", -1000, 1000), ("uint8", 0, 256), ("int16", -300, 300)):
        a = rng.integers(lo, hi, size=(157, 333)).astype(dt)
        d = to_dev(hb, a, dev)
        for mode, want in ((A.SUM, a.astype(np.int64).sum()), (A.MIN, a.min()), (A.MAX, a.max())):
            hb.reduce_async(d, mode, out)
            got = int(out[0].item())
            blocking = hb.reduce(d, mode)
            assert np.array(got).astype(dt) == np.array(want).astype(dt) == blocking, (dt, mode, got, want, blocking)
    d = to_dev(hb, rng.integers(-5, 6, size=(300, 500)).astype("int32"), dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        hb.reduce_async(d, A.SUM, out, stream=stream)
        torch.cuda.synchronize()
        with hb.Graph(stream) as g:
            hb.reduce_async(d, A.SUM, out, stream=stream)
        for k in range(3):
            d.fill_(k + 1)
            g.launch()
            torch.cuda.synchronize()
            assert int(out[0].item()) == (k + 1) * 300 * 500
        g.destroy()
    f = to_dev(hb, (1.0 + rng.random((40, 50)) * 1e-3).astype("float32"), dev)
    outf = torch.zeros(1, dtype=torch.float64, device=dev)
    hb.reduce_async(f, A.PROD, outf)
    assert abs(float(outf.item()) / float(hb.reduce(f, A.PROD)) - 1) < 1e-5
    with pytest.raises(RuntimeError):
        hb.reduce_async(f, A.SUM, outf)     # float SUM / MIN / MAX: the fused hb_reduce_minmaxsum_f32_async


# ------------------------------------------------------------------ Harris
@pytest.mark.parametrize("shape", [cases.HARRIS_SHAPE, (200, 333), (33, 129), (5, 7)])
def test_harris_fused_and_unfused(hb, oracle, dev, shape):
    img = synth.blocks_np(shape[1], shape[0], seed=5)
    want, gx, gy, gxy = oracle.harris(img, return_intermediates=True)
    d = to_dev(hb, img, dev)
    out_u, dgx, dgy, dgxy = hb.harris_unfused(d)
    np.testing.assert_array_equal(to_np(dgx), gx)
    np.testing.assert_array_equal(to_np(dgy), gy)
    np.testing.assert_array_equal(to_np(dgxy), gxy)
    np.testing.assert_array_equal(to_np(out_u), want)
    np.testing.assert_array_equal(to_np(hb.harris(d)), want)          # fused kernel == 9-kernel pipeline
    if shape == cases.HARRIS_SHAPE:
        np.testing.assert_array_equal(want, cases.golden()["harris_out"])
        assert 0 < want.sum() < want.size // 4


def test_harris_noise_image(hb, oracle, dev):
    img = synth.image_np("uint8", 260, 140, seed=6)       # white noise: every stage far from trivial
    np.testing.assert_array_equal(to_np(hb.harris(to_dev(hb, img, dev))), oracle.harris(img))


def test_harris_extreme_gradients(hb, oracle, dev):
    """Binary noise, diagonal steps, checkerboards and stripes: the Sobel sums and the biased dx*dy plane of the fused kernel
    at their extremes (tests/test_oracle.py::test_harris_product_bound_behind_the_biased_xy_plane states the bound)."""
    from test_oracle import _harris_extreme_images
    for k, img in enumerate(_harris_extreme_images()):
        big = np.tile(img, (3, 3))     # several tiles: interior (TMA-staged) and border tiles
        np.testing.assert_array_equal(to_np(hb.harris(to_dev(hb, big, dev))), oracle.harris(big), err_msg=f"image {k}")


@pytest.mark.parametrize("R", [2, 3])
def test_harris_strips_with_ghost_rows_equal_the_whole(hb, oracle, dev, R):
    """C4 sharding: each strip + R >= 2 ghost rows of real neighbour data gives exactly its rows of the full result
    (the 5x5 receptive field of the fused pipeline; CLAMP of image AND intermediates only at the global edge)."""
    H, W = 230, 300
    img = synth.image_np("uint8", W, H, seed=8)
    full = oracle.harris(img)
    for n in (2, 3, 5):
        for r in range(n):
            y0, y1 = H * r // n, H * (r + 1) // n
            g0, g1 = min(R, y0), min(R, H - y1)
            strip = np.ascontiguousarray(img[y0 - g0:y1 + g1])
            roi = (W, y1 - y0, 0, g0)
            out = hb.harris(to_dev(hb, strip, dev), roi=roi, ghost=(g0, g1))
            np.testing.assert_array_equal(to_np(out)[g0:g0 + y1 - y0], full[y0:y1])


@pytest.mark.parametrize("ox", [16, 32, 5])
def test_harris_roi_offsets_take_the_tma_or_the_loader_path(hb, oracle, dev, ox):
    """ROI accessors: CLAMP applies at the ROI edge (dsl/image.hpp:574-580), so the ROI result equals the pipeline run on
    the cropped image.  A 16-byte aligned ROI origin keeps the TMA-staged kernel (interior tiles several tiles away from
    every ROI edge); any other origin takes the all-threads loader of version 2."""
    img = synth.image_np("uint8", 900, 220, seed=21)
    W, H, oy = 700, 170, 9
    d = to_dev(hb, img, dev)
    out = hb.harris(d, roi=(W, H, ox, oy))
    np.testing.assert_array_equal(to_np(out)[oy:oy + H, ox:ox + W], oracle.harris(np.ascontiguousarray(img[oy:oy + H, ox:ox + W])))


@pytest.mark.parametrize("version", ["1", "2"])
def test_harris_earlier_kernel_versions_stay_exact(version):
    """Version 2 of the fused kernel is the fallback for images TMA cannot address (and version 1 an A/B knob): the knob
    is read once per process, so they run in a child process against the oracle."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np, torch; sys.path.insert(0, %r); import hipacc_b200 as hb; from hipacc_b200 import synth; from oracle import oracle as O; "
            "hb.init(0); O.emit_lib(); o = O; "
            "ok = all(np.array_equal(hb.harris(torch.from_numpy(i).cuda()).cpu().numpy(), o.harris(i)) "
            "for i in (synth.image_np('uint8', 517, 300, seed=3), synth.blocks_np(640, 200, seed=5), synth.image_np('uint8', 7, 5, seed=1))); "
            "print('HARRIS_OK' if ok else 'HARRIS_DIFF')") % root
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, "HB_HARRIS_VERSION": version}, capture_output=True, text=True, timeout=300)
    assert "HARRIS_OK" in r.stdout, r.stdout + r.stderr


# ------------------------------------------------------------------ randomised sweep
def test_randomised_sweep_gpu_vs_oracle(hb, oracle, dev):
    """Seeded fuzz over shapes (1 px up to several tiles, widths that break every alignment assumption), boundary modes,
    padded / dense rows and operator families: the CUDA path must equal the oracle bit for bit."""
    import torch
    rng = np.random.default_rng(4242)
    ops = ("gauss3", "gauss5", "gauss7", "dom3", "dom5", "conv5", "conv7", "max3", "min5", "box3", "sobel5i", "lap5i", "rgba5", "hist")
    for case in range(56):
        op = ops[case % len(ops)]
        b = int(rng.choice([A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT]))
        size = 5 if op in ("sobel5i", "lap5i", "rgba5", "hist") else int(op[-1])
        lo = 1 if b in (A.CLAMP, A.CONSTANT) else size // 2 + 1
        h = int(rng.integers(lo, 150))
        w = int(rng.choice([rng.integers(lo, 40), rng.integers(100, 700), 128, 256, 257, 511]))
        w = max(w, lo)
        padded = bool(rng.integers(0, 2))
        seed = 3000 + case
        msg = f"case {case}: {op} {h}x{w} boundary {b} padded {padded}"
        if op == "hist":
            img = synth.image_np("float32", w, h, seed=seed, scale=254.99)
            nb = int(rng.choice([16, 256, 1000]))
            np.testing.assert_array_equal(hb.binning(to_dev(hb, img, dev, padded), nb), oracle.binning(img, nb), err_msg=msg)
            continue
        if op == "rgba5":
            img = cases.rgba_image(h, w, seed=seed)
            spec = S.gaussian_blur(M.GAUSS5, b)
            np.testing.assert_array_equal(to_np(hb.local_op(spec, to_dev(hb, img, dev))), oracle.local_op_x4(spec, img), err_msg=msg)
            continue
        if op.startswith("gauss"):
            img, spec = synth.image_np("uint8", w, h, seed=seed), S.gaussian_blur(M.GAUSS[size], b)
        elif op in ("dom3", "dom5"):
            img = synth.image_np("float32", w, h, seed=seed)
            spec = S.domain_reduce_f32((M.SOBEL3_Y if op == "dom3" else M.SOBEL5_X).astype(np.float32), b)
        elif op in ("conv5", "conv7"):
            img, spec = synth.image_np("float32", w, h, seed=seed), S.convolve_f32(M.GAUSS[size], b)
        elif op in ("max3", "min5"):
            img, spec = synth.image_np("uint8", w, h, seed=seed), S.minmax_u8(size, size, op == "max3", b)
        elif op == "box3":
            img, spec = synth.image_np("uint8", w, h, seed=seed), S.box_blur_u8(3, 3, b)
        elif op == "sobel5i":
            img, spec = synth.image_np("uint8", w, h, seed=seed), S.sobel_u8(M.SOBEL5_Y, b)      # separable integer mask
        else:
            img, spec = synth.image_np("uint8", w, h, seed=seed), S.laplace_u8(M.LAPLACE5, b)    # not separable
        got = to_np(hb.local_op(spec, to_dev(hb, img, dev, padded)))
        np.testing.assert_array_equal(got, oracle.local_op(spec, img), err_msg=msg)


def test_randomised_sweep_pipelines_gpu_vs_oracle(hb, oracle, dev):
    """Seeded fuzz over the multi-kernel paths: fused Harris, pyramid traversals (even sizes take the fused level
    kernels, odd ones the general mapping), interpolating point operators, point arithmetic and reductions."""
    import torch
    rng = np.random.default_rng(777)
    for case in range(10):                                   # fused Harris vs the 9-kernel oracle pipeline
        h, w = int(rng.integers(1, 140)), int(rng.choice([rng.integers(1, 60), rng.integers(100, 520), 128, 129]))
        img = synth.blocks_np(w, h, seed=500 + case)
        np.testing.assert_array_equal(to_np(hb.harris(to_dev(hb, img, dev, bool(case & 1)))), oracle.harris(img), err_msg=f"harris {h}x{w}")
    for case in range(8):                                    # pyramids
        depth = int(rng.integers(2, 5))
        sz = int(rng.choice([3, 5, 7]))
        unit = 1 << (depth - 1)
        if case & 1:                                         # every level even -> fused kernels
            h, w = unit * int(rng.integers(2, 40)), unit * int(rng.integers(2, 60))
        else:                                                # arbitrary sizes -> mixed / general kernels
            h, w = int(rng.integers(unit * 2, 200)), int(rng.integers(unit * 2, 300))
        img = synth.image_np("float32", w, h, seed=600 + case)
        og, ol = oracle.pyramid(img, depth, M.GAUSS[sz])
        pg = hb.Pyramid(to_dev(hb, img, dev), depth)
        pl = hb.Pyramid(torch.zeros_like(pg.levels[0]), depth)
        hb.pyramid_traverse(pg, pl, M.GAUSS[sz])
        for lv in range(depth):
            np.testing.assert_array_equal(to_np(pg.levels[lv]), og[lv], err_msg=f"pyramid {h}x{w} depth {depth} mask {sz} gaus level {lv}")
            np.testing.assert_array_equal(to_np(pl.levels[lv]), ol[lv], err_msg=f"pyramid {h}x{w} depth {depth} mask {sz} lap level {lv}")
    for case in range(10):                                   # NN / LF accessors at arbitrary scale factors
        h, w = int(rng.integers(2, 90)), int(rng.integers(2, 130))
        oh, ow = int(rng.integers(1, 150)), int(rng.integers(1, 200))
        img = synth.image_np("float32", w, h, seed=700 + case)
        ip = [A.INTERP_NN, A.INTERP_LF][case & 1]
        got = hb.point_op(A.POINT_COPY, [to_dev(hb, img, dev)], A.F32, (oh, ow), [ip])
        np.testing.assert_array_equal(to_np(got), oracle.point_op(A.POINT_COPY, [img], A.F32, (oh, ow), [ip]), err_msg=f"interp {ip} {h}x{w}->{oh}x{ow}")
    for case in range(8):                                    # point arithmetic (streaming path incl. row tails) + reductions
        h, w = int(rng.integers(1, 70)), int(rng.choice([rng.integers(1, 50), rng.integers(300, 1100), 1024, 1027]))
        a, b2 = synth.image_np("float32", w, h, seed=800 + case), synth.image_np("float32", w, h, seed=900 + case)
        da, db = to_dev(hb, a, dev, bool(case & 1)), to_dev(hb, b2, dev, bool(case & 1))
        for op in (A.POINT_SUB, A.POINT_ADD, A.POINT_BLEND, A.POINT_MUL):
            np.testing.assert_array_equal(to_np(hb.point_op(op, [da, db], A.F32)), oracle.point_op(op, [a, b2], A.F32), err_msg=f"point {op} {h}x{w}")
        mn, mx, sm = hb.reduce_minmaxsum(da)
        assert np.float32(mn) == a.min() and np.float32(mx) == a.max() and abs(sm - a.astype(np.float64).sum()) <= 1e-5 * abs(a.astype(np.float64).sum()) + 1e-6
        u8 = synth.image_np("uint8", w, h, seed=950 + case)
        assert hb.reduce(to_dev(hb, u8, dev, bool(case & 1)), A.MAX) == u8.max() and hb.reduce(to_dev(hb, u8, dev), A.MIN) == u8.min()


# ------------------------------------------------------------------ CUDA IPC export guard
def test_ipc_export_refuses_pointers_inside_an_allocation(hb, dev):
    """a CUDA IPC handle names a whole allocation: exporting a pointer into the middle of one must fail loudly
    (the peer would map the allocation's base and the halo rows would land in the wrong place)"""
    import ctypes as C
    buf = hb.alloc_image(A.F32, 256, 64, device=dev)          # a whole allocation (hb_image_create)
    mem = A.hb_ipc_mem()
    assert hb.lib().hb_ipc_export(C.c_void_p(buf.data_ptr()), C.byref(mem)) == 0
    assert hb.lib().hb_ipc_export(C.c_void_p(buf.data_ptr() + 4096), C.byref(mem)) == A.HB_ERR_INVALID
    assert b"not the base of its allocation" in hb.lib().hb_last_error()


# ------------------------------------------------------------------ CUDA graphs (the reference's -use-graph mode)
def test_graph_replay_of_multi_kernel_pipelines(hb, oracle, dev):
    """hb_graph_begin / hb_graph_end capture the nine kernels of the unfused Harris pipeline and a pyramid traversal;
    replays on fresh inputs equal the oracle (no per-kernel host work between the kernels)."""
    import torch
    stream = torch.cuda.Stream(device=dev)
    img0, img1 = synth.blocks_np(640, 333, seed=5), synth.blocks_np(640, 333, seed=6)
    src = to_dev(hb, img0, dev)
    with torch.cuda.stream(stream):
        hb.harris_unfused(src, stream=stream)          # warm-up outside the capture (lazy kernel loading)
        stream.synchronize()
        n0 = hb.launch_count()
        with hb.Graph(stream) as g:
            out = hb.harris_unfused(src, stream=stream)[0]
        assert hb.launch_count() - n0 == 9
        for img in (img1, img0):
            src.copy_(torch.from_numpy(img).to(dev))
            stream.synchronize()
            g.launch()
            stream.synchronize()
            np.testing.assert_array_equal(to_np(out), oracle.harris(img))
        g.destroy()
        f0 = synth.image_np("float32", 256, 192, seed=31)
        pg = hb.Pyramid(to_dev(hb, f0, dev), 4)
        pl = hb.Pyramid(torch.zeros_like(pg.levels[0]), 4)
        hb.pyramid_traverse(pg, pl, M.GAUSS5, stream=stream)
        stream.synchronize()
        with hb.Graph(stream) as g2:
            hb.pyramid_traverse(pg, pl, M.GAUSS5, stream=stream)
        f1 = synth.image_np("float32", 256, 192, seed=32)
        pg.levels[0].copy_(torch.from_numpy(f1).to(dev))
        stream.synchronize()
        g2.launch()
        stream.synchronize()
        og, ol = oracle.pyramid(f1, 4, M.GAUSS5)
        for lv in range(4):
            np.testing.assert_array_equal(to_np(pg.levels[lv]), og[lv])
            np.testing.assert_array_equal(to_np(pl.levels[lv]), ol[lv])
        g2.destroy()


# ------------------------------------------------------------------ vector pixel types (uchar4)
@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT])
def test_rgba_local_ops_vs_golden(hb, dev, b):
    import os
    g = np.load(os.path.join(os.path.dirname(cases.GOLDEN_PATH), "reference_rgba.npz"))
    img = to_dev(hb, cases.rgba_image(*cases.RGBA_SHAPE), dev)
    for sz in (3, 5):
        np.testing.assert_array_equal(to_np(hb.local_op(S.gaussian_blur(M.GAUSS[sz], b), img)), g[f"gauss_rgba_{sz}_{b}"])
    np.testing.assert_array_equal(to_np(hb.local_op(S.laplace_u8(M.LAPLACE3, b, add=0), img)), g[f"laplace_rgba_3_{b}"])
    np.testing.assert_array_equal(to_np(hb.local_op(S.laplace_u8(M.LAPLACE5, b, add=0), img)), g[f"laplace_rgba_5_{b}"])
    np.testing.assert_array_equal(to_np(hb.local_op(S.minmax_u8(3, 3, True, b), img)), g[f"dilate_rgba_3_{b}"])
    np.testing.assert_array_equal(to_np(hb.local_op(S.box_blur_u8(5, 5, b), img)), g[f"box_rgba_5_{b}"])


@pytest.mark.parametrize("shape", [(300, 517), (33, 40), (2, 3), (1, 1), (131, 1024)])
@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT, A.UNDEFINED])
def test_rgba_local_ops_vs_oracle(hb, oracle, dev, shape, b):
    """several 128-element (32-pixel) tiles, partial tiles, images smaller than the halo; Gaussian 7x7 (float4
    accumulate), Laplace 5x5 (int4 accumulate over the Domain), dilate 3x3 (max over the Domain)"""
    if b == A.UNDEFINED and min(shape) < 8:
        pytest.skip("UNDEFINED reads outside tiny images are unspecified")
    h, w = shape
    img = cases.rgba_image(h, w, seed=17)
    d = to_dev(hb, img, dev)
    for spec in (S.gaussian_blur(M.GAUSS[7], b), S.laplace_u8(M.LAPLACE5, b), S.minmax_u8(3, 3, True, b)):
        got, want = to_np(hb.local_op(spec, d)), oracle.local_op_x4(spec, img)
        if b == A.UNDEFINED:   # only interior pixels are defined
            r = spec.size_y // 2
            got, want = got[r:h - r, r:w - r], want[r:h - r, r:w - r]
        np.testing.assert_array_equal(got, want)


def test_rgba_point_ops_and_padded_rows(hb, oracle, dev):
    import torch
    img = cases.rgba_image(77, 130, seed=19)
    a = to_dev(hb, img, dev)
    np.testing.assert_array_equal(to_np(hb.point_op(A.POINT_COPY, [a])), img)
    b2 = to_dev(hb, cases.rgba_image(77, 130, seed=20), dev)
    want = (img.astype(np.int32) + to_np(b2).astype(np.int32)).astype(np.uint8)    # uchar4 + uchar4 wraps per channel
    np.testing.assert_array_equal(to_np(hb.point_op(A.POINT_ADD, [a, b2])), want)
    # rows padded to 256 bytes (stride 192 px for 130): the view's stride counts pixels
    buf = torch.zeros((77, 192, 4), dtype=torch.uint8, device=dev)
    buf[:, :130] = a
    spec = S.gaussian_blur(M.GAUSS[5], A.MIRROR)
    np.testing.assert_array_equal(to_np(hb.local_op(spec, buf[:, :130])), oracle.local_op_x4(spec, img))


@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.CONSTANT])
def test_four_channel_intermediates_and_float4(hb, oracle, dev, b):
    """the other 4-channel pixel types (dsl/types.hpp): Sobel_RGBA's uchar4 -> int4 derivative and its int4, int4 -> uchar4
    combine (samples-public/3_Preprocessing/Sobel_RGBA/src/main.cpp:55-98), a uchar4 -> short4 derivative, float4 -> float4"""
    img = cases.rgba_image(61, 83, seed=29)
    d = to_dev(hb, img, dev)
    sx, sy = S.sobel_u8(M.SOBEL3_X, b), S.sobel_u8(M.SOBEL3_Y, b)
    gx, gy = hb.local_op(sx, d), hb.local_op(sy, d)
    assert str(gx.dtype) == "torch.int32" and tuple(gx.shape) == (61, 83, 4)
    wx, wy = oracle.local_op_x4(sx, img), oracle.local_op_x4(sy, img)
    np.testing.assert_array_equal(to_np(gx), wx)
    np.testing.assert_array_equal(to_np(gy), wy)
    mag = to_np(hb.point_op(A.POINT_SOBEL_COMBINE, [gx, gy], A.U8, p=(4, 0)))
    want = np.stack([oracle.point_op(A.POINT_SOBEL_COMBINE, [np.ascontiguousarray(wx[..., c]), np.ascontiguousarray(wy[..., c])], A.U8, p=(4, 0))
                     for c in range(4)], axis=-1)
    np.testing.assert_array_equal(mag, want)
    hd = S.harris_deriv(M.HARRIS_DX)
    np.testing.assert_array_equal(to_np(hb.local_op(hd, d)), oracle.local_op_x4(hd, img))   # short4 out
    f4 = np.ascontiguousarray(synth.image_np("float32", 70 * 4, 45, seed=30).reshape(45, 70, 4))
    for spec in (S.convolve_f32(M.GAUSS5, b), S.domain_reduce_f32(M.LAPLACE3.astype(np.float32), b)):
        np.testing.assert_array_equal(to_np(hb.local_op(spec, to_dev(hb, f4, dev))), oracle.local_op_x4(spec, f4))


def test_rgba_full_size_sample_shape(hb, oracle, dev):
    """Gaussian_Blur_RGBA's own size (4032 x 3024 uchar4): windows against the per-channel oracle"""
    import torch
    h, w = 3024, 4032
    img = torch.from_numpy(cases.rgba_image(h, w, seed=23)).to(dev)
    spec = S.gaussian_blur(M.GAUSS[5], A.CLAMP)
    out = hb.local_op(spec, img)
    for (y0, x0) in ((0, 0), (h - 64, w - 96), (1500, 2000), (0, w - 96), (h - 64, 0)):
        win = np.ascontiguousarray(to_np(img[max(y0 - 2, 0):y0 + 66, max(x0 - 2, 0):x0 + 98]))
        ref = oracle.local_op_x4(spec, win)
        oy, ox = y0 - max(y0 - 2, 0), x0 - max(x0 - 2, 0)
        # compare the part of the window whose 5x5 neighbourhood lies inside the window or at a true image border
        ys = slice(oy if y0 > 0 else 0, oy + 62 if y0 + 66 < h else None)
        xs = slice(ox if x0 > 0 else 0, ox + 94 if x0 + 98 < w else None)
        got = to_np(out[max(y0 - 2, 0):y0 + 66, max(x0 - 2, 0):x0 + 98])
        np.testing.assert_array_equal(got[ys, xs], ref[ys, xs])


# ------------------------------------------------------------------ binning / histogram
@pytest.mark.parametrize("shape", [(61, 83), (300, 1031), (2, 5), (1, 1), (257, 4096)])
@pytest.mark.parametrize("nb", [256, 64, 1000, 20000])
def test_histogram_f32_vs_oracle(hb, oracle, dev, shape, nb):
    img = synth.image_np("float32", shape[1], shape[0], seed=9, scale=254.99)
    got = hb.binning(to_dev(hb, img, dev), nb)
    np.testing.assert_array_equal(got, oracle.binning(img, nb))
    assert int(got.sum()) == img.size


def test_histogram_vs_golden(hb, dev):
    import os
    g = np.load(os.path.join(os.path.dirname(cases.GOLDEN_PATH), "reference_hist.npz"))
    img = synth.image_np("float32", cases.HIST_SHAPE[1], cases.HIST_SHAPE[0], seed=9, scale=254.99)
    for nb in cases.HIST_BINS:
        np.testing.assert_array_equal(hb.binning(to_dev(hb, img, dev), nb), g[f"hist_{nb}"])


def test_binning_kinds_roi_and_dropped_indices(hb, oracle, dev):
    u8 = synth.image_np("uint8", 1000, 333, seed=12)
    d = to_dev(hb, u8, dev)
    for nb, vk in ((256, A.BIN_VALUE_ONE), (100, A.BIN_VALUE_ONE), (256, A.BIN_VALUE_PIXEL)):
        np.testing.assert_array_equal(hb.binning(d, nb, A.BIN_INDEX_PIXEL, vk), oracle.binning(u8, nb, A.BIN_INDEX_PIXEL, vk))
    roi = (701, 200, 13, 7)
    np.testing.assert_array_equal(hb.binning(d, 256, A.BIN_INDEX_PIXEL, roi=roi), oracle.binning(u8, 256, A.BIN_INDEX_PIXEL, roi=roi))
    f = synth.image_np("float32", 515, 129, seed=13, scale=300.0) - 20.0   # pixels < 0 and >= 255 are dropped
    np.testing.assert_array_equal(hb.binning(to_dev(hb, f, dev), 256), oracle.binning(f, 256))
    np.testing.assert_array_equal(hb.binning(to_dev(hb, f, dev), 256, roi=(300, 100, 5, 3)), oracle.binning(f, 256, roi=(300, 100, 5, 3)))
    const = np.full((64, 512), 17.0, np.float32)   # every pixel in one bin: worst-case shared-memory contention
    got = hb.binning(to_dev(hb, const, dev), 256)
    assert got[17] == const.size and got.sum() == const.size


def test_binning_pixels_of_no_bin_take_the_spare_word(hb, dev):
    """The branch-free shared-memory path steers pixels of no bin to a spare word instead of branching around the atomic:
    NaN, +-inf, values <= -1 (after scaling), >= num_bins and huge magnitudes are dropped, (-1, 0) counts in bin 0 like
    the C conversion `(uint)v` -- the contract stated in hipacc_b200.h; the rest of the image is counted exactly."""
    rng = np.random.default_rng(5)
    base = (rng.random((97, 260)) * 254.0).astype(np.float32)
    want = np.bincount((base / np.float32(255.0) * np.float32(256.0)).astype(np.int64).ravel(), minlength=256).astype(np.uint32)
    img = np.concatenate([base, np.zeros((1, 260), np.float32)], axis=0)
    special = [np.nan, np.inf, -np.inf, -1.0, -300.0, 255.0, 1e9, 3e38, -3e38, 1e30, -0.25, -0.9, 254.9999]
    img[-1, :] = 10.0
    img[-1, :len(special)] = special
    want[int(np.float32(10.0) / np.float32(255.0) * np.float32(256.0))] += 260 - len(special)
    want[0] += 2            # -0.25 and -0.9 scale into (-1, 0): (uint) truncation gives bin 0
    want[255] += 1          # 254.9999
    for padded in (True, False):
        np.testing.assert_array_equal(hb.binning(to_dev(hb, img, dev, padded), 256), want)
    # unscaled index kind on float pixels
    f = np.array([[0.0, 0.99, 1.0, 63.5, 64.0, -0.5, -1.0, np.nan, np.inf, 1e20] * 8], np.float32)
    w2 = np.zeros(64, np.uint32)
    w2[0] = 3 * 8; w2[1] = 8; w2[63] = 8
    np.testing.assert_array_equal(hb.binning(to_dev(hb, f, dev), 64, A.BIN_INDEX_PIXEL), w2)


def test_histogram_full_size_properties(hb, dev):
    """8192^2 float (C3's image): counts sum to the pixel count and match torch.histc-free integer binning on the device"""
    import torch
    img = synth.image_torch("float32", 8192, 8192, seed=3, scale=254.99, device=dev)
    got = hb.binning(img, 256)
    assert int(got.sum()) == 8192 * 8192
    # tensor / tensor is an IEEE division (tensor / python scalar multiplies by the reciprocal on CUDA)
    idx = (img / torch.full_like(img, 255.0) * 256.0).to(torch.int64).flatten()
    want = torch.bincount(idx, minlength=256).cpu().numpy().astype(np.uint32)
    np.testing.assert_array_equal(got, want)


# ------------------------------------------------------------------ pyramid
@pytest.mark.parametrize("idx", range(len(cases.PYR_CASES)))
@pytest.mark.parametrize("with_tmp", [False, True], ids=["fused_down", "unfused_down"])
def test_pyramid_vs_golden_and_oracle(hb, oracle, dev, idx, with_tmp):
    h, w, depth, sz = cases.PYR_CASES[idx]
    img = synth.image_np("float32", w, h, seed=7 + idx)
    import torch
    pg = hb.Pyramid(to_dev(hb, img, dev), depth)
    pl = hb.Pyramid(torch.zeros_like(pg.levels[0]), depth)
    pt = hb.Pyramid(torch.zeros_like(pg.levels[0]), depth) if with_tmp else None
    hb.pyramid_traverse(pg, pl, M.GAUSS[sz], ptmp=pt)
    g = cases.golden()
    for lv in range(depth):
        np.testing.assert_array_equal(to_np(pg.levels[lv]), g[f"pyr{idx}_g{lv}"])
        np.testing.assert_array_equal(to_np(pl.levels[lv]), g[f"pyr{idx}_l{lv}"])


def test_pyramid_odd_sizes_vs_oracle(hb, oracle, dev):
    import torch
    img = synth.image_np("float32", 203, 131, seed=51)
    og, ol = oracle.pyramid(img, 4, M.GAUSS5)
    pg = hb.Pyramid(to_dev(hb, img, dev), 4)
    pl = hb.Pyramid(torch.zeros_like(pg.levels[0]), 4)
    hb.pyramid_traverse(pg, pl, M.GAUSS5)
    for lv in range(4):
        np.testing.assert_array_equal(to_np(pg.levels[lv]), og[lv])
        np.testing.assert_array_equal(to_np(pl.levels[lv]), ol[lv])


# ------------------------------------------------------------------ runtime layer
def test_image_memory_roundtrip(hb, dev):
    import ctypes as C
    L = hb.lib()
    v = A.hb_view()
    assert L.hb_image_create(A.F32, 101, 37, 0, C.byref(v)) == 0
    assert v.stride % 64 == 0 and v.stride >= 101           # 256-byte row alignment
    src = synth.image_np("float32", 101, 37, seed=61)
    assert L.hb_image_write(C.byref(v), src.ctypes.data_as(C.c_void_p), None) == 0
    v2 = A.hb_view()
    assert L.hb_image_create(A.F32, 101, 37, 4, C.byref(v2)) == 0 and v2.stride == 101
    assert L.hb_image_copy(C.byref(v), C.byref(v2), None) == 0
    back = np.zeros_like(src)
    assert L.hb_image_read(C.byref(v2), back.ctypes.data_as(C.c_void_p), None) == 0
    np.testing.assert_array_equal(back, src)
    # region copy
    r_src = A.make_view(v.data, A.F32, 101, 37, v.stride, (20, 10, 5, 3))
    r_dst = A.make_view(v2.data, A.F32, 101, 37, v2.stride, (20, 10, 50, 20))
    assert L.hb_image_copy_region(C.byref(r_src), C.byref(r_dst), None) == 0
    assert L.hb_image_read(C.byref(v2), back.ctypes.data_as(C.c_void_p), None) == 0
    want = src.copy()
    want[20:30, 50:70] = src[3:13, 5:25]
    np.testing.assert_array_equal(back, want)
    assert L.hb_image_destroy(C.byref(v)) == 0 and L.hb_image_destroy(C.byref(v2)) == 0


def test_timing_and_launch_counter(hb, dev):
    d = to_dev(hb, synth.image_np("float32", 256, 256, seed=62), dev)
    n0 = hb.launch_count()
    hb.set_timing(True)
    hb.local_op(S.convolve_f32(M.GAUSS3), d)
    ms = hb.last_kernel_ms()
    hb.set_timing(False)
    assert hb.launch_count() == n0 + 1 and 0.0 < ms < 50.0


# ------------------------------------------------------------------ BASELINE.json full sizes
def test_full_size_c1_gaussian_u8_4096(hb, oracle, dev):
    img = synth.image_np("uint8", 4096, 4096, seed=1)
    spec = S.gaussian_blur(M.GAUSS5, A.CLAMP)
    got = to_np(hb.local_op(spec, to_dev(hb, img, dev)))
    np.testing.assert_array_equal(got, oracle.local_op(spec, img))


@pytest.mark.parametrize("name", ["sobel_x", "sobel_y", "laplace"])
def test_full_size_c2_float_8192_mirror(hb, oracle, dev, name):
    m = {"sobel_x": M.SOBEL3_X, "sobel_y": M.SOBEL3_Y, "laplace": M.LAPLACE3}[name].astype(np.float32)
    img = synth.image_np("float32", 8192, 8192, seed=2)
    spec = S.domain_reduce_f32(m, A.MIRROR)
    got = to_np(hb.local_op(spec, to_dev(hb, img, dev)))
    np.testing.assert_array_equal(got, oracle.local_op(spec, img))
    # linearity property (size independent): op(2*img) == 2*op(img) exactly for power-of-two scaling
    got2 = to_np(hb.local_op(spec, to_dev(hb, img * np.float32(2.0), dev)))
    np.testing.assert_array_equal(got2, got * np.float32(2.0))


def test_full_size_c3_reduce_8192(hb, dev):
    img = synth.image_np("float32", 8192, 8192, seed=3, scale=255.0)
    mn, mx, sm = hb.reduce_minmaxsum(to_dev(hb, img, dev))
    s64 = img.astype(np.float64).sum()
    assert np.float32(mn) == img.min() and np.float32(mx) == img.max()
    assert abs(sm - s64) <= 1e-5 * s64
    # the reference's own orders for context: serial float fold saturates


def test_full_size_c3_bilateral_windows(hb, oracle, dev):
    """8192^2 x 169 expf is minutes on the CPU: check border bands and interior windows of the full-size result."""
    img = synth.image_np("float32", 8192, 8192, seed=3, scale=255.0)
    cm = M.bilateral_mask(13)
    got = to_np(hb.bilateral(to_dev(hb, img, dev), 13, cm, 16, A.MIRROR))
    for (y0, x0) in [(0, 0), (0, 8192 - 160), (8192 - 96, 0), (8192 - 96, 8192 - 160), (4000, 4000), (1234, 7000)]:
        ys, xs = slice(max(0, y0 - 6), min(8192, y0 + 96 + 6)), slice(max(0, x0 - 6), min(8192, x0 + 160 + 6))
        crop = np.ascontiguousarray(img[ys, xs])
        want = oracle.bilateral(crop, 13, cm, 16, A.MIRROR)
        oy, ox = y0 - ys.start, x0 - xs.start
        # the crop's own borders are only valid where they coincide with the image borders
        np.testing.assert_allclose(got[y0:y0 + 96, x0:x0 + 160], want[oy:oy + 96, ox:ox + 160], rtol=1e-5, atol=0)


@pytest.mark.parametrize("h,w,depth,sz", [(200, 392, 3, 5), (264, 520, 4, 3), (96, 644, 3, 7), (130, 258, 2, 5)])
def test_pyramid_exact_halving_multi_tile_vs_oracle(hb, oracle, dev, h, w, depth, sz):
    """Even level sizes take the fused kernels (blur+subsample+DoG in one, Restore+Blend in one); several
    128 x 32 tiles per level incl. partial tiles at the right / bottom edge."""
    import torch
    img = synth.image_np("float32", w, h, seed=60 + sz)
    og, ol = oracle.pyramid(img, depth, M.GAUSS[sz])
    pg = hb.Pyramid(to_dev(hb, img, dev), depth)
    pl = hb.Pyramid(torch.zeros_like(pg.levels[0]), depth)
    hb.pyramid_traverse(pg, pl, M.GAUSS[sz])
    for lv in range(depth):
        np.testing.assert_array_equal(to_np(pg.levels[lv]), og[lv])
        np.testing.assert_array_equal(to_np(pl.levels[lv]), ol[lv])


@pytest.mark.parametrize("world,h,w,depth,sz", [(2, 256, 392, 3, 5), (4, 512, 264, 4, 3), (3, 384, 520, 3, 7), (8, 1024, 256, 4, 5)])
def test_pyramid_row_strips_equal_unsharded(hb, dev, world, h, w, depth, sz):
    """SURVEY 8e pyramid sharding: every rank's strips (ghost rows filled by the halo exchange, emulated here by
    device copies between the strip buffers of all ranks in one process) give bit-identical levels."""
    import torch
    from hipacc_b200 import strips
    img = to_dev(hb, synth.image_np("float32", w, h, seed=70 + world), dev)
    pg = hb.Pyramid(img.clone(), depth)
    pl = hb.Pyramid(torch.zeros_like(img), depth)
    hb.pyramid_traverse(pg, pl, M.GAUSS[sz])

    R = sz // 2 + 2
    pgs = [strips.StripPyramid(w, h, depth, world, r, R, dev) for r in range(world)]
    pls = [strips.StripPyramid(w, h, depth, world, r, R, dev) for r in range(world)]
    for r in range(world):
        p0 = pgs[r].plans[0]
        pgs[r].owned(0).copy_(img[p0.y0:p0.y1])
        for l in range(depth):   # poison the ghost rows: stale data must never be read
            pgs[r].bufs[l][:pgs[r].plans[l].ghost_top] = float("nan")
            pls[r].bufs[l][:pls[r].plans[l].ghost_top] = float("nan")

    def exchange_all(pyrs, l):
        for r in range(world):
            pl_, buf = pyrs[r].plans[l], pyrs[r].bufs[l]
            if pl_.ghost_top:
                q, src = pyrs[r - 1].plans[l], pyrs[r - 1].bufs[l]
                buf[0:pl_.ghost_top] = src[q.ghost_top + q.rows - R:q.ghost_top + q.rows]
            if pl_.ghost_bottom:
                q, src = pyrs[r + 1].plans[l], pyrs[r + 1].bufs[l]
                buf[pl_.ghost_top + pl_.rows:pl_.ghost_top + pl_.rows + R] = src[q.ghost_top:q.ghost_top + R]

    for l in range(1, depth):
        exchange_all(pgs, l - 1)
        for r in range(world):
            strips.pyramid_down_step(hb, pgs[r], pls[r], l, M.GAUSS[sz])
    for l in range(depth - 2, -1, -1):
        exchange_all(pgs, l + 1)
        exchange_all(pls, l + 1)
        for r in range(world):
            strips.pyramid_up_step(hb, pgs[r], pls[r], l)
    for l in range(depth):
        got_g = torch.cat([pgs[r].owned(l) for r in range(world)])
        got_l = torch.cat([pls[r].owned(l) for r in range(world)])
        np.testing.assert_array_equal(to_np(got_g), to_np(pg.levels[l]))
        np.testing.assert_array_equal(to_np(got_l), to_np(pl.levels[l]))


def test_full_size_c4_harris_strip_of_32k(hb, oracle, dev):
    """One 32768-wide strip (the per-GPU share at 8 GPUs is 32768 x 4096): fused kernel vs oracle pipeline."""
    img = synth.image_np("uint8", 32768, 512, seed=4)
    np.testing.assert_array_equal(to_np(hb.harris(to_dev(hb, img, dev))), oracle.harris(img))


def test_full_size_c4_harris_32768_windows(hb, oracle, dev):
    """The release configuration itself: the fused Harris kernel on the whole 32768 x 32768 image (1 GiB in, 1 GiB out),
    checked bit for bit against the oracle's 9-kernel pipeline on windows at the four corners, at the row-strip edges of
    the 2 / 4 / 8 GPU partitions and in the interior.  A window's own borders are valid only where they are image borders,
    so every window is computed with an 8-pixel margin that is cut off unless it is the image edge."""
    n, H, W, m = 32768, 96, 640, 8
    img = hb.empty_image(A.U8, n, n, device=dev)
    for y in range(0, n, 4096):   # generated on the device, strip by strip
        img[y:y + 4096].copy_(synth.image_torch("uint8", n, 4096, seed=4, y0=y, device=dev))
    out = hb.harris(img)
    wins = [(0, 0), (0, n - W), (n - H, 0), (n - H, n - W), (16000, 16000), (7777, 30001)]
    wins += [(k * 4096 - H // 2, x) for k in range(1, 8) for x in (0, 12345, n - W)]   # strip edges of every partition
    for (y0, x0) in wins:
        ys, xs = slice(max(0, y0 - m), min(n, y0 + H + m)), slice(max(0, x0 - m), min(n, x0 + W + m))
        crop = synth.image_np("uint8", xs.stop - xs.start, ys.stop - ys.start, seed=4, x0=xs.start, y0=ys.start)
        want = oracle.harris(crop)
        oy, ox = y0 - ys.start, x0 - xs.start
        np.testing.assert_array_equal(to_np(out[y0:y0 + H, x0:x0 + W]), want[oy:oy + H, ox:ox + W], err_msg=f"window at ({y0}, {x0})")
    assert 0 < int(out[:4096].sum().item())   # the detector fires


def test_full_size_c5_pyramid_down_pass_windows(hb, oracle, dev):
    """16384 x 16384, the fused down step (blur + subsample + DoG) at full size for three transitions: gaus(1..3) and
    lap(0..2) are local functions of the input, so windows of them are compared bit for bit with the oracle's unfused
    operators run on crops of the level-0 image (corners, partition edges, interior).  Crops start on multiples of 64 so
    that every level's sampling grid coincides with the global one."""
    n, S0, m = 16384, 512, 64
    g = [hb.empty_image(A.F32, n, n, device=dev)]
    for y in range(0, n, 2048):
        g[0][y:y + 2048].copy_(synth.image_torch("float32", n, 2048, seed=5, y0=y, device=dev))
    lap = []
    for l in range(1, 4):
        g.append(hb.empty_image(A.F32, n >> l, n >> l, device=dev))
        lap.append(hb.empty_image(A.F32, n >> (l - 1), n >> (l - 1), device=dev))
        hb.pyr_down(g[l - 1], g[l], M.GAUSS5, lap_fine=lap[l - 1])
    blur = S.convolve_f32(M.GAUSS5, A.CLAMP)
    for (y0, x0) in [(0, 0), (0, n - S0), (n - S0, 0), (n - S0, n - S0), (8192 - 256, 4096), (2048 - 256, n - S0), (6144, 8192 - 256)]:
        ys, xs = slice(max(0, y0 - m), min(n, y0 + S0 + m)), slice(max(0, x0 - m), min(n, x0 + S0 + m))
        og = [synth.image_np("float32", xs.stop - xs.start, ys.stop - ys.start, seed=5, x0=xs.start, y0=ys.start)]
        ol = []
        for l in range(1, 4):   # the sample's way down (Gaussian_Laplacian_Pyramid/src/main.cpp:199-225), unfused
            tmp = oracle.local_op(blur, og[l - 1])
            og.append(oracle.point_op(A.POINT_COPY, [tmp], A.F32, (og[l - 1].shape[0] // 2, og[l - 1].shape[1] // 2), [A.INTERP_NN]))
            ol.append(oracle.point_op(A.POINT_SUB, [og[l - 1], og[l]], A.F32, og[l - 1].shape, [A.INTERP_NO, A.INTERP_LF]))
        for l in range(0, 4):
            # the window at level l, shrunk by 8 pixels where the crop side is not an image side (the crop's own CLAMP)
            a0, a1 = (y0 >> l) + (8 if ys.start > 0 else 0), ((y0 + S0) >> l) - (8 if ys.stop < n else 0)
            b0, b1 = (x0 >> l) + (8 if xs.start > 0 else 0), ((x0 + S0) >> l) - (8 if xs.stop < n else 0)
            cy, cx = ys.start >> l, xs.start >> l
            if l >= 1:
                np.testing.assert_array_equal(to_np(g[l][a0:a1, b0:b1]), og[l][a0 - cy:a1 - cy, b0 - cx:b1 - cx], err_msg=f"gaus({l}) window ({y0}, {x0})")
            if l <= 2:
                np.testing.assert_array_equal(to_np(lap[l][a0:a1, b0:b1]), ol[l][a0 - cy:a1 - cy, b0 - cx:b1 - cx], err_msg=f"lap({l}) window ({y0}, {x0})")


def test_reduce_min_max_with_nan_pixels(hb, dev):
    """Contract for non-finite input (include/hipacc_b200.h, hb_reduce): MIN / MAX are the IEEE minimum / maximum of the
    pixels that are numbers -- a NaN pixel never wins -- and +-inf take part normally.  (The DSL's `l < r ? l : r` fold
    returns a value that depends on WHERE the NaN sits in the iteration order, which no parallel reduction reproduces;
    the reference's own CUDA reduction does not either, runtime/hipacc_cu_red.hpp:140-346.)"""
    f = synth.image_np("float32", 700, 300, seed=50) - np.float32(0.5)
    f[17, 333] = np.nan
    f[0, 0] = np.nan
    f[299, 699] = np.nan
    f[100, 100] = np.inf
    mn, mx, _ = hb.reduce_minmaxsum(to_dev(hb, f, dev))
    assert np.float32(mn) == np.nanmin(f) and np.float32(mx) == np.inf
    assert np.float32(hb.reduce(to_dev(hb, f, dev), A.MIN)) == np.nanmin(f)
    f[100, 100] = -np.inf
    assert np.float32(hb.reduce(to_dev(hb, f, dev), A.MIN)) == -np.inf and np.float32(hb.reduce(to_dev(hb, f, dev), A.MAX)) == np.nanmax(f)


def test_full_size_c5_pyramid_properties(hb, dev):
    """16384^2, 8 levels: restored level 0 equals the input (the sample's own check) and the coarsest level is 128^2."""
    import torch
    img = synth.image_torch("float32", 16384, 16384, seed=5, device=dev)
    base = hb.empty_image(A.F32, 16384, 16384, device=dev)
    base.copy_(img)
    pg = hb.Pyramid(base, 8)
    pl = hb.Pyramid(hb.empty_image(A.F32, 16384, 16384, device=dev).zero_(), 8)
    hb.pyramid_traverse(pg, pl, M.GAUSS5)
    assert tuple(pg.levels[7].shape) == (128, 128)
    err = (pg.levels[0] - img).abs().max().item()
    assert err <= 1e-5 * 1.0 + 1e-6   # (g - LF(c)) + LF(c) == g up to one float rounding
    assert torch.isfinite(pl.levels[0]).all().item()


@pytest.mark.parametrize("split", [False, True], ids=["one_launch_level0", "three_band_level0"])
@pytest.mark.parametrize("world,h,w,depth,sz,G", [(2, 512, 392, 4, 5, 2), (4, 1024, 264, 4, 3, 2), (8, 4096, 256, 5, 5, 3), (3, 768, 520, 3, 7, 1), (2, 512, 128, 3, 5, 2)])
def test_sharded_pyramid_one_exchange_one_gather_equals_unsharded(hb, dev, world, h, w, depth, sz, G, split):
    """strips.ShardedPyramid (recompute-in-halo: ONE level-0 halo exchange + ONE all-gather of level G, no exchange
    on the way up) gives bit-identical levels; all ranks emulated in one process, communication by device copies,
    everything a rank does not own or compute is poisoned with NaN first."""
    import torch
    from hipacc_b200 import strips
    img = to_dev(hb, synth.image_np("float32", w, h, seed=90 + world), dev)
    pg = hb.Pyramid(img.clone(), depth)
    pl = hb.Pyramid(torch.zeros_like(img), depth)
    hb.pyramid_traverse(pg, pl, M.GAUSS[sz])

    plans = [strips.PyramidShardPlan(w, h, depth, world, r, sz, gather_level=G) for r in range(world)]
    sp = [strips.ShardedPyramid(p, dev) for p in plans]
    for r, (p, s) in enumerate(zip(plans, sp)):
        for l in range(depth):
            s.gaus[l].fill_(float("nan"))
            s.lap[l].fill_(float("nan"))
        s.lap[depth - 1].zero_()          # the coarsest Laplacian level is never written (zeros, like the unsharded run)
        s.owned(s.gaus, 0).copy_(img[p.y0(0):p.y1(0)])
    # the one halo exchange: E0 rows of level 0 from each neighbour
    for r, (p, s) in enumerate(zip(plans, sp)):
        a, b = p.buffer_span(0)
        s.gaus[0][:, :w].copy_(img[a:b])
        assert p.E0 <= p.rows(0)
    for s in sp:
        if split:   # the form that overlaps the exchange: interior band (own rows only) first, edge bands after the halo arrived
            assert s._can_split0()
            s._down0(hb, M.GAUSS[sz], None, "interior")
            s._down0(hb, M.GAUSS[sz], None, "edges")
            s.down_sharded(hb, M.GAUSS[sz], first=2)
        else:
            s.down_sharded(hb, M.GAUSS[sz])
    # the one all-gather: every rank publishes ITS rows of gaus(G)
    Gl = plans[0].G
    for r, (p, s) in enumerate(zip(plans, sp)):
        for q, t in zip(plans, sp):
            if t is not s:
                t.gaus[Gl][p.y0(Gl):p.y1(Gl)].copy_(s.gaus[Gl][p.y0(Gl):p.y1(Gl)])
    for s in sp:
        s.coarse_and_up(hb, M.GAUSS[sz])
    for l in range(depth):
        got_g = torch.cat([s.owned(s.gaus, l) for s in sp])
        got_l = torch.cat([s.owned(s.lap, l) for s in sp])
        np.testing.assert_array_equal(to_np(got_g), to_np(pg.levels[l]))
        np.testing.assert_array_equal(to_np(got_l), to_np(pl.levels[l]))
        if l >= Gl:   # replicated levels: every rank holds the whole image
            for s in sp:
                np.testing.assert_array_equal(to_np(s.gaus[l][:, :w >> l]), to_np(pg.levels[l]))


@pytest.mark.parametrize("h,w,depth,sz", [(1024, 1024, 4, 5), (512, 768, 5, 3), (256, 640, 3, 7), (64, 128, 2, 5), (1024, 512, 8, 5)])
def test_coarse_pyramid_in_one_launch_equals_per_level(hb, oracle, dev, h, w, depth, sz):
    """hb_pyr_traverse_coarse (one cooperative kernel, grid-wide barriers between the transitions) is bit-identical
    to the per-level kernels, also when replayed from a CUDA graph and on a second input (barrier words reset)."""
    import torch
    stream = torch.cuda.Stream(device=dev)
    imgs = [synth.image_np("float32", w, h, seed=95 + k) for k in range(2)]
    with torch.cuda.stream(stream):
        pg = hb.Pyramid(to_dev(hb, imgs[0], dev), depth)
        pl = hb.Pyramid(torch.zeros_like(pg.levels[0]), depth)
        qg = hb.Pyramid(to_dev(hb, imgs[0], dev), depth)
        ql = hb.Pyramid(torch.zeros_like(qg.levels[0]), depth)
        for k in range(2):
            for pyr, img in ((pg, imgs[k]), (qg, imgs[k])):
                pyr.levels[0].copy_(torch.from_numpy(img).to(dev))
            for l in range(depth):
                pl.levels[l].zero_(); ql.levels[l].zero_()
            hb.pyramid_traverse(pg, pl, M.GAUSS[sz], stream=stream, fuse_coarse=False)
            assert hb.pyr_traverse_coarse(qg.levels, ql.levels, M.GAUSS[sz], stream=stream)
            stream.synchronize()
            for l in range(depth):
                np.testing.assert_array_equal(to_np(qg.levels[l]), to_np(pg.levels[l]))
                np.testing.assert_array_equal(to_np(ql.levels[l]), to_np(pl.levels[l]))
        # graph replay
        qg.levels[0].copy_(torch.from_numpy(imgs[0]).to(dev))
        for l in range(depth):
            ql.levels[l].zero_()
        stream.synchronize()
        with hb.Graph(stream) as g:
            hb.pyr_traverse_coarse(qg.levels, ql.levels, M.GAUSS[sz], stream=stream)
        g.launch()
        stream.synchronize()
        pg.levels[0].copy_(torch.from_numpy(imgs[0]).to(dev))
        for l in range(depth):
            pl.levels[l].zero_()
        hb.pyramid_traverse(pg, pl, M.GAUSS[sz], stream=stream, fuse_coarse=False)
        stream.synchronize()
        for l in range(depth):
            np.testing.assert_array_equal(to_np(qg.levels[l]), to_np(pg.levels[l]))
        g.destroy()
    if h * w <= 1 << 18:   # and against the oracle
        og, ol = oracle.pyramid(imgs[0], depth, M.GAUSS[sz])
        for l in range(depth):
            np.testing.assert_array_equal(to_np(pg.levels[l]), og[l])


def test_full_pyramid_with_fused_coarse_end_vs_oracle(hb, oracle, dev):
    """pyramid_traverse(fuse_coarse=True) routes the levels of <= 2^20 pixels through the one-launch coarse traversal"""
    import torch
    img = synth.image_np("float32", 2048, 1024, seed=97)
    og, ol = oracle.pyramid(img, 6, M.GAUSS5)
    pg = hb.Pyramid(to_dev(hb, img, dev), 6)
    pl = hb.Pyramid(torch.zeros_like(pg.levels[0]), 6)
    assert hb.coarse_start(pg.levels) == 1
    n0 = hb.launch_count()
    hb.pyramid_traverse(pg, pl, M.GAUSS5, fuse_coarse=True)
    assert hb.launch_count() - n0 == 3          # down 0->1, the coarse end, up 1->0
    for lv in range(6):
        np.testing.assert_array_equal(to_np(pg.levels[lv]), og[lv])
        np.testing.assert_array_equal(to_np(pl.levels[lv]), ol[lv])
