"""CPU tests (-m "not gpu"): pin the restated oracle (oracle/emit_cpu.cpp) against
  (1) the committed golden outputs of the compiled reference DSL (tests/golden/),
  (2) SURVEY.md appendix A known-answer vectors,
  (3) the live compiled reference (oracle/_ref) incl. ROI / crop / ragged cases, when present,
  (4) the reference samples' embedded plain-C checkers (interior pixels).
"""
import ctypes as C

import numpy as np
import pytest

import cases
from hipacc_b200 import _abi as A, masks as M, specs as S, synth


# ------------------------------------------------------------------ (1) golden fixtures
@pytest.mark.parametrize("case", cases.local_cases(), ids=cases.case_id)
def test_local_vs_golden(oracle, case):
    key, inp, spec = case
    got = oracle.local_op(spec, cases.inputs()[inp])
    np.testing.assert_array_equal(got, cases.golden()[key])  # bit-exact, float included


def test_sobel_combine_vs_golden(oracle):
    u8 = cases.inputs()["u8"]
    a = oracle.local_op(S.sobel_u8(M.SOBEL3_X), u8)
    b = oracle.local_op(S.sobel_u8(M.SOBEL3_Y), u8)
    got = oracle.point_op(A.POINT_SOBEL_COMBINE, [a, b], A.U8, p=(4, 0))
    np.testing.assert_array_equal(got, cases.golden()["sobel_combine"])


@pytest.mark.parametrize("size", [3, 5, 13])
def test_bilateral_vs_golden(oracle, size):
    inp = cases.inputs()
    g = cases.golden()
    np.testing.assert_array_equal(oracle.bilateral(inp["u8"], size, M.bilateral_mask(size), 16, A.CLAMP),
                                  g[f"bilateral_u8_{size}"])
    # same libm expf on both sides -> bit-exact on the CPU
    np.testing.assert_array_equal(oracle.bilateral(inp["f255"], size, M.bilateral_mask(size), 16, A.MIRROR),
                                  g[f"bilateral_f32_{size}"])


def test_harris_vs_golden(oracle):
    g = cases.golden()
    out, gx, gy, gxy = oracle.harris(cases.inputs()["harris"], return_intermediates=True)
    np.testing.assert_array_equal(gx, g["harris_gx"])
    np.testing.assert_array_equal(gy, g["harris_gy"])
    np.testing.assert_array_equal(gxy, g["harris_gxy"])
    np.testing.assert_array_equal(out, g["harris_out"])
    assert 0 < out.sum() < out.size // 4  # sparse, non-trivial corner map


def test_interp_vs_golden(oracle):
    g = cases.golden()
    f32 = cases.inputs()["f32"]
    h, w = f32.shape
    nn = oracle.point_op(A.POINT_COPY, [f32], A.F32, (h // 2, w // 2), [A.INTERP_NN])
    np.testing.assert_array_equal(nn, g["interp_nn"])
    lf = oracle.point_op(A.POINT_COPY, [nn], A.F32, (h, w), [A.INTERP_LF])
    np.testing.assert_array_equal(lf, g["interp_lf"])


@pytest.mark.parametrize("idx", range(len(cases.PYR_CASES)))
def test_pyramid_vs_golden(oracle, idx):
    h, w, depth, sz = cases.PYR_CASES[idx]
    img = synth.image_np("float32", w, h, seed=7 + idx)
    gaus, lap = oracle.pyramid(img, depth, M.GAUSS[sz])
    g = cases.golden()
    for lv in range(depth):
        np.testing.assert_array_equal(gaus[lv], g[f"pyr{idx}_g{lv}"])
        np.testing.assert_array_equal(lap[lv], g[f"pyr{idx}_l{lv}"])
    # the sample's own self-check: restored == input (Gaussian_Laplacian_Pyramid/src/main.cpp:259-260)
    np.testing.assert_allclose(gaus[0], img, rtol=1e-3, atol=1e-6)


def test_global_reduce_vs_golden(oracle):
    f32 = cases.inputs()["f32"]
    g = cases.golden()
    for mode, nm in ((A.SUM, "sum"), (A.MIN, "min"), (A.MAX, "max")):
        assert oracle.reduce_serial(f32, mode) == g[f"reduce_{nm}"][0]
    mn, mx, sm, s64 = oracle.reduce_minmaxsum(f32)
    assert mn == g["reduce_min"][0] and mx == g["reduce_max"][0]
    assert abs(float(sm) - s64) <= 1e-5 * abs(s64)
    assert abs(s64 - f32.astype(np.float64).sum()) <= 1e-9 * abs(s64)


@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT])
def test_rgba_vs_golden(oracle, b):
    """uchar4 local operators (Gaussian_Blur_RGBA, Laplace_RGBA executed by the reference DSL) == the scalar operator
    on each channel plane: the per-channel claim the CUDA uchar4 path rests on."""
    import os
    g = np.load(os.path.join(os.path.dirname(cases.GOLDEN_PATH), "reference_rgba.npz"))
    img = cases.rgba_image(*cases.RGBA_SHAPE)
    for sz in (3, 5):
        np.testing.assert_array_equal(oracle.local_op_x4(S.gaussian_blur(M.GAUSS[sz], b), img), g[f"gauss_rgba_{sz}_{b}"])
    np.testing.assert_array_equal(oracle.local_op_x4(S.laplace_u8(M.LAPLACE3, b, add=0), img), g[f"laplace_rgba_3_{b}"])
    np.testing.assert_array_equal(oracle.local_op_x4(S.laplace_u8(M.LAPLACE5, b, add=0), img), g[f"laplace_rgba_5_{b}"])
    np.testing.assert_array_equal(oracle.local_op_x4(S.minmax_u8(3, 3, True, b), img), g[f"dilate_rgba_3_{b}"])
    np.testing.assert_array_equal(oracle.local_op_x4(S.box_blur_u8(5, 5, b), img), g[f"box_rgba_5_{b}"])


def test_histogram_vs_golden(oracle):
    """Histogram sample (binning() + binned_data()) executed by the reference DSL -> tests/golden/reference_hist.npz"""
    import os
    g = np.load(os.path.join(os.path.dirname(cases.GOLDEN_PATH), "reference_hist.npz"))
    img = synth.image_np("float32", cases.HIST_SHAPE[1], cases.HIST_SHAPE[0], seed=9, scale=254.99)
    for nb in cases.HIST_BINS:
        got = oracle.binning(img, nb)
        np.testing.assert_array_equal(got, g[f"hist_{nb}"])
        assert int(got.sum()) == img.size


def test_binning_kinds_and_dropped_indices(oracle):
    """indices >= num_bins are dropped (BINNING Put helper, runtime/hipacc_cpu_red.hpp:71-76); uchar pixel-indexed
    histogram and value = pixel against numpy"""
    u8 = synth.image_np("uint8", 97, 53, seed=12)
    np.testing.assert_array_equal(oracle.binning(u8, 256, A.BIN_INDEX_PIXEL), np.bincount(u8.ravel(), minlength=256).astype(np.uint32))
    np.testing.assert_array_equal(oracle.binning(u8, 100, A.BIN_INDEX_PIXEL), np.bincount(u8.ravel(), minlength=256)[:100].astype(np.uint32))
    sums = np.bincount(u8.ravel(), weights=u8.ravel().astype(np.float64), minlength=256).astype(np.uint32)
    np.testing.assert_array_equal(oracle.binning(u8, 256, A.BIN_INDEX_PIXEL, A.BIN_VALUE_PIXEL), sums)
    f = synth.image_np("float32", 64, 40, seed=13, scale=300.0) - 20.0   # some pixels < 0 and >= 255: dropped
    idx = (f / np.float32(255.0) * np.float32(256)).astype(np.int64)
    want = np.bincount(idx[(idx >= 0) & (idx < 256) & (f > -1.0)], minlength=256).astype(np.uint32)
    # (uint) of a value in (-1, 0) truncates to 0 in C
    np.testing.assert_array_equal(oracle.binning(f, 256), want)


# ------------------------------------------------------------------ (2) appendix A known-answer vectors
@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT])
def test_kat_sum3(oracle, b):
    got = oracle.local_op(cases.sum_domain_spec(3, b), cases.KAT_IMG)
    assert got.ravel().tolist() == cases.KAT_SUM3[b]


@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.REPEAT])
def test_kat_sum5(oracle, b):
    got = oracle.local_op(cases.sum_domain_spec(5, b), cases.KAT_IMG)
    assert got.ravel().tolist() == cases.KAT_SUM5[b]


def test_kat_constant_emitted_semantics(oracle):
    # emitted code uses the real constant: DSL-mode result + 7 * (#out-of-image taps); corner = 26 + 5*7
    got = oracle.local_op(cases.sum_domain_spec(3, A.CONSTANT, 7), cases.KAT_IMG)
    base = np.array(cases.KAT_SUM3[A.CONSTANT]).reshape(4, 6)
    n_out = np.full((4, 6), 0)
    n_out[0, :] += 3; n_out[-1, :] += 3; n_out[:, 0] += 3; n_out[:, -1] += 3
    n_out[0, 0] -= 1; n_out[0, -1] -= 1; n_out[-1, 0] -= 1; n_out[-1, -1] -= 1
    np.testing.assert_array_equal(got, base + 7 * n_out)
    assert got[0, 0] == 61


@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT])
def test_kat_single_tap(oracle, b):
    got = oracle.local_op(cases.single_tap_spec(-2, 2, 5, b), cases.KAT_IMG)
    assert got.ravel().tolist() == cases.KAT_TAP_M2P2[b]


def test_kat_interpolation(oracle):
    img = np.arange(32, dtype=np.float32).reshape(4, 8)
    nn = oracle.point_op(A.POINT_COPY, [img], A.F32, (2, 4), [A.INTERP_NN])
    assert nn.ravel().tolist() == cases.KAT_NN_8x4  # NN by 2 selects pixel (2x+1, 2y+1)
    lf = oracle.point_op(A.POINT_COPY, [nn], A.F32, (4, 8), [A.INTERP_LF])
    assert lf.ravel().tolist() == cases.KAT_LF_4x2_to_8x4
    img = np.arange(35, dtype=np.float32).reshape(5, 7)
    nn = oracle.point_op(A.POINT_COPY, [img], A.F32, (2, 3), [A.INTERP_NN])
    assert nn.ravel().tolist() == cases.KAT_NN_7x5


def test_reduction_hazard_documented(oracle):
    # the serial float fold drifts from the float64 sum; MIN/MAX do not care
    img = synth.image_np("float32", 1024, 1024, seed=4)
    s64 = img.astype(np.float64).sum()
    serial = float(oracle.reduce_serial(img, A.SUM))
    assert abs(serial - s64) / s64 > 1e-6          # the reference's own order is not 1e-5-stable at scale
    _, _, sm, s64b = oracle.reduce_minmaxsum(img)
    assert abs(float(sm) - s64) / s64 < 1e-5 and abs(s64b - s64) / s64 < 1e-12


def test_bilateral_mask_matches_sample_tables():
    m = M.bilateral_mask(13)
    assert m.shape == (13, 13)
    # spot values of Bilateral_Filter/src/main.cpp:128-140
    assert np.float32(0.018316) == m[0, 0] and np.float32(0.033746) == m[0, 1]
    assert np.float32(0.945959) == m[6, 5] and np.float32(1.0) == m[6, 6] and np.float32(0.894839) == m[5, 5]
    m5 = M.bilateral_mask(5)
    assert np.float32(0.082085) == m5[0, 1] and np.float32(0.606531) == m5[1, 2]


# ------------------------------------------------------------------ (3) live compiled reference
ROIS = [  # (iteration space roi, accessor roi) over the U8_SHAPE image (w=70,h=45)
    ((30, 20, 5, 7), (30, 20, 11, 3)),
    ((70, 1, 0, 44), (70, 1, 0, 0)),     # single row
    ((1, 45, 69, 0), (1, 45, 3, 0)),     # single column
    ((17, 13, 53, 32), (17, 13, 0, 0)),  # ragged corner
]


@pytest.mark.parametrize("b", [A.CLAMP, A.MIRROR, A.CONSTANT])
@pytest.mark.parametrize("roi", ROIS)
def test_roi_vs_reference(ref, b, roi):
    u8 = cases.inputs()["u8"]
    ris, racc = roi
    if b == A.MIRROR and min(racc[0], racc[1]) < 2:
        pytest.skip("single reflection with halo > window reads outside the window in the reference (undefined)")
    base = synth.image_np("uint8", u8.shape[1], u8.shape[0], seed=9)
    want = ref.ref_gaussian_u8(u8, M.GAUSS5, b, ris, racc, out=base.copy())
    got = ref.local_op(S.gaussian_blur(M.GAUSS5, b), u8, out=base.copy(), roi_in=racc, roi_out=ris)
    np.testing.assert_array_equal(got, want)   # pixels outside the iteration space untouched, too


def test_randomised_sweep_vs_reference(ref):
    """Seeded sweep over image shapes (incl. smaller than the halo is excluded for MIRROR / REPEAT, where the DSL's
    single reflection / single wrap is undefined), boundary modes and operators: the restated oracle must equal the
    compiled reference DSL pixel for pixel.  ~40 cases, a few hundred kpx in total."""
    rng = np.random.default_rng(2024)
    ops = ("gauss3", "gauss5", "gauss7", "dom3", "dom5", "conv5", "max3", "min5", "box3")
    for case in range(40):
        op = ops[case % len(ops)]
        size = int(op[-1])
        b = int(rng.choice([A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT]))
        lo = 1 if b in (A.CLAMP, A.CONSTANT) else size // 2 + 1   # halo <= window for MIRROR / REPEAT
        h, w = int(rng.integers(lo, 41)), int(rng.integers(lo, 67))
        seed = 1000 + case
        if op.startswith("gauss"):
            img = synth.image_np("uint8", w, h, seed=seed)
            want = ref.ref_gaussian_u8(img, M.GAUSS[size], b)
            got = ref.local_op(S.gaussian_blur(M.GAUSS[size], b), img)
        elif op in ("dom3", "dom5", "conv5"):
            img = synth.image_np("float32", w, h, seed=seed)
            mask = {"dom3": M.SOBEL3_Y, "dom5": M.SOBEL5_X, "conv5": M.GAUSS5}[op].astype(np.float32)
            use_dom = op != "conv5"
            want = ref.ref_local_f32(img, mask, use_dom, A.SUM, b)
            got = ref.local_op(S.domain_reduce_f32(mask, b) if use_dom else S.convolve_f32(mask, b), img)
        elif op in ("max3", "min5"):
            img = synth.image_np("uint8", w, h, seed=seed)
            want = ref.ref_minmax_u8(img, size, size, op == "max3", b)
            got = ref.local_op(S.minmax_u8(size, size, op == "max3", b), img)
        else:
            img = synth.image_np("uint8", w, h, seed=seed)
            want = ref.ref_box_u8(img, size, size, b)
            got = ref.local_op(S.box_blur_u8(size, size, b), img)
        np.testing.assert_array_equal(got, want, err_msg=f"case {case}: {op} {h}x{w} boundary {b}")


def test_randomised_sweep_rgba_and_histogram_vs_reference(ref):
    """uchar4 operators (per-channel oracle) and the histogram against the RGBA / Histogram samples run by the reference
    DSL, over random shapes and boundary modes"""
    rng = np.random.default_rng(99)
    for case in range(16):
        b = int(rng.choice([A.CLAMP, A.MIRROR, A.REPEAT, A.CONSTANT]))
        lo = 1 if b in (A.CLAMP, A.CONSTANT) else 4
        h, w = int(rng.integers(lo, 30)), int(rng.integers(lo, 45))
        img = cases.rgba_image(h, w, seed=2000 + case)
        kind = case % 4
        if kind == 0:
            sz = int(rng.choice([3, 5, 7])) if lo == 4 or min(h, w) >= 1 else 3
            want, got = ref.ref_gaussian_rgba(img, M.GAUSS[sz], b), ref.local_op_x4(S.gaussian_blur(M.GAUSS[sz], b), img)
        elif kind == 1:
            want, got = ref.ref_laplace_rgba(img, M.LAPLACE5, b), ref.local_op_x4(S.laplace_u8(M.LAPLACE5, b, add=0), img)
        elif kind == 2:
            want, got = ref.ref_dilate_rgba(img, 5, 5, b), ref.local_op_x4(S.minmax_u8(5, 5, True, b), img)
        else:
            want, got = ref.ref_box_rgba(img, 3, 3, b), ref.local_op_x4(S.box_blur_u8(3, 3, b), img)
        np.testing.assert_array_equal(got, want, err_msg=f"case {case}: kind {kind} {h}x{w} boundary {b}")
    for case in range(6):
        h, w, nb = int(rng.integers(1, 60)), int(rng.integers(1, 90)), int(rng.choice([2, 17, 256, 999]))
        img = synth.image_np("float32", w, h, seed=2100 + case, scale=254.99)
        np.testing.assert_array_equal(ref.binning(img, nb), ref.ref_sample_histogram_f32(img, nb), err_msg=f"hist {h}x{w} bins {nb}")


def test_repeat_divergence_with_offset_accessor(ref):
    """DSL repeat adds lower+upper once (dsl/image.hpp:296-300); emitted code loops +-size
    (lib/AST/BorderHandling.cpp:59-74).  Equal when lower == 0, different for an offset window:
    the contract follows emitted code."""
    u8 = cases.inputs()["u8"]
    full = ref.ref_gaussian_u8(u8, M.GAUSS5, A.REPEAT)
    np.testing.assert_array_equal(ref.local_op(S.gaussian_blur(M.GAUSS5, A.REPEAT), u8), full)
    ris, racc = (30, 20, 5, 7), (30, 20, 11, 3)
    want = ref.ref_gaussian_u8(u8, M.GAUSS5, A.REPEAT, ris, racc)
    got = ref.local_op(S.gaussian_blur(M.GAUSS5, A.REPEAT), u8, roi_in=racc, roi_out=ris)
    assert (got != want).any()
    # emitted semantics == periodic extension of the cropped window
    crop = np.ascontiguousarray(u8[3:23, 11:41])
    per = ref.local_op(S.gaussian_blur(M.GAUSS5, A.REPEAT), crop)
    np.testing.assert_array_equal(got[7:27, 5:35], per)


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (7, 1), (2, 3), (5, 5)])
@pytest.mark.parametrize("b", [A.CLAMP, A.CONSTANT])
def test_tiny_images_vs_reference(ref, shape, b):
    img = synth.image_np("uint8", shape[1], shape[0], seed=11)
    want = ref.ref_gaussian_u8(img, M.GAUSS3, b)
    np.testing.assert_array_equal(ref.local_op(S.gaussian_blur(M.GAUSS3, b), img), want)
    imgf = synth.image_np("float32", shape[1], shape[0], seed=12)
    wantf = ref.ref_local_f32(imgf, M.LAPLACE3.astype(np.float32), 1, A.SUM, b)
    np.testing.assert_array_equal(ref.local_op(S.domain_reduce_f32(M.LAPLACE3.astype(np.float32), b), imgf), wantf)


@pytest.mark.parametrize("mode", [A.SUM, A.MIN, A.MAX, A.PROD])
def test_reduce_modes_vs_reference(ref, mode):
    f32 = cases.inputs()["f32"]
    m = (M.GAUSS3 * 3.0).astype(np.float32)
    for dom in (0, 1):
        want = ref.ref_local_f32(f32, m, dom, mode, A.MIRROR)
        spec = S.domain_reduce_f32(m, A.MIRROR, mode) if dom else S.convolve_f32(m, A.MIRROR, mode)
        np.testing.assert_array_equal(ref.local_op(spec, f32), want)


def test_histogram_vs_reference_and_sample_checker(ref):
    img = synth.image_np("float32", 130, 77, seed=14, scale=254.99)
    for nb in (256, 32):
        want = ref.ref_sample_histogram_f32(img, nb)
        np.testing.assert_array_equal(ref.binning(img, nb), want)
        np.testing.assert_array_equal(ref.ref_sample_histogram_check(img, nb), want)


def test_pyramid_char_sample_restores_input(ref):
    # the sample itself (char pixels): restored input equals the input within the sample's tolerance
    img = synth.image_np("int8", 96, 64, seed=13)
    g, _ = ref.ref_pyramid_s8(img, 4, M.GAUSS5)
    assert np.abs(g[0].astype(np.int32) - img.astype(np.int32)).max() <= 1


# ------------------------------------------------------------------ (4) the samples' embedded plain-C checkers
def _interior(a, r):
    return a[r:-r, r:-r]


def test_samples_plain_c_checkers(ref):
    lib = ref.ref_lib()
    u8 = cases.inputs()["u8"]
    h, w = u8.shape
    pu8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_ubyte))
    # Gaussian (Gaussian_Blur/src/main.cpp:174-194): float sum starts at 0.5f -> may differ by 1 LSB from the
    # DSL's (uchar)(sum+0.5f); the sample itself tolerates |diff| <= 1 (hipacc_helper.hpp:190-196)
    out = np.zeros_like(u8)
    m = np.ascontiguousarray(M.GAUSS5)
    lib.ref_sample_gaussian_filter(pu8(u8), pu8(out), m.ctypes.data_as(C.POINTER(C.c_float)), 5, 5, w, h)
    mine = ref.local_op(S.gaussian_blur(M.GAUSS5, A.CLAMP), u8)
    assert np.abs(_interior(mine, 2).astype(int) - _interior(out, 2).astype(int)).max() <= 1
    # Laplace (Laplace/src/main.cpp:181-207): integer -> exact
    out = np.zeros_like(u8)
    mi = np.ascontiguousarray(M.LAPLACE5)
    lib.ref_sample_laplace_filter(pu8(u8), pu8(out), mi.ctypes.data_as(C.POINTER(C.c_int)), 5, w, h)
    np.testing.assert_array_equal(_interior(ref.local_op(S.laplace_u8(M.LAPLACE5), u8), 2), _interior(out, 2))
    # Sobel + combine (Sobel/src/main.cpp:257-289)
    sx, sy = np.zeros(u8.shape, np.int32), np.zeros(u8.shape, np.int32)
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    mx, my = np.ascontiguousarray(M.SOBEL3_X), np.ascontiguousarray(M.SOBEL3_Y)
    lib.ref_sample_sobel_filter(pu8(u8), pi(sx), pi(mx), 3, 3, w, h)
    lib.ref_sample_sobel_filter(pu8(u8), pi(sy), pi(my), 3, 3, w, h)
    ox, oy = ref.local_op(S.sobel_u8(M.SOBEL3_X), u8), ref.local_op(S.sobel_u8(M.SOBEL3_Y), u8)
    np.testing.assert_array_equal(_interior(ox, 1), _interior(sx, 1))
    np.testing.assert_array_equal(_interior(oy, 1), _interior(sy, 1))
    comb = np.zeros_like(u8)
    lib.ref_sample_sobel_combine(pi(ox), pi(oy), pu8(comb), w, h, 4)
    np.testing.assert_array_equal(ref.point_op(A.POINT_SOBEL_COMBINE, [ox, oy], A.U8, p=(4, 0)), comb)
    # Bilateral (Bilateral_Filter/src/main.cpp:199-224)
    out = np.zeros_like(u8)
    bm = np.ascontiguousarray(M.bilateral_mask(5))
    lib.ref_sample_bilateral_filter(pu8(u8), pu8(out), bm.ctypes.data_as(C.POINTER(C.c_float)), 5, 16, w, h)
    mine = ref.bilateral(u8, 5, bm, 16, A.CLAMP)
    assert np.abs(_interior(mine, 2).astype(int) - _interior(out, 2).astype(int)).max() <= 1
    # Reduction (Reduction_Sum/src/main.cpp:122-126) == serial fold starting at 0
    f32 = cases.inputs()["f32"]
    r = C.c_float()
    lib.ref_sample_reduction(f32.ctypes.data_as(C.POINTER(C.c_float)), C.byref(r), f32.shape[1], f32.shape[0])
    assert abs(r.value - float(ref.reduce_serial(f32, A.SUM))) <= 1e-3 * abs(r.value)


def test_synth_numpy_equals_torch():
    import torch
    for dt in ("uint8", "float32"):
        a = synth.image_np(dt, 37, 19, seed=5, x0=3, y0=100)
        b = synth.image_torch(dt, 37, 19, seed=5, x0=3, y0=100, rows_per_chunk=7).numpy()
        np.testing.assert_array_equal(a, b)
    full = synth.image_np("float32", 16, 32, seed=1)
    strip = synth.image_np("float32", 16, 8, seed=1, y0=8)
    np.testing.assert_array_equal(full[8:16], strip)


# ------------------------------------------------------------------ the TIMED CPU leg (oracle/emit_cpu_fast.cpp)
@pytest.mark.parametrize("b", [A.CLAMP, A.REPEAT, A.MIRROR, A.CONSTANT, A.UNDEFINED])
@pytest.mark.parametrize("shape", [(39, 66), (1, 1), (2, 3), (131, 257), (64, 1031)])
def test_fast_cpu_leg_equals_generic_oracle(oracle, b, shape):
    """The specialised loops bench.py times on the host (constexpr masks, interior / border split, AVX2) are the SAME
    function as the generic checker: bit-for-bit on ragged sizes, every boundary mode, ROIs and ghost rows."""
    h, w = shape
    f = synth.image_np("float32", w, h, seed=71)
    u = synth.image_np("uint8", w, h, seed=72)
    f32_specs = [S.domain_reduce_f32(m.astype(np.float32), b, const=0.25) for m in (M.SOBEL3_X, M.SOBEL3_Y, M.LAPLACE3, np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]]))]
    f32_specs += [S.convolve_f32(M.GAUSS[sz], b, const=0.25) for sz in (3, 5, 7)]
    f32_specs += [S.convolve_f32(M.SOBEL3_X.astype(np.float32), b, const=0.25)]   # convolve() visits the zero taps too
    for spec in f32_specs:
        np.testing.assert_array_equal(oracle.local_op_fast(spec, f), oracle.local_op(spec, f))
    for sz in (3, 5, 7):
        spec = S.gaussian_blur(M.GAUSS[sz], b)
        np.testing.assert_array_equal(oracle.local_op_fast(spec, u), oracle.local_op(spec, u))
    if h > 20 and w > 40:   # ROI + crop accessor + ghost rows
        roi_in, roi_out, ghost = (w - 9, h - 12, 4, 6), (w - 9, h - 12, 5, 3), (2, 3)
        for spec in (f32_specs[0], f32_specs[5]):
            o1, o2 = np.zeros_like(f), np.zeros_like(f)
            oracle.local_op(spec, f, out=o1, roi_in=roi_in, roi_out=roi_out, ghost=ghost)
            oracle.local_op_fast(spec, f, out=o2, roi_in=roi_in, roi_out=roi_out, ghost=ghost)
            np.testing.assert_array_equal(o1, o2)


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (7, 1), (2, 2), (37, 61), (130, 257)])
def test_fast_cpu_harris_equals_generic_oracle(oracle, shape):
    """The timed CPU leg of C4 (nine specialised kernels, emit_cpu_fast.cpp::ocf_harris) equals the generic oracle pipeline
    bit for bit: noise (every stage far from trivial), blocks (corners fire), extreme values, pitched rows."""
    h, w = shape
    for img in (synth.image_np("uint8", w, h, seed=81), synth.blocks_np(w, h, seed=82), np.full((h, w), 255, np.uint8),
                (np.indices((h, w)).sum(0) % 2 * 255).astype(np.uint8)):
        np.testing.assert_array_equal(oracle.harris_fast(img), oracle.harris(img))
    pitched = np.zeros((h, w + 13), np.uint8)
    pitched[:, :w] = synth.image_np("uint8", w, h, seed=83)
    np.testing.assert_array_equal(oracle.harris_fast(pitched[:, :w]), oracle.harris(np.ascontiguousarray(pitched[:, :w])))


def _harris_extreme_images():
    """images that drive the Sobel sums of the Harris pipeline to their extremes: binary noise, a diagonal step (both
    derivatives at 510 = the largest |dx*6| the corner sharing allows next to an equal |dy*6|), checkerboards, stripes"""
    rng = np.random.default_rng(99)
    h, w = 96, 160
    yy, xx = np.indices((h, w))
    imgs = [(rng.integers(0, 2, size=(h, w)) * 255).astype(np.uint8) for _ in range(4)]
    imgs += [((xx + yy) > 120).astype(np.uint8) * 255, ((xx - yy) > 20).astype(np.uint8) * 255,
             ((xx + yy) % 2 * 255).astype(np.uint8), ((xx // 2 + yy // 2) % 2 * 255).astype(np.uint8),
             ((xx % 3 == 0) * 255).astype(np.uint8), ((yy % 3 == 0) * 255).astype(np.uint8),
             (rng.integers(0, 2, size=(h, w)) * 255 * ((xx + yy) > 100)).astype(np.uint8)]
    return imgs


def test_harris_product_bound_behind_the_biased_xy_plane(oracle):
    """The fused kernel stores dx*dy with a bias of 8192 in an unsigned 16-bit plane and sums three rows of it in packed
    16-bit halves: that needs |dx * dy| <= 7225 (the two Sobel sums share their corner pixels: |dx*6| + |dy*6| <= 1020).
    Checked here on the oracle's intermediates for images built to hit the extremes; the diagonal step reaches the bound."""
    worst = 0
    for img in _harris_extreme_images():
        dx = oracle.local_op(S.harris_deriv(M.HARRIS_DX), img).astype(np.int64)
        dy = oracle.local_op(S.harris_deriv(M.HARRIS_DY), img).astype(np.int64)
        assert np.abs(dx).max() <= 127 and np.abs(dy).max() <= 127
        assert (np.abs(dx) * 6 + np.abs(dy) * 6).max() <= 1020 + 10          # quotients are truncated: |q| * 6 <= |d * 6|
        worst = max(worst, int(np.abs(dx * dy).max()))
    assert worst <= 7225
    assert worst == 7225      # the diagonal step attains it: the bound is tight, not merely safe


def test_harris_oracle_and_fast_leg_equal_the_reference_dsl_on_extreme_images(ref):
    """The compiled reference DSL (the sample's own nine kernel classes, oracle/_ref) on the extreme-gradient images and on
    noise: the restated pipeline and the specialised CPU leg give the same corners and the same three smoothed planes."""
    imgs = _harris_extreme_images() + [synth.image_np("uint8", 131, 77, seed=s) for s in (31, 32)]
    for k, img in enumerate(imgs):
        want, gx, gy, gxy = ref.ref_harris_u8(img)
        got, ogx, ogy, ogxy = ref.harris(img, return_intermediates=True)
        np.testing.assert_array_equal(ogx, gx, err_msg=f"image {k}")
        np.testing.assert_array_equal(ogy, gy, err_msg=f"image {k}")
        np.testing.assert_array_equal(ogxy, gxy, err_msg=f"image {k}")
        np.testing.assert_array_equal(got, want, err_msg=f"image {k}")
        np.testing.assert_array_equal(ref.harris_fast(img), want, err_msg=f"image {k} (fast leg)")


def test_fast_cpu_leg_refuses_what_it_does_not_specialise(oracle):
    u = synth.image_np("uint8", 40, 30, seed=73)
    with pytest.raises(RuntimeError):
        oracle.local_op_fast(S.sobel_u8(M.SOBEL3_X), u)          # integer accumulate: generic checker only
