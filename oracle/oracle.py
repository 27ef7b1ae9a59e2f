"""TEST INFRASTRUCTURE ONLY -- numpy front of the CPU oracle and of the compiled reference.

Two checkers live behind this module:
  * ``emit``  : oracle/liboracle_emitcpu.so, our restatement of Hipacc's -emit-cpu code
                (oracle/emit_cpu.cpp), driven by the product's own C-ABI descriptors with
                host pointers.  Also the timed CPU baseline ("port").
  * ``ref``   : oracle/_ref/libhipacc_ref.so, the reference's OWN DSL headers and sample
                kernel classes compiled where they lie under /root/reference
                (oracle/ref_dsl.cpp).  Spec of record, ~1 Mpx/s.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (hipacc_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from hipacc_b200 import _abi as A
from hipacc_b200 import specs as S
from hipacc_b200 import masks as M

HERE = os.path.dirname(os.path.abspath(__file__))
_EMIT_PATH = os.path.join(HERE, "liboracle_emitcpu.so")
_FAST_PATH = os.path.join(HERE, "liboracle_emitcpu_fast.so")
_REF_PATH = os.path.join(HERE, "_ref", "libhipacc_ref.so")


def build(force=False):
    """Compile the checkers (g++).  _ref is only (re)built where /root/reference exists."""
    args = ["make", "-C", HERE, "all"] + (["-B"] if force else [])
    subprocess.run(args, check=True, capture_output=True)


_emit = None
_fast = None
_ref = None


def emit_lib():
    global _emit
    if _emit is None:
        if not os.path.exists(_EMIT_PATH):
            build()
        _emit = C.CDLL(_EMIT_PATH)
        _emit.oc_local_op.argtypes = [C.POINTER(A.hb_local_desc)]
        _emit.oc_bilateral.argtypes = [C.POINTER(A.hb_bilateral_desc)]
        _emit.oc_point_op.argtypes = [C.POINTER(A.hb_point_desc)]
        _emit.oc_reduce_serial_f32.argtypes = [C.POINTER(A.hb_view), C.c_int, C.POINTER(C.c_float)]
        _emit.oc_binning.argtypes = [C.POINTER(A.hb_binning_desc), C.POINTER(C.c_uint)]
        _emit.oc_reduce_minmaxsum_f32.argtypes = [C.POINTER(A.hb_view), C.POINTER(C.c_float), C.POINTER(C.c_double)]
    return _emit


def fast_lib():
    """oracle/liboracle_emitcpu_fast.so: the specialised (-emit-cpu shaped) loops of the TIMED CPU leg (emit_cpu_fast.cpp)"""
    global _fast
    if _fast is None:
        if not os.path.exists(_FAST_PATH):
            build()
        _fast = C.CDLL(_FAST_PATH)
        _fast.ocf_local_op.argtypes = [C.POINTER(A.hb_local_desc)]
        _fast.ocf_harris.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
    return _fast


def have_ref():
    if not os.path.exists(_REF_PATH) and os.path.isdir("/root/reference/dsl"):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(_REF_PATH)


def ref_lib():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libhipacc_ref.so not built (needs /root/reference at build time)")
        _ref = C.CDLL(_REF_PATH)
    return _ref


def num_threads():
    return emit_lib().oc_num_threads()


def set_num_threads(n):
    emit_lib().oc_set_num_threads(int(n))
    fast_lib().ocf_set_num_threads(int(n))


# ----------------------------------------------------------------------------- helpers
def np_view(a, roi=None, ghost=(0, 0)):
    assert a.ndim == 2 and a.flags.c_contiguous
    return A.make_view(a.ctypes.data, A.NUMPY_DTYPE[a.dtype.name], a.shape[1], a.shape[0], a.shape[1], roi, ghost)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed with status {rc}")


def _roi8(roi_is, roi_acc):
    r = (C.c_int * 8)(*([0] * 8))
    if roi_is is not None:
        r[0:4] = list(roi_is)
    if roi_acc is not None:
        r[4:8] = list(roi_acc)
    return r


# ----------------------------------------------------------------------------- emit (restated oracle)
def local_op(spec: S.LocalSpec, img, out=None, roi_in=None, roi_out=None, ghost=(0, 0)):
    """Run a local operator on a numpy image.  `out` (optional) is written in place inside roi_out."""
    if out is None:
        out = np.zeros(img.shape, dtype=A.DTYPE_NUMPY[spec.out_dtype])
    d = A.hb_local_desc()
    spec.fill(d)
    d.in_ = np_view(img, roi_in, ghost)
    d.out = np_view(out, roi_out)
    _check(emit_lib().oc_local_op(C.byref(d)), "local_op")
    return out


def local_op_fast(spec: S.LocalSpec, img, out=None, roi_in=None, roi_out=None, ghost=(0, 0)):
    """The same operator through the specialised CPU loops (emit_cpu_fast.cpp); raises when no specialisation exists."""
    if out is None:
        out = np.zeros(img.shape, dtype=A.DTYPE_NUMPY[spec.out_dtype])
    d = A.hb_local_desc()
    spec.fill(d)
    d.in_ = np_view(img, roi_in, ghost)
    d.out = np_view(out, roi_out)
    _check(fast_lib().ocf_local_op(C.byref(d)), "local_op_fast")
    return out


def bilateral(img, size, coef, sigma_r, boundary=A.CLAMP, const=0.0):
    out = np.zeros_like(img)
    d = A.hb_bilateral_desc()
    c = np.ascontiguousarray(coef, dtype=np.float32)
    d.in_, d.out = np_view(img), np_view(out)
    d.size, d.coef_f32, d.sigma_r = size, c.ctypes.data_as(C.POINTER(C.c_float)), sigma_r
    d.boundary, d.boundary_const = boundary, const
    _check(emit_lib().oc_bilateral(C.byref(d)), "bilateral")
    return out


def point_op(op, inputs, out_dtype, out_shape=None, interp=None, p=(0.0, 0.0), out=None):
    if out is None:
        out = np.zeros(out_shape if out_shape is not None else inputs[0].shape, dtype=A.DTYPE_NUMPY[out_dtype])
    d = A.hb_point_desc()
    d.n_in = len(inputs)
    for i, a in enumerate(inputs):
        d.in_[i] = np_view(a)
        d.interp[i] = (interp[i] if interp else A.INTERP_NO)
    d.out = np_view(out)
    d.op = op
    d.p[0], d.p[1] = float(p[0]), float(p[1])
    _check(emit_lib().oc_point_op(C.byref(d)), "point_op")
    return out


def reduce_serial(img, mode, roi=None):
    v = np_view(img, roi)
    r = C.c_float()
    _check(emit_lib().oc_reduce_serial_f32(C.byref(v), mode, C.byref(r)), "reduce_serial")
    return np.float32(r.value)


def reduce_minmaxsum(img, roi=None):
    """-> (min, max, float32 sum in -emit-cpu order with the current thread count, float64 sum)"""
    v = np_view(img, roi)
    o = (C.c_float * 3)()
    s = C.c_double()
    _check(emit_lib().oc_reduce_minmaxsum_f32(C.byref(v), o, C.byref(s)), "reduce_minmaxsum")
    return np.float32(o[0]), np.float32(o[1]), np.float32(o[2]), s.value


def local_op_x4(spec, img, **kw):
    """uchar4 image [H, W, 4]: the DSL's float4 / int4 arithmetic and convert_uchar4() are element-wise
    (dsl/types.hpp:56-516), so a local operator on uchar4 pixels is the scalar operator on each channel plane."""
    planes = [local_op(spec, np.ascontiguousarray(img[..., c]), **kw) for c in range(4)]
    return np.ascontiguousarray(np.stack(planes, axis=-1))   # the spec's output type: uchar4 -> uchar4 / short4 / int4, float4 -> float4


def binning(img, num_bins, index_kind=A.BIN_INDEX_SCALE, value_kind=A.BIN_VALUE_ONE, p0=255.0, roi=None):
    """-emit-cpu binning (BINNING_CPU_2D, runtime/hipacc_cpu_red.hpp:70-128) -> uint32[num_bins]"""
    d = A.hb_binning_desc()
    d.in_ = np_view(img, roi)
    d.num_bins, d.index_kind, d.value_kind, d.p0 = int(num_bins), index_kind, value_kind, float(p0)
    out = np.zeros(num_bins, dtype=np.uint32)
    _check(emit_lib().oc_binning(C.byref(d), out.ctypes.data_as(C.POINTER(C.c_uint))), "binning")
    return out


def harris(img, k=M.HARRIS_K, threshold=M.HARRIS_THRESHOLD, return_intermediates=False):
    """The sample's 9-kernel pipeline (Harris_Corner/src/main.cpp:230-305) composed from oracle ops."""
    dx = local_op(S.harris_deriv(M.HARRIS_DX), img)
    dy = local_op(S.harris_deriv(M.HARRIS_DY), img)
    sx = point_op(A.POINT_SQUARE, [dx], A.S16)
    sy = point_op(A.POINT_SQUARE, [dy], A.S16)
    sxy = point_op(A.POINT_MUL, [dx, dy], A.S16)
    gx = local_op(S.harris_gauss(M.HARRIS_GAUSS3), sx)
    gy = local_op(S.harris_gauss(M.HARRIS_GAUSS3), sy)
    gxy = local_op(S.harris_gauss(M.HARRIS_GAUSS3), sxy)
    out = point_op(A.POINT_HARRIS, [gx, gy, gxy], A.U8, p=(k, threshold))
    if return_intermediates:
        return out, gx, gy, gxy
    return out


def harris_fast(img, k=M.HARRIS_K, threshold=M.HARRIS_THRESHOLD, out=None):
    """The same nine kernels through the specialised CPU loops (emit_cpu_fast.cpp::ocf_harris): the TIMED CPU leg of C4."""
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    if out is None:
        out = np.empty(img.shape, dtype=np.uint8)
    _check(fast_lib().ocf_harris(img.ctypes.data, out.ctypes.data, img.shape[1], img.shape[0], img.strides[0], out.strides[0], k, threshold), "harris_fast")
    return out


def pyramid(img, depth, mask):
    """Gaussian/Laplacian pyramid traversal of Gaussian_Laplacian_Pyramid/src/main.cpp:199-248
    (float pixels), composed from oracle ops in the sample's order.  -> (gaus levels, lap levels)"""
    sizes = S.pyramid_sizes(img.shape[1], img.shape[0], depth)
    gaus = [img.copy()] + [np.zeros((h, w), np.float32) for (w, h) in sizes[1:]]
    lap = [np.zeros((h, w), np.float32) for (w, h) in sizes]
    blur = S.convolve_f32(mask, A.CLAMP)
    for l in range(1, depth):  # way down
        tmp = local_op(blur, gaus[l - 1])
        gaus[l] = point_op(A.POINT_COPY, [tmp], A.F32, gaus[l].shape, [A.INTERP_NN])
        lap[l - 1] = point_op(A.POINT_SUB, [gaus[l - 1], gaus[l]], A.F32, gaus[l - 1].shape,
                              [A.INTERP_NO, A.INTERP_LF])
    for l in range(depth - 2, -1, -1):  # way up
        gaus[l] = point_op(A.POINT_ADD, [gaus[l + 1], lap[l]], A.F32, gaus[l].shape, [A.INTERP_LF, A.INTERP_NO])
        lap[l] = point_op(A.POINT_BLEND, [lap[l + 1], lap[l]], A.F32, lap[l].shape, [A.INTERP_LF, A.INTERP_NO])
    return gaus, lap


# ----------------------------------------------------------------------------- ref (reference DSL, compiled)
def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def ref_gaussian_u8(img, mask, boundary, roi_is=None, roi_acc=None, out=None):
    out = np.zeros_like(img) if out is None else out
    m = np.ascontiguousarray(mask, np.float32)
    rc = ref_lib().ref_gaussian_u8(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0],
                                   m.shape[1], m.shape[0], _p(m, C.c_float), boundary, _roi8(roi_is, roi_acc))
    _check(rc, "ref_gaussian_u8")
    return out


def ref_local_f32(img, mask, use_domain, mode, boundary, roi_is=None, roi_acc=None, out=None):
    out = np.zeros_like(img) if out is None else out
    m = np.ascontiguousarray(mask, np.float32)
    rc = ref_lib().ref_local_f32(_p(img, C.c_float), _p(out, C.c_float), img.shape[1], img.shape[0],
                                 m.shape[1], m.shape[0], _p(m, C.c_float), int(use_domain), mode, boundary,
                                 _roi8(roi_is, roi_acc))
    _check(rc, "ref_local_f32")
    return out


def ref_sobel_u8(img, mask, boundary, roi_is=None, roi_acc=None):
    out = np.zeros(img.shape, np.int32)
    m = np.ascontiguousarray(mask, np.int32)
    rc = ref_lib().ref_sobel_u8(_p(img, C.c_ubyte), _p(out, C.c_int), img.shape[1], img.shape[0],
                                m.shape[1], m.shape[0], _p(m, C.c_int), boundary, _roi8(roi_is, roi_acc))
    _check(rc, "ref_sobel_u8")
    return out


def ref_sobel_combine(a, b, norm):
    out = np.zeros(a.shape, np.uint8)
    _check(ref_lib().ref_sobel_combine(_p(a, C.c_int), _p(b, C.c_int), _p(out, C.c_ubyte), a.shape[1], a.shape[0],
                                       int(norm)), "ref_sobel_combine")
    return out


def ref_laplace_u8(img, mask, boundary, roi_is=None, roi_acc=None):
    out = np.zeros_like(img)
    m = np.ascontiguousarray(mask, np.int32)
    rc = ref_lib().ref_laplace_u8(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0],
                                  m.shape[1], m.shape[0], _p(m, C.c_int), boundary, _roi8(roi_is, roi_acc))
    _check(rc, "ref_laplace_u8")
    return out


def ref_minmax_u8(img, sx, sy, is_max, boundary):
    out = np.zeros_like(img)
    _check(ref_lib().ref_minmax_u8(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0], sx, sy,
                                   int(is_max), boundary), "ref_minmax_u8")
    return out


def ref_box_u8(img, sx, sy, boundary):
    out = np.zeros_like(img)
    _check(ref_lib().ref_box_u8(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0], sx, sy,
                                boundary), "ref_box_u8")
    return out


def ref_bilateral(img, size, coef, sigma_r, boundary):
    out = np.zeros_like(img)
    c = np.ascontiguousarray(coef, np.float32)
    if img.dtype == np.uint8:
        rc = ref_lib().ref_bilateral_u8(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0], size,
                                        _p(c, C.c_float), sigma_r, boundary)
    else:
        rc = ref_lib().ref_bilateral_f32(_p(img, C.c_float), _p(out, C.c_float), img.shape[1], img.shape[0], size,
                                         _p(c, C.c_float), sigma_r, boundary)
    _check(rc, "ref_bilateral")
    return out


def ref_tap_u8(img, dx, dy, boundary, wx, wy):
    out = np.zeros_like(img)
    _check(ref_lib().ref_tap_u8(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0], dx, dy,
                                boundary, wx, wy), "ref_tap_u8")
    return out


def ref_harris_u8(img, k=M.HARRIS_K, threshold=M.HARRIS_THRESHOLD):
    h, w = img.shape
    out = np.zeros_like(img)
    gx, gy, gxy = (np.zeros((h, w), np.int16) for _ in range(3))
    rc = ref_lib().ref_harris_u8(_p(img, C.c_ubyte), _p(out, C.c_ubyte), w, h, C.c_float(k), C.c_float(threshold),
                                 _p(gx, C.c_short), _p(gy, C.c_short), _p(gxy, C.c_short))
    _check(rc, "ref_harris_u8")
    return out, gx, gy, gxy


def ref_interp_f32(img, ow, oh, imode):
    out = np.zeros((oh, ow), np.float32)
    _check(ref_lib().ref_interp_f32(_p(img, C.c_float), img.shape[1], img.shape[0], _p(out, C.c_float), ow, oh,
                                    imode), "ref_interp_f32")
    return out


def _unpack_levels(flat, sizes):
    out, off = [], 0
    for (w, h) in sizes:
        out.append(flat[off:off + w * h].reshape(h, w).copy())
        off += w * h
    return out


def ref_pyramid_f32(img, depth, mask):
    h, w = img.shape
    sizes = S.pyramid_sizes(w, h, depth)
    n = sum(a * b for a, b in sizes)
    g, l = np.zeros(n, np.float32), np.zeros(n, np.float32)
    m = np.ascontiguousarray(mask, np.float32)
    rc = ref_lib().ref_pyramid_f32(_p(img, C.c_float), w, h, depth, m.shape[0], _p(m, C.c_float), _p(g, C.c_float),
                                   _p(l, C.c_float), None)
    _check(rc, "ref_pyramid_f32")
    return _unpack_levels(g, sizes), _unpack_levels(l, sizes)


def ref_pyramid_s8(img, depth, mask):
    h, w = img.shape
    sizes = S.pyramid_sizes(w, h, depth)
    n = sum(a * b for a, b in sizes)
    g, l = np.zeros(n, np.int8), np.zeros(n, np.int8)
    m = np.ascontiguousarray(mask, np.float32)
    rc = ref_lib().ref_pyramid_s8(_p(img, C.c_char), w, h, depth, m.shape[0], _p(m, C.c_float), _p(g, C.c_char),
                                  _p(l, C.c_char))
    _check(rc, "ref_pyramid_s8")
    return _unpack_levels(g, sizes), _unpack_levels(l, sizes)


def ref_global_reduce_f32(img, op, roi=None):
    """DSL serial fold; op 0=sum 1=min 2=max"""
    r = C.c_float()
    _check(ref_lib().ref_global_reduce_f32(_p(img, C.c_float), img.shape[1], img.shape[0], op,
                                           _roi8(roi, None), C.byref(r)), "ref_global_reduce_f32")
    return np.float32(r.value)


def ref_gaussian_rgba(img, mask, boundary):
    """sample GaussianBlur of Gaussian_Blur_RGBA (Kernel<uchar4>, float4 accumulate) executed by the reference DSL"""
    out = np.zeros_like(img)
    m = np.ascontiguousarray(mask, np.float32)
    _check(ref_lib().ref_gaussian_rgba(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0], _p(m, C.c_float),
                                       m.shape[1], m.shape[0], boundary), "ref_gaussian_rgba")
    return out


def ref_laplace_rgba(img, mask, boundary):
    """sample LaplaceFilter of Laplace_RGBA (Kernel<uchar4>, int4 accumulate, +128 clamp) executed by the reference DSL"""
    out = np.zeros_like(img)
    m = np.ascontiguousarray(mask, np.int32)
    _check(ref_lib().ref_laplace_rgba(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0], _p(m, C.c_int),
                                      m.shape[0], boundary), "ref_laplace_rgba")
    return out


def ref_dilate_rgba(img, sx, sy, boundary):
    out = np.zeros_like(img)
    _check(ref_lib().ref_dilate_rgba(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0], sx, sy, boundary), "ref_dilate_rgba")
    return out


def ref_box_rgba(img, sx, sy, boundary):
    out = np.zeros_like(img)
    _check(ref_lib().ref_box_rgba(_p(img, C.c_ubyte), _p(out, C.c_ubyte), img.shape[1], img.shape[0], sx, sy, boundary), "ref_box_rgba")
    return out


def ref_sample_histogram_f32(img, num_bins):
    """the Histogram sample's Kernel (binning() + binned_data()) executed by the reference DSL"""
    out = np.zeros(num_bins, dtype=np.uint32)
    _check(ref_lib().ref_sample_histogram_f32(_p(img, C.c_float), img.shape[1], img.shape[0], int(num_bins),
                                              out.ctypes.data_as(C.POINTER(C.c_uint))), "ref_sample_histogram_f32")
    return out


def ref_sample_histogram_check(img, num_bins):
    """the sample's embedded plain-C histogram (Histogram/src/main.cpp:122-127)"""
    out = np.zeros(num_bins, dtype=np.uint32)
    _check(ref_lib().ref_sample_histogram_check(_p(img, C.c_float), out.ctypes.data_as(C.POINTER(C.c_uint)),
                                                img.shape[1], img.shape[0], int(num_bins)), "ref_sample_histogram_check")
    return out


def ref_rt_reduce_f32(img, op):
    """reference CPU runtime REDUCTION_CPU_2D (thread-count dependent for SUM)"""
    r = C.c_float()
    _check(ref_lib().ref_rt_reduce_f32(_p(img, C.c_float), img.shape[1], img.shape[0], img.shape[1], op, 0, 0,
                                       C.byref(r)), "ref_rt_reduce_f32")
    return np.float32(r.value)


def ref_sample_checker(name, *args):
    """The samples' embedded plain-C reference loops (interior pixels only)."""
    return getattr(ref_lib(), "ref_sample_" + name)(*args)
