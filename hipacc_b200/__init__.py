"""hipacc_b200 -- B200-native execution path for Hipacc operators.

The product is the C-ABI shared library ``hipacc_b200/lib/libhipacc_b200.so`` (hand-written
sm_100a CUDA kernels behind include/hipacc_b200.h).  This package is the thin Python host side
used by the tests and bench.py: it loads the library with ctypes and wraps device buffers
(torch tensors are used only as HBM allocations / streams -- plumbing, not compute).

There is NO CPU fallback: ``lib()`` raises if the CUDA library has not been built, and every
operator returns an error status (raised as HbError) when no device kernel exists.
"""
import ctypes as C
import os

from . import _abi as A
from . import masks, specs, synth  # noqa: F401  (re-exported)
from ._abi import *  # noqa: F401,F403  (enum constants)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HIPACC_B200_LIB") or os.path.join(_HERE, "lib", "libhipacc_b200.so")   # the override is for A/B builds (tools/)
_lib = None


class HbError(RuntimeError):
    def __init__(self, status, what, detail=""):
        super().__init__(f"{what} failed with status {status}: {detail}")
        self.status = status


def lib():
    """Load libhipacc_b200.so (once).  Fails loudly when the CUDA extension is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m hipacc_b200.build` "
                "(hipacc_b200 has no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        L.hb_last_error.restype = C.c_char_p
        L.hb_last_kernel_ms.restype = C.c_float
        L.hb_launch_count.restype = C.c_longlong
        L.hb_local_op.argtypes = [C.POINTER(A.hb_local_desc), C.c_void_p]
        L.hb_bilateral.argtypes = [C.POINTER(A.hb_bilateral_desc), C.c_void_p]
        L.hb_point_op.argtypes = [C.POINTER(A.hb_point_desc), C.c_void_p]
        L.hb_reduce.argtypes = [C.POINTER(A.hb_view), C.c_int, C.c_void_p, C.c_void_p]
        L.hb_reduce_async.argtypes = [C.POINTER(A.hb_view), C.c_int, C.c_void_p, C.c_void_p]
        L.hb_reduce_minmaxsum_f32.argtypes = [C.POINTER(A.hb_view), C.POINTER(C.c_float), C.c_void_p]
        L.hb_reduce_minmaxsum_f32_async.argtypes = [C.POINTER(A.hb_view), C.c_void_p, C.c_void_p]
        L.hb_binning.argtypes = [C.POINTER(A.hb_binning_desc), C.c_void_p, C.c_void_p]
        L.hb_binning_async.argtypes = [C.POINTER(A.hb_binning_desc), C.c_void_p, C.c_void_p]
        L.hb_harris.argtypes = [C.POINTER(A.hb_harris_desc), C.c_void_p]
        L.hb_pyr_down.argtypes = [C.POINTER(A.hb_pyr_down_desc), C.c_void_p]
        L.hb_pyr_up.argtypes = [C.POINTER(A.hb_pyr_up_desc), C.c_void_p]
        L.hb_pyr_dog.argtypes = [C.POINTER(A.hb_pyr_dog_desc), C.c_void_p]
        L.hb_pyr_traverse_coarse.argtypes = [C.POINTER(A.hb_pyr_coarse_desc), C.c_void_p]
        L.hb_image_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(A.hb_view)]
        L.hb_image_destroy.argtypes = [C.POINTER(A.hb_view)]
        L.hb_image_wrap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(A.hb_view)]
        L.hb_image_write.argtypes = [C.POINTER(A.hb_view), C.c_void_p, C.c_void_p]
        L.hb_image_read.argtypes = [C.POINTER(A.hb_view), C.c_void_p, C.c_void_p]
        L.hb_image_write_region_async.argtypes = [C.POINTER(A.hb_view), C.c_void_p, C.c_size_t, C.c_void_p]
        L.hb_image_read_region_async.argtypes = [C.POINTER(A.hb_view), C.c_void_p, C.c_size_t, C.c_void_p]
        L.hb_image_copy.argtypes = [C.POINTER(A.hb_view), C.POINTER(A.hb_view), C.c_void_p]
        L.hb_image_copy_region.argtypes = [C.POINTER(A.hb_view), C.POINTER(A.hb_view), C.c_void_p]
        L.hb_stream_synchronize.argtypes = [C.c_void_p]
        L.hb_debug_timestamp.argtypes = [C.c_void_p, C.c_void_p]
        L.hb_stream_create.argtypes = [C.POINTER(C.c_void_p)]
        L.hb_stream_destroy.argtypes = [C.c_void_p]
        L.hb_graph_begin.argtypes = [C.c_void_p]
        L.hb_graph_end.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.hb_graph_launch.argtypes = [C.c_void_p, C.c_void_p]
        L.hb_graph_destroy.argtypes = [C.c_void_p]
        L.hb_ipc_export.argtypes = [C.c_void_p, C.POINTER(A.hb_ipc_mem)]
        L.hb_ipc_open.argtypes = [C.POINTER(A.hb_ipc_mem), C.POINTER(C.c_void_p)]
        L.hb_ipc_close.argtypes = [C.c_void_p]
        L.hb_halo_ctrl_create.argtypes = [C.POINTER(C.c_void_p)]
        L.hb_halo_ctrl_destroy.argtypes = [C.c_void_p]
        L.hb_halo_status.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hb_halo_exchange.argtypes = [C.POINTER(A.hb_halo_desc), C.c_void_p]
        L.hb_halo_exchange_batch.argtypes = [C.POINTER(C.POINTER(A.hb_halo_desc)), C.c_int, C.c_void_p]
        L.hb_allgather_rows.argtypes = [C.POINTER(A.hb_gather_desc), C.c_void_p]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise HbError(rc, what, (lib().hb_last_error() or b"").decode(errors="replace"))


def init(device=0):
    _check(lib().hb_init(int(device)), "hb_init")


def set_timing(on):
    lib().hb_set_timing(1 if on else 0)


def last_kernel_ms():
    return float(lib().hb_last_kernel_ms())


def timestamp(marks, index, stream=None):
    """diagnostics: write the GPU nanosecond timer into marks[index] (a CUDA int64 tensor) in stream order"""
    _check(lib().hb_debug_timestamp(C.c_void_p(marks.data_ptr() + 8 * index), stream_ptr(stream)), "hb_debug_timestamp")


def launch_count():
    return int(lib().hb_launch_count())


# ----------------------------------------------------------------------------- torch plumbing
_TORCH_DTYPES = None


def _torch_dtype_map():
    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        import torch
        _TORCH_DTYPES = {torch.uint8: A.U8, torch.int8: A.S8, torch.int16: A.S16, torch.int32: A.S32,
                         torch.float32: A.F32}
    return _TORCH_DTYPES


def torch_dtype(dt):
    import torch
    return {A.U8: torch.uint8, A.S8: torch.int8, A.S16: torch.int16, A.S32: torch.int32, A.F32: torch.float32}[dt]


def view(t, roi=None, ghost=(0, 0)):
    """hb_view over a 2-D CUDA tensor (row stride may exceed the width; unit pixel stride), or over a uint8 tensor of
    shape [H, W, 4] (uchar4 pixels, HB_U8X4; width / stride / roi count pixels)."""
    import torch
    if t.dim() == 3:
        x4 = {torch.uint8: A.U8X4, torch.int8: A.S8X4, torch.int16: A.S16X4, torch.int32: A.S32X4, torch.float32: A.F32X4}
        assert t.is_cuda and t.dtype in x4 and t.shape[2] == 4 and t.stride(2) == 1 and t.stride(1) == 4 and t.stride(0) % 4 == 0, \
            "4-channel images (uchar4, char4, short4, int4, float4) are CUDA tensors of shape [H, W, 4] with interleaved channels"
        return A.make_view(t.data_ptr(), x4[t.dtype], t.shape[1], t.shape[0], t.stride(0) // 4, roi, ghost)
    assert t.is_cuda and t.dim() == 2 and t.stride(1) == 1, "need a 2-D CUDA tensor with contiguous rows"
    return A.make_view(t.data_ptr(), _torch_dtype_map()[t.dtype], t.shape[1], t.shape[0], t.stride(0), roi, ghost)


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def empty_image(dtype, width, height, device="cuda", align_bytes=256):
    """Allocate an image in HBM whose rows start on `align_bytes` boundaries (returns the w x h view)."""
    import torch
    es = A.DTYPE_SIZE[dtype]
    per = max(1, align_bytes // es)
    stride = (width + per - 1) // per * per
    buf = torch.empty((height, stride), dtype=torch_dtype(dtype), device=device)
    return buf[:, :width]


class Graph:
    """A captured pipeline (hb_graph_begin / hb_graph_end): `with hb.Graph(stream) as g: ...operators on stream...`,
    then g.launch() replays every kernel of the block with one launch (the reference's -use-graph mode)."""

    def __init__(self, stream):
        self.stream, self.handle = stream, C.c_void_p()

    def __enter__(self):
        _check(lib().hb_graph_begin(stream_ptr(self.stream)), "hb_graph_begin")
        return self

    def __exit__(self, et, ev, tb):
        rc = lib().hb_graph_end(stream_ptr(self.stream), C.byref(self.handle))
        if et is None:
            _check(rc, "hb_graph_end")
        return False

    def launch(self, stream=None):
        _check(lib().hb_graph_launch(self.handle, stream_ptr(self.stream if stream is None else stream)), "hb_graph_launch")

    def destroy(self):
        if self.handle:
            lib().hb_graph_destroy(self.handle)
            self.handle = C.c_void_p()


class _DevArray:
    """__cuda_array_interface__ carrier so torch can wrap memory the library allocated"""

    def __init__(self, ptr, shape, strides, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "strides": strides, "typestr": typestr, "data": (ptr, False), "version": 2}


def alloc_image(dtype, width, height, device="cuda", align_bytes=256):
    """An image allocated by the LIBRARY (hb_image_create = one cudaMalloc), wrapped as a torch tensor of shape
    [height, stride].  Unlike a torch allocation its base pointer is a whole CUDA allocation, which is what CUDA IPC
    (hb_ipc_export, the peer-to-peer halo exchange) needs.  Returns the full-stride tensor; the memory is released
    when the returned tensor's `hb_owner` is destroyed explicitly (kept alive for the life of the process otherwise)."""
    import torch
    v = A.hb_view()
    with torch.cuda.device(device):
        _check(lib().hb_image_create(dtype, width, height, align_bytes, C.byref(v)), "hb_image_create")
    es = A.DTYPE_SIZE[dtype]
    typestr = {A.U8: "|u1", A.S8: "|i1", A.S16: "<i2", A.S32: "<i4", A.F32: "<f4"}[dtype]
    t = torch.as_tensor(_DevArray(v.data, (height, v.stride), (v.stride * es, es), typestr), device=device)
    t.hb_owner = v
    return t


# ----------------------------------------------------------------------------- operators
def local_op(spec, src, dst=None, roi_in=None, roi_out=None, ghost=(0, 0), stream=None):
    """Run a local operator (specs.LocalSpec) on CUDA tensors through hb_local_op."""
    import torch
    if dst is None:
        dst = torch.zeros(src.shape, dtype=torch_dtype(spec.out_dtype), device=src.device)   # [H, W, 4] for 4-channel pixels
    d = A.hb_local_desc()
    spec.fill(d)
    d.in_ = view(src, roi_in, ghost)
    d.out = view(dst, roi_out)
    _check(lib().hb_local_op(C.byref(d), stream_ptr(stream)), "hb_local_op")
    return dst


def bilateral(src, size, coef, sigma_r, boundary=A.CLAMP, const=0.0, dst=None, stream=None):
    import numpy as np
    import torch
    if dst is None:
        dst = torch.zeros_like(src)
    c = np.ascontiguousarray(coef, dtype=np.float32)
    d = A.hb_bilateral_desc()
    d.in_, d.out = view(src), view(dst)
    d.size, d.coef_f32, d.sigma_r = size, c.ctypes.data_as(C.POINTER(C.c_float)), int(sigma_r)
    d.boundary, d.boundary_const = boundary, float(const)
    _check(lib().hb_bilateral(C.byref(d), stream_ptr(stream)), "hb_bilateral")
    return dst


def point_op(op, inputs, out_dtype=None, out_shape=None, interp=None, p=(0.0, 0.0), dst=None, stream=None):
    import torch
    if dst is None:
        shape = tuple(out_shape) if out_shape is not None else tuple(inputs[0].shape)
        dst = torch.zeros(shape, dtype=torch_dtype(out_dtype if out_dtype is not None else A.U8), device=inputs[0].device)
    d = A.hb_point_desc()
    d.n_in = len(inputs)
    for i, t in enumerate(inputs):
        d.in_[i] = view(t)
        d.interp[i] = interp[i] if interp else A.INTERP_NO
    d.out = view(dst)
    d.op = op
    d.p[0], d.p[1] = float(p[0]), float(p[1])
    _check(lib().hb_point_op(C.byref(d), stream_ptr(stream)), "hb_point_op")
    return dst


def reduce_minmaxsum(src, roi=None, stream=None):
    """Fused one-pass global reduction of an f32 image -> (min, max, sum) as Python floats (blocking)."""
    v = view(src, roi)
    out = (C.c_float * 3)()
    _check(lib().hb_reduce_minmaxsum_f32(C.byref(v), out, stream_ptr(stream)), "hb_reduce_minmaxsum_f32")
    return float(out[0]), float(out[1]), float(out[2])


def reduce_minmaxsum_async(src, partials, roi=None, stream=None):
    """Leave {float min, float max, double sum} in `partials` (a 16-byte CUDA buffer) without syncing."""
    v = view(src, roi)
    _check(lib().hb_reduce_minmaxsum_f32_async(C.byref(v), C.c_void_p(partials.data_ptr()), stream_ptr(stream)),
           "hb_reduce_minmaxsum_f32_async")


def reduce(src, mode, roi=None, stream=None):
    import numpy as np
    v = view(src, roi)
    res = np.zeros(1, dtype=A.DTYPE_NUMPY[v.dtype])
    _check(lib().hb_reduce(C.byref(v), mode, res.ctypes.data_as(C.c_void_p), stream_ptr(stream)), "hb_reduce")
    return res[0]


def reduce_async(src, mode, out, roi=None, stream=None):
    """hb_reduce_async: integer images (every mode) / float PROD; the scalar is left in `out` (a CUDA tensor of >= 8
    bytes: int32 accumulator for integer pixels, float64 for float PROD).  Capturable."""
    assert out.is_cuda and out.numel() * out.element_size() >= 8 and out.data_ptr() % 8 == 0
    v = view(src, roi)
    _check(lib().hb_reduce_async(C.byref(v), mode, C.c_void_p(out.data_ptr()), stream_ptr(stream)), "hb_reduce_async")
    return out


def _binning_desc(src, num_bins, index_kind, value_kind, p0, roi):
    d = A.hb_binning_desc()
    d.in_ = view(src, roi)
    d.num_bins, d.index_kind, d.value_kind, d.p0 = int(num_bins), index_kind, value_kind, float(p0)
    return d


def binning(src, num_bins, index_kind=A.BIN_INDEX_SCALE, value_kind=A.BIN_VALUE_ONE, p0=255.0, roi=None, stream=None):
    """Kernel::binned_data(num_bins) for `bin(INDEX(pixel)) = VALUE(pixel)`, reduce = + (hb_binning, blocking)
    -> numpy uint32[num_bins].  Defaults are the Histogram sample: bin(pixel/255.0f*num_bins) = 1."""
    import numpy as np
    d = _binning_desc(src, num_bins, index_kind, value_kind, p0, roi)
    out = np.zeros(num_bins, dtype=np.uint32)
    _check(lib().hb_binning(C.byref(d), out.ctypes.data_as(C.c_void_p), stream_ptr(stream)), "hb_binning")
    return out


def binning_async(src, bins, index_kind=A.BIN_INDEX_SCALE, value_kind=A.BIN_VALUE_ONE, p0=255.0, roi=None, stream=None):
    """Same, bins left in the CUDA int32/uint32 tensor `bins` (zeroed by the call), no synchronisation."""
    d = _binning_desc(src, bins.numel(), index_kind, value_kind, p0, roi)
    _check(lib().hb_binning_async(C.byref(d), C.c_void_p(bins.data_ptr()), stream_ptr(stream)), "hb_binning_async")
    return bins


def harris(src, k=masks.HARRIS_K, threshold=masks.HARRIS_THRESHOLD, dst=None, roi=None, ghost=(0, 0), stream=None):
    """Fused Harris corner detector (uchar -> uchar), one kernel (hb_harris).  `roi` / `ghost` select the
    owned rows of a strip buffer whose ghost rows hold real neighbour data (row-strip sharding)."""
    import torch
    if dst is None:
        dst = torch.zeros_like(src)
    d = A.hb_harris_desc()
    d.in_, d.out = view(src, roi, ghost), view(dst, roi)
    d.k, d.threshold = float(k), float(threshold)
    _check(lib().hb_harris(C.byref(d), stream_ptr(stream)), "hb_harris")
    return dst


def harris_unfused(src, k=masks.HARRIS_K, threshold=masks.HARRIS_THRESHOLD, stream=None):
    """The sample's nine-kernel pipeline (Harris_Corner/src/main.cpp:230-305) through the generic
    local / point operators -- the shape Hipacc's rewritten host code has; used to cross-check the
    fused kernel and to measure what fusion buys."""
    dx = local_op(specs.harris_deriv(masks.HARRIS_DX), src, stream=stream)
    dy = local_op(specs.harris_deriv(masks.HARRIS_DY), src, stream=stream)
    sx = point_op(A.POINT_SQUARE, [dx], A.S16, stream=stream)
    sy = point_op(A.POINT_SQUARE, [dy], A.S16, stream=stream)
    sxy = point_op(A.POINT_MUL, [dx, dy], A.S16, stream=stream)
    gx = local_op(specs.harris_gauss(masks.HARRIS_GAUSS3), sx, stream=stream)
    gy = local_op(specs.harris_gauss(masks.HARRIS_GAUSS3), sy, stream=stream)
    gxy = local_op(specs.harris_gauss(masks.HARRIS_GAUSS3), sxy, stream=stream)
    out = point_op(A.POINT_HARRIS, [gx, gy, gxy], A.U8, p=(k, threshold), stream=stream)
    return out, gx, gy, gxy


class Pyramid:
    """Device-resident image pyramid: level l is (w >> l) x (h >> l) (hipaccCreatePyramid,
    runtime/hipacc_cu.tpp:482-497); level 0 aliases the user image like the emitted code."""

    def __init__(self, base, depth):
        import torch
        self.depth = depth
        sizes = specs.pyramid_sizes(base.shape[1], base.shape[0], depth)
        self.levels = [base] + [empty_image(A.F32, w, h, device=base.device).zero_() for (w, h) in sizes[1:]]
        assert base.dtype == torch.float32


def _v(t):
    """hb_view of a tensor, or of a (tensor, roi, ghost) triple (row-strip buffers)"""
    return view(*t) if isinstance(t, tuple) else view(t)


def pyr_down(fine, coarse, mask, lap_fine=None, tmp=None, stream=None):
    """fine / coarse / lap_fine: 2-D CUDA tensors, or (tensor, roi, ghost) triples for row strips."""
    import numpy as np
    m = np.ascontiguousarray(mask, dtype=np.float32)
    d = A.hb_pyr_down_desc()
    d.fine, d.coarse = _v(fine), _v(coarse)
    if tmp is not None:
        d.tmp = _v(tmp)
    if lap_fine is not None:
        d.lap_fine = _v(lap_fine)
    d.size, d.coef_f32 = m.shape[0], m.ctypes.data_as(C.POINTER(C.c_float))
    _check(lib().hb_pyr_down(C.byref(d), stream_ptr(stream)), "hb_pyr_down")


def pyr_dog(fine, coarse, lap_fine, stream=None):
    """lap_fine = fine - LF(coarse)  (hb_pyr_dog); operands as in pyr_down"""
    d = A.hb_pyr_dog_desc()
    d.fine, d.coarse, d.lap_fine = _v(fine), _v(coarse), _v(lap_fine)
    _check(lib().hb_pyr_dog(C.byref(d), stream_ptr(stream)), "hb_pyr_dog")


def pyr_up(coarse_gaus, coarse_lap, fine_gaus, fine_lap, stream=None):
    d = A.hb_pyr_up_desc()
    d.coarse_gaus, d.coarse_lap, d.fine_gaus, d.fine_lap = _v(coarse_gaus), _v(coarse_lap), _v(fine_gaus), _v(fine_lap)
    _check(lib().hb_pyr_up(C.byref(d), stream_ptr(stream)), "hb_pyr_up")


COARSE_PIXELS = 1 << 20   # levels of at most this many pixels form the "coarse end" (one cooperative launch)


def pyr_traverse_coarse(gaus_levels, lap_levels, mask, stream=None):
    """hb_pyr_traverse_coarse: way down and way up through a small pyramid (2..8 levels) in ONE launch.  Returns False
    when the levels are not aligned exact halvings (the caller then walks them level by level)."""
    import numpy as np
    m = np.ascontiguousarray(mask, dtype=np.float32)
    d = A.hb_pyr_coarse_desc()
    d.levels = len(gaus_levels)
    for i, (g, l) in enumerate(zip(gaus_levels, lap_levels)):
        d.gaus[i], d.lap[i] = _v(g), _v(l)
    d.size, d.coef_f32 = m.shape[0], m.ctypes.data_as(C.POINTER(C.c_float))
    rc = lib().hb_pyr_traverse_coarse(C.byref(d), stream_ptr(stream))
    if rc == A.HB_ERR_UNSUPPORTED:
        return False
    _check(rc, "hb_pyr_traverse_coarse")
    return True


def _halving(levels, l):
    return levels[l - 1].shape[0] == 2 * levels[l].shape[0] and levels[l - 1].shape[1] == 2 * levels[l].shape[1]


def coarse_start(levels, first=1):
    """index c >= first of the finest level from which the rest of the pyramid is small enough for the one-launch
    coarse traversal (exact halvings only, at least two levels), or None"""
    depth = len(levels)
    for c in range(first, depth - 1):
        if levels[c].shape[0] * levels[c].shape[1] <= COARSE_PIXELS and depth - c <= 8 and all(_halving(levels, l) for l in range(c + 1, depth)):
            return c
    return None


def pyramid_traverse(pgaus, plap, mask, ptmp=None, stream=None, fuse_coarse=False):
    """The traversal of Gaussian_Laplacian_Pyramid/src/main.cpp:199-248: way down builds the Gaussian and
    Laplacian pyramids, way up restores / blends.  Host recursion only (dsl/pyramid.hpp:182-208).  fuse_coarse: run the coarse
    end (levels of <= COARSE_PIXELS pixels) as one cooperative launch (hb_pyr_traverse_coarse).  Off by default: in a
    CUDA-graph replay the six per-level kernels of the 1024^2 .. 128^2 tail take 43 us, the one-launch kernel 57 us
    (tools/coarse_probe.py) -- every phase is a latency chain either way and a grid-wide barrier costs about what a
    graph node does; it pays only for hosts that launch every kernel themselves."""
    depth = pgaus.depth
    c = coarse_start(pgaus.levels) if (fuse_coarse and ptmp is None) else None
    last_down = depth - 1 if c is None else c
    for l in range(1, last_down + 1):
        pyr_down(pgaus.levels[l - 1], pgaus.levels[l], mask, lap_fine=plap.levels[l - 1],
                 tmp=None if ptmp is None else ptmp.levels[l - 1], stream=stream)
    first_up = depth - 2
    if c is not None:
        if pyr_traverse_coarse(pgaus.levels[c:], plap.levels[c:], mask, stream=stream):
            first_up = c - 1
        else:
            for l in range(c + 1, depth):
                pyr_down(pgaus.levels[l - 1], pgaus.levels[l], mask, lap_fine=plap.levels[l - 1], stream=stream)
    for l in range(first_up, -1, -1):
        pyr_up(pgaus.levels[l + 1], plap.levels[l + 1], pgaus.levels[l], plap.levels[l], stream=stream)
