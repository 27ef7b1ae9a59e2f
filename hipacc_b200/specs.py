"""Host-side operator specifications: the Python mirror of the DSL-level description of an
operator (what a Kernel subclass' kernel() body says), lowered onto the C-ABI descriptors
of include/hipacc_b200.h.  No compute happens here.

The named constructors follow the reference samples (paths relative to samples-public/):
    gaussian_blur   1_Local_Operators/Gaussian_Blur/src/main.cpp:48-66
    sobel / laplace 3_Preprocessing/Sobel/src/main.cpp:55-73, 1_Local_Operators/Laplace/src/main.cpp:50-72
    dilate / erode  1_Local_Operators/Dilate/src/main.cpp:48-64
    box_blur        1_Local_Operators/Box_Blur/src/main.cpp:49-66
    harris_*        3_Preprocessing/Harris_Corner/src/main.cpp:55-164
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _abi as A


@dataclass
class LocalSpec:
    size_x: int
    size_y: int
    kind: int = A.CONVOLVE
    reduce_mode: int = A.SUM
    tap: int = A.TAP_MUL
    acc_dtype: int = A.F32
    coef: Optional[np.ndarray] = None        # float32 or int32, shape (size_y, size_x)
    domain: Optional[np.ndarray] = None      # uint8 0/1
    boundary: int = A.CLAMP
    boundary_const: float = 0.0
    epilogue: int = A.EPI_CAST
    epi_p: Sequence[float] = (0.0, 0.0, 0.0)
    out_dtype: int = A.F32
    _keep: list = field(default_factory=list, repr=False)

    def fill(self, d: "A.hb_local_desc"):
        """Write everything but the two views into a descriptor (host arrays are kept alive)."""
        d.kind, d.reduce_mode, d.tap, d.acc_dtype = self.kind, self.reduce_mode, self.tap, self.acc_dtype
        d.size_x, d.size_y = self.size_x, self.size_y
        d.coef_f32 = None
        d.coef_s32 = None
        d.domain = None
        if self.coef is not None:
            c = np.ascontiguousarray(self.coef)
            assert c.shape == (self.size_y, self.size_x), c.shape
            if c.dtype == np.float32:
                d.coef_f32 = c.ctypes.data_as(C.POINTER(C.c_float))
            else:
                c = np.ascontiguousarray(c.astype(np.int32))
                d.coef_s32 = c.ctypes.data_as(C.POINTER(C.c_int))
            self._keep.append(c)
        if self.domain is not None:
            dm = np.ascontiguousarray(self.domain.astype(np.uint8))
            d.domain = dm.ctypes.data_as(C.POINTER(C.c_ubyte))
            self._keep.append(dm)
        d.boundary, d.boundary_const = self.boundary, float(self.boundary_const)
        d.epilogue = self.epilogue
        for i in range(3):
            d.epi_p[i] = float(self.epi_p[i]) if i < len(self.epi_p) else 0.0
        return d

    @property
    def radius(self):
        return self.size_x // 2, self.size_y // 2


def gaussian_blur(mask, boundary=A.CLAMP, out_dtype=A.U8):
    """uchar -> uchar: (uchar)(convolve(mask, SUM, mask()*input(mask)) + 0.5f)"""
    m = np.asarray(mask, dtype=np.float32)
    return LocalSpec(m.shape[1], m.shape[0], A.CONVOLVE, A.SUM, A.TAP_MUL, A.F32, m, None, boundary, 0.0,
                     A.EPI_ADD_CAST, (0.5, 0, 0), out_dtype)


def convolve_f32(mask, boundary=A.CLAMP, mode=A.SUM, const=0.0):
    """float -> float: convolve(mask, mode, mask()*input(mask)) -- pyramid Gaussian, generic float"""
    m = np.asarray(mask, dtype=np.float32)
    return LocalSpec(m.shape[1], m.shape[0], A.CONVOLVE, mode, A.TAP_MUL, A.F32, m, None, boundary, const,
                     A.EPI_CAST, (0, 0, 0), A.F32)


def domain_reduce_f32(mask, boundary=A.MIRROR, mode=A.SUM, const=0.0):
    """float -> float: reduce(dom, mode, mask(dom)*in(dom)) over the non-zero taps (config C2)"""
    m = np.asarray(mask, dtype=np.float32)
    return LocalSpec(m.shape[1], m.shape[0], A.REDUCE_DOMAIN, mode, A.TAP_MUL, A.F32, m, None, boundary, const,
                     A.EPI_CAST, (0, 0, 0), A.F32)


def sobel_u8(mask, boundary=A.CLAMP, out_dtype=A.S32):
    """uchar -> int: (data_t)reduce(dom, SUM, mask(dom)*input(dom)) with an int mask"""
    m = np.asarray(mask, dtype=np.int32)
    return LocalSpec(m.shape[1], m.shape[0], A.REDUCE_DOMAIN, A.SUM, A.TAP_MUL, A.S32, m, None, boundary, 0.0,
                     A.EPI_CAST, (0, 0, 0), out_dtype)


def laplace_u8(mask, boundary=A.CLAMP, add=128):
    """uchar -> uchar: sum += 128; min(sum,255); max(sum,0).
    uchar4 (Laplace_RGBA): the reference's `int4 += int` takes its left operand BY VALUE (dsl/types.hpp:146-148,
    runtime/hipacc_types.hpp:173-175), so `sum += 128` has no effect in the DSL, in the sample's own checker and in
    the emitted CUDA / C++ code alike -- the reference's RGBA result is `add=0`."""
    m = np.asarray(mask, dtype=np.int32)
    return LocalSpec(m.shape[1], m.shape[0], A.REDUCE_DOMAIN, A.SUM, A.TAP_MUL, A.S32, m, None, boundary, 0.0,
                     A.EPI_ADD_CLAMP_CAST, (add, 0, 255), A.U8)


def minmax_u8(size_x, size_y, is_max, boundary=A.CLAMP):
    """Dilate (MAX) / Erode (MIN) over a full Domain"""
    return LocalSpec(size_x, size_y, A.REDUCE_DOMAIN, A.MAX if is_max else A.MIN, A.TAP_IN, A.S32, None,
                     np.ones((size_y, size_x), np.uint8), boundary, 0.0, A.EPI_CAST, (0, 0, 0), A.U8)


def box_blur_u8(size_x, size_y, boundary=A.CLAMP):
    """reduce(dom, SUM, in(dom)) / (float)(size_x*size_y)"""
    return LocalSpec(size_x, size_y, A.REDUCE_DOMAIN, A.SUM, A.TAP_IN, A.S32, None,
                     np.ones((size_y, size_x), np.uint8), boundary, 0.0, A.EPI_DIVF_CAST,
                     (size_x * size_y, 0, 0), A.U8)


def harris_deriv(mask):
    """Harris Sobel: uchar -> short, short accumulate, /6"""
    m = np.asarray(mask, dtype=np.int32)
    return LocalSpec(3, 3, A.REDUCE_DOMAIN, A.SUM, A.TAP_MUL, A.S16, m, None, A.CLAMP, 0.0,
                     A.EPI_DIVI_CAST, (6, 0, 0), A.S16)


def harris_gauss(mask, norm=16):
    """Harris Gaussian: short -> short, int accumulate over all taps, /norm"""
    m = np.asarray(mask, dtype=np.int32)
    return LocalSpec(m.shape[1], m.shape[0], A.CONVOLVE, A.SUM, A.TAP_MUL, A.S32, m, None, A.CLAMP, 0.0,
                     A.EPI_DIVI_CAST, (norm, 0, 0), A.S16)


def pyramid_sizes(w, h, depth):
    """Level extents of hipaccCreatePyramid (runtime/hipacc_cu.tpp:482-497, dsl/pyramid.hpp:103-114)."""
    out = [(w, h)]
    for _ in range(1, depth):
        w, h = w // 2, h // 2
        assert w * h > 0, "Pyramid stages too deep for image size"
        out.append((w, h))
    return out
