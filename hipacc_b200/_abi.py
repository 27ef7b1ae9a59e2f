"""ctypes mirror of include/hipacc_b200.h (the C ABI of the B200 operator path).

Only layout definitions live here -- no compute, no library loading -- so that the
product wrapper (hipacc_b200/__init__.py, device pointers) and the test-only CPU oracle
wrapper (oracle/oracle.py, host pointers) build byte-identical descriptors.
"""
import ctypes as C

# hb_status
HB_OK, HB_ERR_INVALID, HB_ERR_UNSUPPORTED, HB_ERR_CUDA, HB_ERR_NO_DEVICE = 0, -1, -2, -3, -4
# hb_dtype
U8, S8, U16, S16, S32, U32, F32, U8X4, S8X4, U16X4, S16X4, S32X4, U32X4, F32X4 = range(14)
DTYPE_SIZE = {U8: 1, S8: 1, U16: 2, S16: 2, S32: 4, U32: 4, F32: 4, U8X4: 4, S8X4: 4, U16X4: 8, S16X4: 8, S32X4: 16, U32X4: 16, F32X4: 16}
DTYPE_NUMPY = {U8: "uint8", S8: "int8", U16: "uint16", S16: "int16", S32: "int32", U32: "uint32", F32: "float32"}
NUMPY_DTYPE = {v: k for k, v in DTYPE_NUMPY.items()}
# hb_boundary == hipacc::Boundary (dsl/image.hpp:46-52)
UNDEFINED, CLAMP, REPEAT, MIRROR, CONSTANT = range(5)
BOUNDARY_NAMES = {UNDEFINED: "UNDEFINED", CLAMP: "CLAMP", REPEAT: "REPEAT", MIRROR: "MIRROR", CONSTANT: "CONSTANT"}
# hb_interp == hipacc::Interpolate (dsl/image.hpp:54-61)
INTERP_NO, INTERP_NN, INTERP_LF, INTERP_B5, INTERP_CF, INTERP_L3 = range(6)
# hb_reduce_mode == hipacc::Reduce (dsl/kernel.hpp:48-54)
SUM, MIN, MAX, PROD = range(4)
# hb_local_kind / hb_tap / hb_epilogue
CONVOLVE, REDUCE_DOMAIN = 0, 1
TAP_MUL, TAP_IN = 0, 1
EPI_CAST, EPI_ADD_CAST, EPI_ADD_CLAMP_CAST, EPI_DIVI_CAST, EPI_DIVF_CAST = range(5)
# hb_point_kind
(POINT_COPY, POINT_SQUARE, POINT_MUL, POINT_SUB, POINT_ADD, POINT_BLEND,
 POINT_SOBEL_COMBINE, POINT_HARRIS) = range(8)
# hb_bin_index / hb_bin_value
BIN_INDEX_SCALE, BIN_INDEX_PIXEL = 0, 1
BIN_VALUE_ONE, BIN_VALUE_PIXEL = 0, 1


class hb_view(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("dtype", C.c_int),
        ("img_width", C.c_int), ("img_height", C.c_int),
        ("stride", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("offset_x", C.c_int), ("offset_y", C.c_int),
        ("ghost_top", C.c_int), ("ghost_bottom", C.c_int),
    ]


class hb_local_desc(C.Structure):
    _fields_ = [
        ("in_", hb_view), ("out", hb_view),
        ("kind", C.c_int), ("reduce_mode", C.c_int), ("tap", C.c_int), ("acc_dtype", C.c_int),
        ("size_x", C.c_int), ("size_y", C.c_int),
        ("coef_f32", C.POINTER(C.c_float)),
        ("coef_s32", C.POINTER(C.c_int)),
        ("domain", C.POINTER(C.c_ubyte)),
        ("boundary", C.c_int),
        ("boundary_const", C.c_double),
        ("epilogue", C.c_int),
        ("epi_p", C.c_double * 3),
    ]


class hb_bilateral_desc(C.Structure):
    _fields_ = [
        ("in_", hb_view), ("out", hb_view),
        ("size", C.c_int),
        ("coef_f32", C.POINTER(C.c_float)),
        ("sigma_r", C.c_int),
        ("boundary", C.c_int),
        ("boundary_const", C.c_double),
    ]


class hb_point_desc(C.Structure):
    _fields_ = [
        ("in_", hb_view * 3),
        ("interp", C.c_int * 3),
        ("n_in", C.c_int),
        ("out", hb_view),
        ("op", C.c_int),
        ("p", C.c_double * 2),
    ]


class hb_binning_desc(C.Structure):
    _fields_ = [("in_", hb_view), ("num_bins", C.c_int), ("index_kind", C.c_int), ("value_kind", C.c_int), ("p0", C.c_double)]


class hb_harris_desc(C.Structure):
    _fields_ = [("in_", hb_view), ("out", hb_view), ("k", C.c_float), ("threshold", C.c_float)]


class hb_pyr_down_desc(C.Structure):
    _fields_ = [
        ("fine", hb_view), ("tmp", hb_view), ("coarse", hb_view), ("lap_fine", hb_view),
        ("size", C.c_int),
        ("coef_f32", C.POINTER(C.c_float)),
    ]


class hb_pyr_dog_desc(C.Structure):
    _fields_ = [("fine", hb_view), ("coarse", hb_view), ("lap_fine", hb_view)]


class hb_pyr_up_desc(C.Structure):
    _fields_ = [("coarse_gaus", hb_view), ("coarse_lap", hb_view), ("fine_gaus", hb_view), ("fine_lap", hb_view)]


class hb_pyr_coarse_desc(C.Structure):
    _fields_ = [("levels", C.c_int), ("gaus", hb_view * 8), ("lap", hb_view * 8), ("size", C.c_int), ("coef_f32", C.POINTER(C.c_float))]


class hb_ipc_mem(C.Structure):
    _fields_ = [("handle", C.c_ubyte * 64)]


class hb_halo_desc(C.Structure):
    _fields_ = [
        ("buf", C.c_void_p), ("pitch_bytes", C.c_size_t), ("row_bytes", C.c_size_t),
        ("ghost_top", C.c_int), ("rows", C.c_int), ("radius", C.c_int),
        ("ctrl", C.c_void_p),
        ("up_buf", C.c_void_p), ("up_ctrl", C.c_void_p), ("up_pitch_bytes", C.c_size_t),
        ("up_ghost_top", C.c_int), ("up_rows", C.c_int),
        ("down_buf", C.c_void_p), ("down_ctrl", C.c_void_p), ("down_pitch_bytes", C.c_size_t),
        ("down_ghost_top", C.c_int),
    ]


HB_MAX_PEERS = 15


class hb_gather_desc(C.Structure):
    _fields_ = [
        ("buf", C.c_void_p), ("pitch_bytes", C.c_size_t), ("row_bytes", C.c_size_t),
        ("row0", C.c_int), ("rows", C.c_int),
        ("ctrl", C.c_void_p),
        ("my_slot", C.c_int), ("n_peers", C.c_int),
        ("peer_buf", C.c_void_p * HB_MAX_PEERS), ("peer_ctrl", C.c_void_p * HB_MAX_PEERS),
        ("peer_slot", C.c_int * HB_MAX_PEERS),
    ]


def make_view(ptr, dtype, img_w, img_h, stride=None, roi=None, ghost=(0, 0)):
    """Build an hb_view.  roi = (w, h, ox, oy) or None for the whole image."""
    v = hb_view()
    v.data = ptr
    v.dtype = dtype
    v.img_width, v.img_height = img_w, img_h
    v.stride = img_w if stride is None else stride
    if roi is None:
        v.width, v.height, v.offset_x, v.offset_y = img_w, img_h, 0, 0
    else:
        v.width, v.height, v.offset_x, v.offset_y = roi
    v.ghost_top, v.ghost_bottom = ghost
    return v


# Every symbol include/hipacc_b200.h declares (tests check the built library exports all of them)
EXPORTS = [
    "hb_init", "hb_device_count", "hb_sm_count", "hb_set_log_callback", "hb_last_error",
    "hb_image_create", "hb_image_destroy", "hb_image_wrap", "hb_image_write", "hb_image_read",
    "hb_image_copy", "hb_image_copy_region", "hb_image_write_region_async", "hb_image_read_region_async", "hb_set_timing", "hb_last_kernel_ms", "hb_launch_count",
    "hb_stream_synchronize", "hb_debug_timestamp", "hb_stream_create", "hb_stream_destroy", "hb_graph_begin", "hb_graph_end", "hb_graph_launch", "hb_graph_destroy",
    "hb_local_op", "hb_bilateral", "hb_point_op",
    "hb_reduce", "hb_reduce_async", "hb_reduce_minmaxsum_f32", "hb_reduce_minmaxsum_f32_async",
    "hb_binning", "hb_binning_async",
    "hb_harris", "hb_pyr_down", "hb_pyr_up", "hb_pyr_dog", "hb_pyr_traverse_coarse",
    "hb_ipc_export", "hb_ipc_open", "hb_ipc_close", "hb_halo_ctrl_create", "hb_halo_ctrl_destroy", "hb_halo_status", "hb_halo_exchange", "hb_halo_exchange_batch", "hb_allgather_rows",
]
