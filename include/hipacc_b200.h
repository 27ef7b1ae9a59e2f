/*
 * hipacc_b200.h -- C ABI of the B200-native execution path for Hipacc operators.
 *
 * This is the drop-in boundary: what Hipacc's rewritten host code (today emitted by
 * lib/Rewrite/CreateHostStrings.cpp against runtime/hipacc_cu.hpp) would bind instead of the
 * header-only CUDA runtime plus one generated .cu file per Kernel instance.  Every entry
 * point cites the reference interface it replaces (paths relative to the Hipacc tree).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  Pixel buffers are DEVICE pointers (HBM);
 *     masks / domains / scalar results are HOST pointers.
 *   - every call returns an int status: 0 = ok, <0 = error (hb_status).  Like the
 *     reference's checkErr (runtime/hipacc_cu.hpp:69-75, hipacc_base.hpp:112-129) errors are
 *     also logged through a callback and execution continues; there is NO CPU fallback --
 *     an operator that cannot run on the device fails with HB_ERR_UNSUPPORTED.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream), the stream the
 *     reference takes from HipaccExecutionParameterCuda (runtime/hipacc_cu.hpp:234-245).
 *   - enum values equal the DSL's (dsl/image.hpp:46-61, dsl/kernel.hpp:48-54).
 */
#ifndef HIPACC_B200_H
#define HIPACC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_VERSION 100

typedef enum {
  HB_OK = 0,
  HB_ERR_INVALID = -1,      /* malformed descriptor */
  HB_ERR_UNSUPPORTED = -2,  /* no device kernel for this (type, size, op) combination */
  HB_ERR_CUDA = -3,         /* CUDA runtime / driver error (logged) */
  HB_ERR_NO_DEVICE = -4
} hb_status;

typedef enum {  /* pixel / accumulator types */
  HB_U8 = 0, HB_S8 = 1, HB_U16 = 2, HB_S16 = 3, HB_S32 = 4, HB_U32 = 5, HB_F32 = 6,
  /* vector pixel (dsl/types.hpp:56-516): uchar4 = 4 interleaved uchar channels (RGBA); width / stride / offsets count
   * PIXELS.  Local and point operators treat the channels independently, which is what the DSL's element-wise
   * float4 / int4 arithmetic and convert_uchar4() do (samples-public/1_Local_Operators, the _RGBA samples). */
  HB_U8X4 = 7,
  /* the other 4-channel pixel types of dsl/types.hpp (char4, ushort4, short4, int4, uint4, float4): images of these types
   * can be created, copied, read and written, are the intermediates of the RGBA pipelines (Sobel_RGBA: short4 / int4) and
   * are what kernels compiled from a kernel() body (include/hipacc_b200/hipacc.hpp under nvcc) operate on. */
  HB_S8X4 = 8, HB_U16X4 = 9, HB_S16X4 = 10, HB_S32X4 = 11, HB_U32X4 = 12, HB_F32X4 = 13
} hb_dtype;
#define HB_DTYPE_LAST HB_F32X4

typedef enum {  /* hipacc::Boundary, dsl/image.hpp:46-52 */
  HB_BOUNDARY_UNDEFINED = 0, HB_BOUNDARY_CLAMP = 1, HB_BOUNDARY_REPEAT = 2,
  HB_BOUNDARY_MIRROR = 3, HB_BOUNDARY_CONSTANT = 4
} hb_boundary;

typedef enum {  /* hipacc::Interpolate, dsl/image.hpp:54-61: nearest neighbour, bilinear, binomial 5 (4 x 4 taps), bicubic
                 * convolution (4 x 4), Lanczos 3 (6 x 6) */
  HB_INTERP_NO = 0, HB_INTERP_NN = 1, HB_INTERP_LF = 2, HB_INTERP_B5 = 3, HB_INTERP_CF = 4, HB_INTERP_L3 = 5
} hb_interp;

typedef enum {  /* hipacc::Reduce, dsl/kernel.hpp:48-54 */
  HB_REDUCE_SUM = 0, HB_REDUCE_MIN = 1, HB_REDUCE_MAX = 2, HB_REDUCE_PROD = 3
} hb_reduce_mode;

/*
 * A view = image + rectangular region: the reference's HipaccAccessor<T>{img,width,height,
 * offset_x,offset_y} (runtime/hipacc_cu.hpp:162-186), used for Accessor and IterationSpace
 * alike (lib/Rewrite/Rewrite.cpp:1161,1249).  `data` points at pixel (0,0) of the image;
 * `stride` is in pixels (runtime/hipacc_cu.tpp:51-68).
 *
 * ghost_top / ghost_bottom (extension for row-strip sharding, 0 = reference semantics):
 * rows of valid neighbour data that exist above / below the region inside the same
 * allocation.  Boundary handling treats the region as [offset_y - ghost_top,
 * offset_y + height + ghost_bottom) vertically, so ghost rows are read as real data and the
 * boundary mode is applied only at the global image edge.
 */
typedef struct {
  void *data;
  int dtype;                 /* hb_dtype */
  int img_width, img_height; /* allocation extent in pixels */
  int stride;                /* pixels per row */
  int width, height;         /* region size; 0,0 = whole image */
  int offset_x, offset_y;
  int ghost_top, ghost_bottom;
} hb_view;

/* ------------------------------------------------------------------ device / runtime layer */

/* hipaccInitCUDA (runtime/hipacc_cu_standalone.hpp:113-163): select the device this process
 * drives (one process per GPU; the reference always picks device 0). */
int hb_init(int device);
int hb_device_count(void);
/* number of SMs of the selected device */
int hb_sm_count(void);

typedef void (*hb_log_fn)(int level /*0 info,1 warn,2 error*/, const char *msg);
void hb_set_log_callback(hb_log_fn fn);
const char *hb_last_error(void);

/* createMemory / hipaccCreateMemory<T> (runtime/hipacc_cu.tpp:43-68): allocate a w x h image;
 * `alignment` in bytes (0 = the library default of 256 B so every row is TMA-addressable;
 * the reference's default is stride == width).  Returns the device pointer and the stride
 * (in pixels) through the view, with region = whole image. */
int hb_image_create(int dtype, int width, int height, int alignment, hb_view *out);
/* ~HipaccImageCudaRaw (runtime/hipacc_cu.hpp:141-144) */
int hb_image_destroy(hb_view *img);
/* hipaccMapMemory (runtime/hipacc_cu.hpp:286): wrap an already resident device buffer
 * (e.g. a torch tensor's data_ptr()) without copying. */
int hb_image_wrap(void *device_ptr, int dtype, int width, int height, int stride, hb_view *out);
/* hipaccWriteMemory / hipaccReadMemory (runtime/hipacc_cu.tpp:85-116,166-193): blocking 2-D
 * copies between a dense host array (stride == width) and the image. */
int hb_image_write(const hb_view *img, const void *host, void *stream);
int hb_image_read(const hb_view *img, void *host, void *stream);
/* Asynchronous region transfers (an addition: the reference's hipaccWriteMemory / hipaccReadMemory block).  Copy the
 * view's REGION (width x height at its offsets) from / to a host array whose rows are `host_pitch_bytes` apart; `host`
 * points at the region's first pixel.  Enqueued on `stream` without synchronising: with pinned host memory the
 * transfers of one row strip overlap the operators of another, and host->device overlaps device->host. */
int hb_image_write_region_async(const hb_view *region, const void *host, size_t host_pitch_bytes, void *stream);
int hb_image_read_region_async(const hb_view *region, void *host, size_t host_pitch_bytes, void *stream);
/* hipaccCopyMemory / hipaccCopyMemoryRegion (runtime/hipacc_cu_standalone.hpp:166-216) */
int hb_image_copy(const hb_view *src, const hb_view *dst, void *stream);
int hb_image_copy_region(const hb_view *src, const hb_view *dst, void *stream);

/* print_timing / hipacc_last_kernel_timing (runtime/hipacc_cu_standalone.hpp:297-326,
 * hipacc_base.hpp:64-66): when enabled every operator call is bracketed by CUDA events on
 * its stream and synchronised; the elapsed milliseconds are returned by hb_last_kernel_ms. */
void hb_set_timing(int enabled);
float hb_last_kernel_ms(void);
/* launches issued by this library since process start (bench.py's gpu_launches) */
long long hb_launch_count(void);
int hb_stream_synchronize(void *stream);
/* diagnostics: a one-thread kernel that writes the GPU's nanosecond %globaltimer into *slot_device (8 bytes) in stream
 * order -- a timeline of the kernels of a captured pipeline, rank by rank (bench.py HB_BENCH_PHASES) */
int hb_debug_timestamp(void *slot_device, void *stream);
/* The stream a HipaccExecutionParameterCuda carries (runtime/hipacc_cu.hpp:234-245) is the user's cudaStream_t; host
 * code that does not include the CUDA headers creates one here (non-blocking with respect to the default stream). */
int hb_stream_create(void **stream);
int hb_stream_destroy(void *stream);

/* ------------------------------------------------------------------ CUDA graphs */
/*
 * The reference's -use-graph mode builds a cudaGraph node by node (hipaccLaunchKernelCudaGraph,
 * hipaccWriteMemoryCudaGraph, runtime/hipacc_cu_standalone.hpp:331-356, hipacc_cu.hpp:260-285).  Here
 * every non-blocking entry point of this library (operators, hb_halo_exchange, hb_*_async, the
 * region copies) is capturable: bracket a pipeline with hb_graph_begin / hb_graph_end on a
 * non-default stream and replay it with hb_graph_launch -- one launch for a multi-kernel program
 * (the unfused Harris pipeline: 9 kernels; a pyramid traversal: 14), no per-kernel host cost.
 * Blocking calls (hb_reduce, hb_binning, hb_image_write / read) must stay outside a capture.
 */
typedef struct hb_graph hb_graph;
int hb_graph_begin(void *stream);
int hb_graph_end(void *stream, hb_graph **out);
int hb_graph_launch(hb_graph *g, void *stream);
int hb_graph_destroy(hb_graph *g);

/* ------------------------------------------------------------------ local operators */
/*
 * Replaces one generated local-operator kernel + its launch (kernel text
 * lib/Rewrite/Rewrite.cpp:2726-2883, body lib/AST/ASTTranslate.cpp:510-1197, launch
 * runtime/hipacc_cu_standalone.hpp:66-110,277-329) for kernel() bodies of the form
 *
 *     acc = convolve(mask, mode, [&]{ return tap(mask(), in(mask)); })     HB_LOCAL_CONVOLVE
 *     acc = reduce(dom,  mode, [&]{ return tap(mask(dom), in(dom)); })     HB_LOCAL_REDUCE_DOMAIN
 *     output() = epilogue(acc)
 *
 * (dsl/kernel.hpp:241-296).  CONVOLVE visits all size_x*size_y taps, REDUCE_DOMAIN only the
 * non-zero domain taps (dsl/mask.hpp:112-126), both in row-major order; the first visited tap
 * initialises the accumulator, the rest are folded with `mode`.  Neighbour fetches go
 * through the boundary mode of the input accessor (dsl/image.hpp:574-612,
 * lib/AST/BorderHandling.cpp:41-120); CONSTANT uses the real constant like emitted code.
 */
typedef enum { HB_LOCAL_CONVOLVE = 0, HB_LOCAL_REDUCE_DOMAIN = 1 } hb_local_kind;
typedef enum {
  HB_TAP_MUL = 0, /* coef * in   (Gaussian, Sobel, Laplace, ...) */
  HB_TAP_IN = 1   /* in          (Dilate / Erode / Box blur over a Domain) */
} hb_tap;
typedef enum {
  HB_EPI_CAST = 0,           /* out = (Tout)acc                                           */
  HB_EPI_ADD_CAST = 1,       /* out = (Tout)(acc + p0)            Gaussian_Blur: +0.5f    */
  HB_EPI_ADD_CLAMP_CAST = 2, /* v = acc + p0; v = min(v,p2); v = max(v,p1); out = (Tout)v  Laplace */
  HB_EPI_DIVI_CAST = 3,      /* out = (Tout)(acc / (int)p0)        Harris: sum / norm      */
  HB_EPI_DIVF_CAST = 4       /* out = (Tout)(acc / (float)p0)      Box_Blur               */
} hb_epilogue;

typedef struct {
  hb_view in;            /* input Accessor  (region = boundary-handling window) */
  hb_view out;           /* IterationSpace over the output image */
  int kind;              /* hb_local_kind */
  int reduce_mode;       /* hb_reduce_mode */
  int tap;               /* hb_tap */
  int acc_dtype;         /* HB_F32 or HB_S32: the lambda's return type */
  int size_x, size_y;    /* Mask / Domain size; centre at size/2 (dsl/mask.hpp:177-184) */
  const float *coef_f32; /* HOST, size_y x size_x row-major; one of coef_f32 / coef_s32 for HB_TAP_MUL */
  const int *coef_s32;
  const unsigned char *domain; /* HOST, optional 0/1 footprint; NULL = derived from coef != 0
                                  (Mask ctor, dsl/mask.hpp:238-250) or all ones for HB_TAP_IN */
  int boundary;          /* hb_boundary */
  double boundary_const; /* value for HB_BOUNDARY_CONSTANT */
  int epilogue;          /* hb_epilogue */
  double epi_p[3];
} hb_local_desc;

int hb_local_op(const hb_local_desc *desc, void *stream);

/*
 * Bilateral filter: the `iterate(dom, ...)` body of
 * samples-public/3_Preprocessing/Bilateral_Filter/src/main.cpp:63-77 --
 *   c_r = 0.5f/(sigma_r*sigma_r); for each domain tap: diff = in(dom) - centre;
 *   s = expf(-c_r*diff*diff) * mask(dom); d += s; p += s*in(dom);
 *   out = u8: (uchar)(p/d + 0.5f)   f32: p/d
 */
typedef struct {
  hb_view in, out;
  int size;               /* sigma_s: mask is size x size */
  const float *coef_f32;  /* HOST */
  int sigma_r;
  int boundary;
  double boundary_const;
} hb_bilateral_desc;
int hb_bilateral(const hb_bilateral_desc *desc, void *stream);

/* ------------------------------------------------------------------ point operators */
/*
 * Replaces generated point-operator kernels (bodies using only in() / output(),
 * lib/Analysis/KernelStatistics.cpp:366-388).  Inputs may be interpolating accessors
 * (dsl/image.hpp:390-422; emitted mapping lib/AST/Interpolate.cpp:85-113): NN and LF with the
 * DSL's default CLAMP boundary (dsl/image.hpp:616-620).
 */
typedef enum {
  HB_POINT_COPY = 0,          /* out = a                    (Reduction_Sum, Subsample)              */
  HB_POINT_SQUARE = 1,        /* out = a*a                  (Harris Square1)                        */
  HB_POINT_MUL = 2,           /* out = a*b                  (Harris Square2)                        */
  HB_POINT_SUB = 3,           /* out = a-b                  (DifferenceOfGaussian)                  */
  HB_POINT_ADD = 4,           /* out = a+b                  (Restore)                               */
  HB_POINT_BLEND = 5,         /* out = a + b/2              (Blend)                                 */
  HB_POINT_SOBEL_COMBINE = 6, /* Sobel/src/main.cpp:76-98, p0 = norm                                */
  HB_POINT_HARRIS = 7         /* Harris_Corner/src/main.cpp:152-163, p0 = k, p1 = threshold (a,b,c) */
} hb_point_kind;

typedef struct {
  hb_view in[3];
  int interp[3];  /* hb_interp per input */
  int n_in;
  hb_view out;
  int op;         /* hb_point_kind */
  double p[2];
} hb_point_desc;
int hb_point_op(const hb_point_desc *desc, void *stream);

/* ------------------------------------------------------------------ global reductions */
/*
 * Kernel::reduce()/reduced_data() (dsl/kernel.hpp:121-161): fold a binary op over all pixels of
 * the region.  Replaces hipaccApplyReductionShared + hipacc_shared_reduction
 * (runtime/hipacc_cu.tpp:312-408, runtime/hipacc_cu_red.hpp:140-346).  Blocking like the
 * reference; the scalar is written to *result_host in the view's dtype.
 * MIN/MAX are bit-exact; float SUM is accumulated pairwise (float per thread, double across
 * threads) -- see DESIGN.md for why the reference's serial float fold is not a stable target.
 * Non-finite pixels: MIN / MAX are the IEEE minimum / maximum of the pixels that are numbers (a NaN never wins; +-inf take
 * part normally); SUM / PROD propagate NaN and inf as the arithmetic does.  The DSL's `l < r ? l : r` fold returns a value
 * that depends on where in the iteration order a NaN sits, which no parallel fold (the reference's included) reproduces.
 */
int hb_reduce(const hb_view *in, int reduce_mode, void *result_host, void *stream);
/* fused single pass over HBM: out[0]=min out[1]=max out[2]=sum (f32 images) */
int hb_reduce_minmaxsum_f32(const hb_view *in, float result_host[3], void *stream);
/* asynchronous form for multi-GPU: partial (min,max) as float and sum as double are left in
 * device memory ({float min, float max, double sum}, 16 bytes) for an NCCL all-reduce */
int hb_reduce_minmaxsum_f32_async(const hb_view *in, void *partials_device, void *stream);
/* asynchronous (capturable) form of hb_reduce for integer images (every mode) and float PROD: the scalar stays in device
 * memory at result_device (8 bytes, 8-byte aligned) -- a 32-bit accumulator for integer pixels (the C result is its
 * truncation to the pixel type, as hb_reduce returns it), a double for float PROD.  Float MIN / MAX / SUM come fused from
 * hb_reduce_minmaxsum_f32_async.  One reduction in flight per stream (per-stream scratch). */
int hb_reduce_async(const hb_view *in, int reduce_mode, void *result_device, void *stream);

/* ------------------------------------------------------------------ binning (histograms) */
/*
 * Kernel::binning() + binned_data(num_bins) (dsl/kernel.hpp:163-209): every pixel of the region
 * contributes `bin(INDEX(pixel)) = VALUE(pixel)` and equal indices are combined with reduce() = +
 * (the Histogram sample, samples-public/2_Global_Operators/Histogram/src/main.cpp:48-70).
 * Replaces hipaccApplyBinningSegmented + the generated segmented-binning kernel
 * (runtime/hipacc_cu.tpp:410-464, runtime/hipacc_cu_red.hpp:527-641).  Bins are uint32; indices
 * >= num_bins are dropped like the emitted Put helper does (runtime/hipacc_cpu_red.hpp:71-76).
 * A float index value v is defined (C conversion rules) for v > -1; v <= -1, +-inf and NaN are
 * undefined in C (x86-64 wraps them to a huge index or 0, the reference's CUDA backend
 * saturates) and are dropped here.  num_bins <= 2^22.
 * hb_binning is blocking and writes num_bins values to host memory (the reference returns
 * `new T[num_bins]`); the _async form leaves them in device memory (zeroed by the call) for an
 * NCCL all-reduce across row strips.
 */
typedef enum {
  HB_BIN_INDEX_SCALE = 0, /* idx = (uint)(pixel / p0 * num_bins), float arithmetic (Histogram sample: p0 = 255) */
  HB_BIN_INDEX_PIXEL = 1  /* idx = (uint)pixel */
} hb_bin_index;
typedef enum { HB_BIN_VALUE_ONE = 0 /* count */, HB_BIN_VALUE_PIXEL = 1 /* (uint)pixel */ } hb_bin_value;
typedef struct {
  hb_view in;      /* HB_F32 or HB_U8 */
  int num_bins;
  int index_kind;  /* hb_bin_index */
  int value_kind;  /* hb_bin_value */
  double p0;
} hb_binning_desc;
int hb_binning(const hb_binning_desc *desc, uint32_t *bins_host, void *stream);
int hb_binning_async(const hb_binning_desc *desc, uint32_t *bins_device, void *stream);

/* ------------------------------------------------------------------ fused pipelines */
/*
 * Harris corner detector, the 9-kernel pipeline of
 * samples-public/3_Preprocessing/Harris_Corner/src/main.cpp:230-305 (3x3 masks, CLAMP) fused
 * into one uchar -> uchar kernel: Sobel dx,dy (/6) -> squares -> 3x3 binomial (/16) -> response.
 * Intermediates keep the reference's short/int truncations; results are bit-identical to the
 * unfused pipeline.
 */
typedef struct {
  hb_view in, out;  /* HB_U8 */
  float k, threshold;
} hb_harris_desc;
int hb_harris(const hb_harris_desc *desc, void *stream);

/*
 * Pyramid level transitions of
 * samples-public/5_Other/Gaussian_Laplacian_Pyramid/src/main.cpp:199-248 (float pixels).
 * down: tmp = Gaussian(fine) [optional], coarse = NN-subsample(tmp), lap_fine = fine - LF(coarse)
 * up  : fine_gaus = LF(coarse_gaus) + lap_fine ; lap_fine = LF(coarse_lap) + lap_fine/2
 * Level sizes follow hipaccCreatePyramid (runtime/hipacc_cu.tpp:482-497): w/2, h/2 truncating.
 */
typedef struct {
  hb_view fine;      /* gaus(l-1), read */
  hb_view tmp;       /* blurred fine; data == NULL: not materialised */
  hb_view coarse;    /* gaus(l), written */
  hb_view lap_fine;  /* lap(l-1), written; data == NULL: skip DoG */
  int size;          /* Gaussian mask size (3,5,7) */
  const float *coef_f32;
} hb_pyr_down_desc;
int hb_pyr_down(const hb_pyr_down_desc *desc, void *stream);

/* DifferenceOfGaussian on its own (Gaussian_Laplacian_Pyramid/src/main.cpp:83-98 with an LF accessor): lap_fine =
 * fine - LF(coarse).  hb_pyr_down with lap_fine.data == NULL followed by hb_pyr_dog equals the fused hb_pyr_down bit
 * for bit; the split lets a sharded traversal run the DoG of a level in the shadow of the latency-bound coarse levels.
 * `coarse` may declare ghost rows (row strips). */
typedef struct {
  hb_view fine, coarse, lap_fine;
} hb_pyr_dog_desc;
int hb_pyr_dog(const hb_pyr_dog_desc *desc, void *stream);

typedef struct {
  hb_view coarse_gaus; /* gaus(l+1), read */
  hb_view coarse_lap;  /* lap(l+1), read */
  hb_view fine_gaus;   /* gaus(l), written */
  hb_view fine_lap;    /* lap(l), read + written */
} hb_pyr_up_desc;
int hb_pyr_up(const hb_pyr_up_desc *desc, void *stream);

/*
 * The COARSE END of a pyramid traversal in one launch: levels 0 .. levels-1 of a small pyramid (level 0 = its finest
 * level, e.g. level 4 of the 16384^2 pyramid), equivalent to
 *     for l = 1 .. levels-1 : hb_pyr_down(gaus[l-1] -> gaus[l], lap[l-1])        (fused form, no tmp)
 *     for l = levels-2 .. 0 : hb_pyr_up(gaus[l+1], lap[l+1] -> gaus[l], lap[l])
 * bit for bit, but as ONE cooperative kernel with grid-wide barriers between the transitions: below ~1024^2 a
 * transition is a few microseconds of work behind a launch of its own.  Needs exact halving at every transition
 * and 16-byte aligned rows (HB_ERR_UNSUPPORTED otherwise: use the per-level calls); lap[levels-1] is only read.
 */
#define HB_PYR_MAX_COARSE_LEVELS 8
typedef struct {
  int levels;
  hb_view gaus[HB_PYR_MAX_COARSE_LEVELS], lap[HB_PYR_MAX_COARSE_LEVELS];
  int size;               /* Gaussian mask size (3,5,7) */
  const float *coef_f32;
} hb_pyr_coarse_desc;
int hb_pyr_traverse_coarse(const hb_pyr_coarse_desc *desc, void *stream);

/* ------------------------------------------------------------------ peer-to-peer halo exchange */
/*
 * Row-strip sharding across the GPUs of one box (BASELINE.json: halo rows exchanged peer to peer over
 * NVLink).  The reference has no multi-device code; these entry points are the B200-native addition.
 * One process per GPU.  A rank exports its strip buffer and a control block (hb_halo_ctrl_create)
 * with hb_ipc_export, sends the 64-byte handles to its neighbours (any host channel), and maps
 * theirs with hb_ipc_open.  hb_halo_exchange then launches ONE kernel that pushes this rank's
 * `radius` top / bottom owned rows into the neighbours' ghost rows with peer stores and waits
 * (device side, system-scope flags in the control blocks) until its own ghost rows have arrived
 * and the previous ones have been consumed.  No host synchronisation, CUDA-graph replayable.
 * Buffers must be whole allocations (hb_image_create / cudaMalloc base pointers).
 */
#define HB_IPC_HANDLE_BYTES 64
typedef struct { unsigned char handle[HB_IPC_HANDLE_BYTES]; } hb_ipc_mem;
int hb_ipc_export(const void *device_ptr, hb_ipc_mem *out);
int hb_ipc_open(const hb_ipc_mem *in, void **peer_ptr);
int hb_ipc_close(void *peer_ptr);
int hb_halo_ctrl_create(void **ctrl);
int hb_halo_ctrl_destroy(void *ctrl);
/* exchanges completed so far and whether a wait ever timed out (~2 s; a neighbour never arrived) */
int hb_halo_status(const void *ctrl, int *exchanges, int *timed_out);

typedef struct {
  void *buf;                 /* this rank's strip buffer, row 0 = first ghost row */
  size_t pitch_bytes, row_bytes;
  int ghost_top, rows, radius;
  void *ctrl;                /* this rank's control block */
  void *up_buf, *up_ctrl;    /* upper neighbour's buffer / control block (peer pointers) or NULL */
  size_t up_pitch_bytes;
  int up_ghost_top, up_rows; /* its layout: my rows land at row up_ghost_top + up_rows */
  void *down_buf, *down_ctrl;
  size_t down_pitch_bytes;
  int down_ghost_top;        /* my rows land at row down_ghost_top - radius */
} hb_halo_desc;
int hb_halo_exchange(const hb_halo_desc *desc, void *stream);
/* up to 4 strip buffers exchanged concurrently by one launch (one CTA each), e.g. the Gaussian and the
 * Laplacian level before a pyramid up-transition */
int hb_halo_exchange_batch(const hb_halo_desc *const *descs, int n, void *stream);

/*
 * All-gather of row strips over peer memory (coarse pyramid levels): every rank keeps the FULL image of a level,
 * owns rows [row0, row0 + rows) of it and pushes them into the same rows of every peer's copy with one launch
 * (one CTA per peer, device-side flags like hb_halo_exchange; CUDA-graph replayable).  All copies share the
 * layout (pitch, height); `slot` = rank numbers used to index the flag arrays of the control blocks
 * (hb_halo_ctrl_create).  With this the sharded pyramid needs no halo exchange below the gather level.
 */
#define HB_MAX_PEERS 15
typedef struct {
  void *buf;                 /* this rank's full copy */
  size_t pitch_bytes, row_bytes;
  int row0, rows;            /* the rows this rank owns and publishes */
  void *ctrl;                /* this rank's control block */
  int my_slot, n_peers;
  void *peer_buf[HB_MAX_PEERS], *peer_ctrl[HB_MAX_PEERS];
  int peer_slot[HB_MAX_PEERS];
} hb_gather_desc;
int hb_allgather_rows(const hb_gather_desc *desc, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HIPACC_B200_H */
