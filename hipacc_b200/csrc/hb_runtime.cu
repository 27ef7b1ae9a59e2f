// hb_runtime.cu -- device / runtime layer of the C ABI: init, image memory, copies, timing, logging.
// Replaces runtime/hipacc_cu_standalone.hpp:113-216,277-329 and runtime/hipacc_cu.tpp:43-193
// (paths relative to the Hipacc tree).  No textures, no NVRTC, no OpenCL, no CPU fallback.
#include "hb_internal.h"
#include "hb_tma.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

namespace hb {

std::atomic<long long> g_launches{0};
bool g_timing = false;
static hb_log_fn g_log = nullptr;
static thread_local std::string g_last_error;
static thread_local float g_last_ms = 0.0f;  // thread_local like HipaccKernelTimingBase (hipacc_base_standalone.hpp:35-39)
static int g_device = -1;
static int g_sms = 0;
// timing events: one pair per host thread and device (the reference's timing singleton is thread_local too), so
// two threads that time operators on different streams do not share an event
struct TimingEvents {
    cudaEvent_t ev0[16] = {}, ev1[16] = {};
};
static thread_local TimingEvents t_events;

void set_last_error(const std::string &s) { g_last_error = s; }

void log_msg(int level, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (level >= 2) g_last_error = buf;
    if (g_log) g_log(level, buf);
    else fprintf(level == 0 ? stdout : stderr, "<HIPACC-B200:> %s\n", buf);
}

int sm_count() {
    if (g_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}

bool stream_is_capturing(cudaStream_t s) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return false; }
    return st != cudaStreamCaptureStatusNone;
}

// Timing brackets the operator with events and SYNCHRONISES (the reference's print_timing).  Inside a stream capture
// (hb_graph_begin .. hb_graph_end) an event synchronise is illegal and would invalidate the capture, so a captured
// operator is never timed: the launch is recorded into the graph and hb_last_kernel_ms keeps its previous value.
OpScope::OpScope(cudaStream_t s, const char *n) : stream(s), name(n), timed(false) {
    if (g_timing && !stream_is_capturing(s)) {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 15;
        if (!t_events.ev0[dev]) { cudaEventCreate(&t_events.ev0[dev]); cudaEventCreate(&t_events.ev1[dev]); }
        ev0 = t_events.ev0[dev]; ev1 = t_events.ev1[dev];
        timed = cudaEventRecord(ev0, stream) == cudaSuccess;
    }
}
int OpScope::finish() {
    int rc = check_cuda(cudaGetLastError(), name);
    if (timed) {
        cudaEventRecord(ev1, stream);
        rc |= check_cuda(cudaEventSynchronize(ev1), name);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        g_last_ms = ms;
    }
    return rc ? HB_ERR_CUDA : HB_OK;
}

// ---- TMA tensor maps (hb_tma.cuh).  cuTensorMapEncodeTiled is resolved through cudart so the
// library does not link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else {
            cudaGetLastError();
            log_msg(1, "WARNING: cuTensorMapEncodeTiled unavailable; local operators use the all-threads tile loader");
        }
    }
    return fn;
}

long long stream_grid(long long chunks, int per_sm) {
    static int override_ = -2;
    if (override_ == -2) {
        const char *e = getenv("HB_STREAM_CTAS_PER_SM");
        override_ = e ? atoi(e) : -1;
    }
    if (override_ >= 0) per_sm = override_;
    if (chunks < 1) chunks = 1;
    if (per_sm <= 0) return chunks;
    const long long cap = (long long)sm_count() * per_sm;
    return chunks < cap ? chunks : cap;
}

bool tma_addressable(const void *base, int dtype, int stride_px) {
    const size_t es = dtype_size(dtype);
    return (reinterpret_cast<uintptr_t>(base) % 16 == 0) && (((size_t)stride_px * es) % 16 == 0);
}

bool make_tile_map(CUtensorMap *out, const void *base, int dtype, int img_w, int img_h, int stride_px, int box_w, int box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || !tma_addressable(base, dtype, stride_px)) return false;
    const size_t es = dtype_size(dtype);
    if (((size_t)box_w * es) % 16 != 0 || box_w > 256 || box_h > 256 || box_w <= 0 || box_h <= 0) return false;
    CUtensorMapDataType dt;
    switch (dtype) {
    case HB_U8: case HB_S8: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    case HB_U16: case HB_S16: dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; break;
    case HB_S32: dt = CU_TENSOR_MAP_DATA_TYPE_INT32; break;
    case HB_U32: dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; break;
    case HB_F32: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    default: return false;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)img_w, (cuuint64_t)img_h};
    const cuuint64_t gstride[1] = {(cuuint64_t)stride_px * es};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    static int promo = -1;  // HB_TMA_L2PROMO = 0 none, 1 64B, 2 128B, 3 256B (tuning knob)
    if (promo < 0) {
        const char *e = getenv("HB_TMA_L2PROMO");
        promo = e ? atoi(e) : 2;
    }
    const CUtensorMapL2promotion l2 = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    // descriptors are cached per (pointer, extent, pitch, box): an operator that runs every frame on the same image
    // does not pay the driver's encode on each launch
    struct Key {
        const void *base; int dtype, w, h, stride, bw, bh, promo;
        bool operator<(const Key &o) const { return memcmp(this, &o, sizeof(Key)) < 0; }
    };
    static std::mutex mu;
    static std::map<Key, CUtensorMap> cache;
    Key key;
    memset(&key, 0, sizeof(key));
    key.base = base; key.dtype = dtype; key.w = img_w; key.h = img_h; key.stride = stride_px; key.bw = box_w; key.bh = box_h; key.promo = promo;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    const CUresult r = fn(out, dt, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        log_msg(1, "WARNING: cuTensorMapEncodeTiled failed (%d) for a %dx%d image, box %dx%d", (int)r, img_w, img_h, box_w, box_h);
        return false;
    }
    if (cache.size() >= 256) cache.clear();
    cache.emplace(key, *out);
    return true;
}

__global__ void timestamp_kernel(unsigned long long *slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}

}  // namespace hb

using namespace hb;

extern "C" {

int hb_debug_timestamp(void *slot_device, void *stream) {
    HB_REQUIRE(slot_device, HB_ERR_INVALID, "hb_debug_timestamp: null slot");
    timestamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(static_cast<unsigned long long *>(slot_device));
    return check_cuda(cudaGetLastError(), "hb_debug_timestamp");
}

int hb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hb_init(int device) {
    int n = hb_device_count();
    HB_REQUIRE(n > 0, HB_ERR_NO_DEVICE, "hb_init: no CUDA device visible (this library has no CPU fallback)");
    HB_REQUIRE(device >= 0 && device < n, HB_ERR_INVALID, "hb_init: device %d out of range (0..%d)", device, n - 1);
    int rc = check_cuda(cudaSetDevice(device), "cudaSetDevice()");
    if (rc) return rc;
    cudaDeviceProp p;
    rc = check_cuda(cudaGetDeviceProperties(&p, device), "cudaGetDeviceProperties()");
    if (rc) return rc;
    g_device = device;
    g_sms = p.multiProcessorCount;
    HB_REQUIRE(p.major >= 10, HB_ERR_UNSUPPORTED, "hb_init: device %d (%s, sm_%d%d) is not a Blackwell part; kernels are built for sm_100a only",
               device, p.name, p.major, p.minor);
    return HB_OK;
}

int hb_sm_count(void) { return sm_count(); }
void hb_set_log_callback(hb_log_fn fn) { g_log = fn; }
const char *hb_last_error(void) { return g_last_error.c_str(); }
void hb_set_timing(int enabled) { g_timing = enabled != 0; }
float hb_last_kernel_ms(void) { return g_last_ms; }
long long hb_launch_count(void) { return g_launches.load(); }
int hb_stream_synchronize(void *stream) { return check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize()"); }

int hb_stream_create(void **stream) {
    HB_REQUIRE(stream, HB_ERR_INVALID, "hb_stream_create: null argument");
    cudaStream_t s = nullptr;
    int rc = check_cuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreateWithFlags()");
    *stream = s;
    return rc;
}
int hb_stream_destroy(void *stream) {
    if (!stream) return HB_OK;
    return check_cuda(cudaStreamDestroy((cudaStream_t)stream), "cudaStreamDestroy()");
}

int hb_image_create(int dtype, int width, int height, int alignment, hb_view *out) {
    HB_REQUIRE(out && width > 0 && height > 0 && dtype >= HB_U8 && dtype <= HB_DTYPE_LAST, HB_ERR_INVALID, "hb_image_create: bad arguments");
    const int es = dtype_size(dtype);
    if (alignment <= 0) alignment = 256;
    // alignment has to be a multiple of sizeof(T) (runtime/hipacc_cu.tpp:54-57)
    alignment = ((alignment + es - 1) / es) * es;
    const size_t per = (size_t)alignment / es;
    const size_t stride = (((size_t)width + per - 1) / per) * per;
    void *p = nullptr;
    int rc = check_cuda(cudaMalloc(&p, stride * (size_t)height * es), "cudaMalloc()");
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    out->data = p; out->dtype = dtype; out->img_width = width; out->img_height = height; out->stride = (int)stride;
    out->width = width; out->height = height;
    return HB_OK;
}

int hb_image_destroy(hb_view *img) {
    if (!img || !img->data) return HB_OK;
    int rc = check_cuda(cudaFree(img->data), "cudaFree()");
    img->data = nullptr;
    return rc;
}

int hb_image_wrap(void *device_ptr, int dtype, int width, int height, int stride, hb_view *out) {
    HB_REQUIRE(out && device_ptr && width > 0 && height > 0 && stride >= width, HB_ERR_INVALID, "hb_image_wrap: bad arguments");
    memset(out, 0, sizeof(*out));
    out->data = device_ptr; out->dtype = dtype; out->img_width = width; out->img_height = height; out->stride = stride;
    out->width = width; out->height = height;
    return HB_OK;
}

int hb_image_write(const hb_view *img, const void *host, void *stream) {
    HB_REQUIRE(img && img->data && host, HB_ERR_INVALID, "hb_image_write: bad arguments");
    const size_t es = dtype_size(img->dtype);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemcpy2DAsync(img->data, (size_t)img->stride * es, host, (size_t)img->img_width * es,
                                          (size_t)img->img_width * es, img->img_height, cudaMemcpyHostToDevice, s),
                        "cudaMemcpy2DAsync(H2D)");
    if (rc) return rc;
    return check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize()");  // blocking like the reference
}

int hb_image_read(const hb_view *img, void *host, void *stream) {
    HB_REQUIRE(img && img->data && host, HB_ERR_INVALID, "hb_image_read: bad arguments");
    const size_t es = dtype_size(img->dtype);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemcpy2DAsync(host, (size_t)img->img_width * es, img->data, (size_t)img->stride * es,
                                          (size_t)img->img_width * es, img->img_height, cudaMemcpyDeviceToHost, s),
                        "cudaMemcpy2DAsync(D2H)");
    if (rc) return rc;
    return check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize()");
}

static int region_copy_async(const hb_view *v_, void *host, size_t host_pitch, bool to_device, void *stream, const char *who) {
    HB_REQUIRE(v_ && v_->data && host, HB_ERR_INVALID, "%s: bad arguments", who);
    hb_view v = norm_view(*v_);
    HB_REQUIRE(view_ok(v), HB_ERR_INVALID, "%s: malformed view", who);
    const size_t es = dtype_size(v.dtype), row = (size_t)v.width * es;
    HB_REQUIRE(host_pitch >= row, HB_ERR_INVALID, "%s: host pitch smaller than a region row", who);
    char *d = (char *)v.data + ((size_t)v.offset_y * v.stride + v.offset_x) * es;
    const cudaError_t e = to_device ? cudaMemcpy2DAsync(d, (size_t)v.stride * es, host, host_pitch, row, v.height, cudaMemcpyHostToDevice, (cudaStream_t)stream)
                                    : cudaMemcpy2DAsync(host, host_pitch, d, (size_t)v.stride * es, row, v.height, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    return check_cuda(e, who);
}
int hb_image_write_region_async(const hb_view *region, const void *host, size_t host_pitch_bytes, void *stream) {
    return region_copy_async(region, const_cast<void *>(host), host_pitch_bytes, true, stream, "hb_image_write_region_async");
}
int hb_image_read_region_async(const hb_view *region, void *host, size_t host_pitch_bytes, void *stream) {
    return region_copy_async(region, host, host_pitch_bytes, false, stream, "hb_image_read_region_async");
}

int hb_graph_begin(void *stream) {
    HB_REQUIRE(stream, HB_ERR_INVALID, "hb_graph_begin: capture needs a non-default stream");
    // allocations are illegal while capturing: the scratch a captured reduction bakes into the graph is reserved now
    // (one set per stream, never reallocated, so replays stay valid)
    int rc = reserve_reduce_scratch((cudaStream_t)stream);
    if (!rc) rc = reserve_pyramid_scratch((cudaStream_t)stream);
    if (rc) return rc;
    return check_cuda(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture()");
}
int hb_graph_end(void *stream, hb_graph **out) {
    HB_REQUIRE(stream && out, HB_ERR_INVALID, "hb_graph_end: null argument");
    cudaGraph_t g = nullptr;
    int rc = check_cuda(cudaStreamEndCapture((cudaStream_t)stream, &g), "cudaStreamEndCapture()");
    if (rc) return rc;
    cudaGraphExec_t ex = nullptr;
    rc = check_cuda(cudaGraphInstantiate(&ex, g, 0), "cudaGraphInstantiate()");
    cudaGraphDestroy(g);
    if (rc) return rc;
    *out = reinterpret_cast<hb_graph *>(ex);
    return HB_OK;
}
int hb_graph_launch(hb_graph *g, void *stream) {
    HB_REQUIRE(g, HB_ERR_INVALID, "hb_graph_launch: null graph");
    return check_cuda(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(g), (cudaStream_t)stream), "cudaGraphLaunch()");
}
int hb_graph_destroy(hb_graph *g) {
    if (!g) return HB_OK;
    return check_cuda(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(g)), "cudaGraphExecDestroy()");
}

int hb_image_copy(const hb_view *src, const hb_view *dst, void *stream) {
    HB_REQUIRE(src && dst && src->data && dst->data, HB_ERR_INVALID, "hb_image_copy: bad arguments");
    HB_REQUIRE(src->img_width == dst->img_width && src->img_height == dst->img_height && src->dtype == dst->dtype, HB_ERR_INVALID,
               "hb_image_copy: extents / types differ");
    const size_t es = dtype_size(src->dtype);
    return check_cuda(cudaMemcpy2DAsync(dst->data, (size_t)dst->stride * es, src->data, (size_t)src->stride * es,
                                        (size_t)src->img_width * es, src->img_height, cudaMemcpyDeviceToDevice, (cudaStream_t)stream),
                      "cudaMemcpy2DAsync(D2D)");
}

int hb_image_copy_region(const hb_view *src_, const hb_view *dst_, void *stream) {
    HB_REQUIRE(src_ && dst_ && src_->data && dst_->data, HB_ERR_INVALID, "hb_image_copy_region: bad arguments");
    hb_view src = norm_view(*src_), dst = norm_view(*dst_);
    HB_REQUIRE(src.dtype == dst.dtype && view_ok(src) && view_ok(dst), HB_ERR_INVALID, "hb_image_copy_region: bad views");
    HB_REQUIRE(src.width <= dst.img_width - dst.offset_x && src.height <= dst.img_height - dst.offset_y, HB_ERR_INVALID,
               "hb_image_copy_region: source region does not fit at the destination offset");
    const size_t es = dtype_size(src.dtype);
    char *d = (char *)dst.data + ((size_t)dst.offset_y * dst.stride + dst.offset_x) * es;
    const char *s = (const char *)src.data + ((size_t)src.offset_y * src.stride + src.offset_x) * es;
    return check_cuda(cudaMemcpy2DAsync(d, (size_t)dst.stride * es, s, (size_t)src.stride * es, (size_t)src.width * es, src.height,
                                        cudaMemcpyDeviceToDevice, (cudaStream_t)stream),
                      "cudaMemcpy2DAsync(D2D region)");
}

}  // extern "C"
