// hb_internal.h -- host-side plumbing shared by the translation units of libhipacc_b200.so
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/hipacc_b200.h"

namespace hb {

void log_msg(int level, const char *fmt, ...);
void set_last_error(const std::string &s);
extern std::atomic<long long> g_launches;
extern bool g_timing;

// checkErr of the reference (runtime/hipacc_cu.hpp:69-75): log and continue, but also report a status
inline int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return HB_OK;
    log_msg(2, "ERROR: %s (%d): %s: %s", what, (int)e, cudaGetErrorName(e), cudaGetErrorString(e));
    return HB_ERR_CUDA;
}

// Brackets one operator call with events when hb_set_timing(1) (the reference's print_timing,
// runtime/hipacc_cu_standalone.hpp:297-326), and checks the launch.
struct OpScope {
    cudaStream_t stream;
    const char *name;
    bool timed;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    OpScope(cudaStream_t s, const char *n);
    int finish();  // returns status of the launches issued inside the scope
};
bool stream_is_capturing(cudaStream_t s);
// per-(device, stream) reduction scratch must exist before a capture starts (hb_reduce.cu)
int reserve_reduce_scratch(cudaStream_t s);
int reserve_pyramid_scratch(cudaStream_t s);   // grid-barrier words of hb_pyr_traverse_coarse (hb_pyramid.cu)

inline hb_view norm_view(const hb_view &v) {
    hb_view o = v;
    if (o.width <= 0 || o.height <= 0) { o.width = o.img_width; o.height = o.img_height; o.offset_x = 0; o.offset_y = 0; }
    return o;
}
inline int dtype_size(int dt) {
    switch (dt) {
    case HB_U8: case HB_S8: return 1;
    case HB_U16: case HB_S16: return 2;
    case HB_U16X4: case HB_S16X4: return 8;
    case HB_S32X4: case HB_U32X4: case HB_F32X4: return 16;
    default: return 4;   // 32-bit scalars, HB_U8X4, HB_S8X4
    }
}
inline bool view_ok(const hb_view &v) {
    return v.data && v.img_width > 0 && v.img_height > 0 && v.stride >= v.img_width && v.dtype >= HB_U8 && v.dtype <= HB_DTYPE_LAST &&
           v.offset_x >= 0 && v.offset_y >= 0 && v.offset_x + v.width <= v.img_width && v.offset_y + v.height <= v.img_height &&
           v.ghost_top >= 0 && v.ghost_bottom >= 0 && v.offset_y - v.ghost_top >= 0 &&
           v.offset_y + v.height + v.ghost_bottom <= v.img_height;
}

// a uchar4 view as the uchar image of its channel elements (4x wider); `unit` = elements per pixel
inline bool is_x4(int dt) { return dt >= HB_U8X4 && dt <= HB_F32X4; }
inline int channel_dtype(int dt) {
    switch (dt) {
    case HB_U8X4: return HB_U8; case HB_S8X4: return HB_S8; case HB_U16X4: return HB_U16; case HB_S16X4: return HB_S16;
    case HB_S32X4: return HB_S32; case HB_U32X4: return HB_U32; case HB_F32X4: return HB_F32;
    default: return dt;
    }
}
inline hb_view as_channels(const hb_view &v) {
    hb_view o = v;
    if (is_x4(v.dtype)) {
        o.dtype = channel_dtype(v.dtype);
        o.img_width *= 4; o.stride *= 4; o.width *= 4; o.offset_x *= 4;
    }
    return o;
}

int sm_count();
// grid of a streaming kernel that walks `chunks` work units: `per_sm` CTAs per SM (tuning override: HB_STREAM_CTAS_PER_SM,
// 0 = one CTA per chunk).  Measured on B200 (tools/probe/copy_probe.cu): a one-shot grid streams ~9 % faster than a
// persistent grid whose CTAs march over memory in lockstep.
long long stream_grid(long long chunks, int per_sm);

}  // namespace hb

#define HB_REQUIRE(cond, status, ...)                       \
    do {                                                    \
        if (!(cond)) {                                      \
            hb::log_msg(2, __VA_ARGS__);                    \
            return (status);                                \
        }                                                   \
    } while (0)
