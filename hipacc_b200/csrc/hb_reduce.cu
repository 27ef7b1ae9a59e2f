// hb_reduce.cu -- global reductions (Kernel::reduce()/reduced_data(), dsl/kernel.hpp:121-161) for sm_100a.
//
// Replaces hipaccApplyReductionShared + hipacc_shared_reduction (runtime/hipacc_cu.tpp:312-408,
// runtime/hipacc_cu_red.hpp:140-346: one pixel per thread, warp-synchronous shared-memory tree,
// 3 cudaMalloc/cudaFree + a blocking memcpy per call).
//
// One pass over HBM: 16 CTAs per SM stride over the row chunks with
// 16-byte loads, each thread keeps private accumulators, then warp shuffles -> one partial per
// CTA -> the last CTA to finish (atomic ticket) folds the partials in a FIXED order, so results are
// deterministic for a given grid.  min, max and sum are produced by the same pass (fused).
// Float sums: float per thread, double across threads (DESIGN.md "Float SUM").
#include "hb_common.cuh"
#include "hb_internal.h"

#include <cfloat>
#include <climits>
#include <cstring>

namespace hb {

struct MMS {  // min, max, sum partial
    float mn, mx;
    double sum;
};

struct ReduceParams {
    const void *in;
    int stride, w, h, ox, oy;
    MMS *partials;          // gridDim.x entries
    unsigned *ticket;       // zero before launch, reset by the last CTA
    void *result;           // device: MMS (f32 fused) or 8-byte scalar slot (generic)
    int mode;
};

__device__ __forceinline__ MMS mms_combine(MMS a, MMS b) {
    MMS r;
    r.mn = b.mn < a.mn ? b.mn : a.mn;
    r.mx = b.mx > a.mx ? b.mx : a.mx;
    r.sum = a.sum + b.sum;
    return r;
}
__device__ __forceinline__ MMS mms_shfl_down(MMS v, int d) {
    MMS r;
    r.mn = __shfl_down_sync(0xffffffffu, v.mn, d);
    r.mx = __shfl_down_sync(0xffffffffu, v.mx, d);
    r.sum = __shfl_down_sync(0xffffffffu, v.sum, d);
    return r;
}

constexpr int RT = 256;  // threads per CTA

// block-level fold of per-thread partials; valid in thread 0
__device__ __forceinline__ MMS block_fold(MMS v) {
    __shared__ MMS warp_part[RT / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = mms_combine(v, mms_shfl_down(v, d));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < RT / 32 ? warp_part[lane] : MMS{FLT_MAX * 2.0f, -FLT_MAX * 2.0f, 0.0};
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = mms_combine(v, mms_shfl_down(v, d));
    }
    return v;
}

// Work unit = one "chunk": CHUNK_V consecutive 16-byte vectors of one row (RT threads x UNR independent loads in
// flight per thread).  Rows are split relative to a 16-byte aligned column; the few scalar pixels before / after
// the aligned span of a row are folded by the first lanes of the row's first chunk.
constexpr int UNR = 4;
constexpr int CHUNK_V = RT * UNR;

__global__ void __launch_bounds__(RT) reduce_mms_f32_kernel(const __grid_constant__ ReduceParams p) {
    const float *in = static_cast<const float *>(p.in);
    const float INF = __int_as_float(0x7f800000);
    float mn = INF, mx = -INF;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;  // 4 independent float chains per thread

    const uintptr_t base_addr = reinterpret_cast<uintptr_t>(in) + (size_t)p.ox * sizeof(float);
    const bool aligned_rows = (p.stride % 4 == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0);
    const int head = aligned_rows ? (int)(((16 - (base_addr & 15)) & 15) / 4) : 0;  // scalar pixels before alignment
    const int head_n = head < p.w ? head : p.w;
    const int nvec = aligned_rows ? (p.w - head_n) / 4 : 0;
    const int tail0 = head_n + nvec * 4;
    const int cpr = nvec > 0 ? (nvec + CHUNK_V - 1) / CHUNK_V : 1;  // chunks per row
    const long long total = (long long)cpr * p.h;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int y = (int)(u / cpr), c = (int)(u - (long long)y * cpr);
        const float *row = in + (size_t)(p.oy + y) * p.stride + p.ox;
        const float4 *vrow = reinterpret_cast<const float4 *>(row + head_n);
        const int v0 = c * CHUNK_V + threadIdx.x;
        float4 v[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const int i = v0 + k * RT;
            // neutral element for lanes past the end of the row: min/max ignore +-inf halves, sum adds 0
            v[k] = i < nvec ? __ldcs(vrow + i) : make_float4(INF, INF, INF, INF);
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            if (v0 + k * RT < nvec) {
                mn = fminf(mn, fminf(fminf(v[k].x, v[k].y), fminf(v[k].z, v[k].w)));
                mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
                s0 += v[k].x; s1 += v[k].y; s2 += v[k].z; s3 += v[k].w;
            }
        }
        if (c == 0) {  // scalar head / tail pixels of this row
            const int nscal = head_n + (p.w - tail0);
            for (int k = threadIdx.x; k < nscal; k += RT) {
                const float e = row[k < head_n ? k : tail0 + (k - head_n)];
                mn = fminf(mn, e); mx = fmaxf(mx, e); s0 += e;
            }
        }
    }
    MMS v{mn, mx, ((double)s0 + (double)s1) + ((double)s2 + (double)s3)};
    v = block_fold(v);

    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        p.partials[blockIdx.x] = v;
        __threadfence();
        const unsigned t = atomicAdd(p.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fixed-order fold of the per-CTA partials: thread t takes partials t, t+RT, ... then block fold
        MMS a{INF, -INF, 0.0};
        for (unsigned i = threadIdx.x; i < gridDim.x; i += RT) a = mms_combine(a, p.partials[i]);
        __syncthreads();
        a = block_fold(a);
        if (threadIdx.x == 0) {
            *static_cast<MMS *>(p.result) = a;
            *p.ticket = 0;  // ready for the next call
        }
    }
}

// generic single-op reduction for integer pixel types (accumulated in the pixel type's modular arithmetic)
template <typename T>
__global__ void __launch_bounds__(RT) reduce_int_kernel(const __grid_constant__ ReduceParams p, long long *gpart) {
    const T *in = static_cast<const T *>(p.in);
    long long acc = p.mode == HB_REDUCE_SUM ? 0 : p.mode == HB_REDUCE_PROD ? 1 : p.mode == HB_REDUCE_MIN ? LLONG_MAX : LLONG_MIN;
    const long long total = (long long)p.w * p.h;
    for (long long u = blockIdx.x * (long long)RT + threadIdx.x; u < total; u += (long long)gridDim.x * RT) {
        const int y = (int)(u / p.w), x = (int)(u - (long long)y * p.w);
        const long long e = (long long)in[(size_t)(p.oy + y) * p.stride + p.ox + x];
        acc = p.mode == HB_REDUCE_SUM ? acc + e : p.mode == HB_REDUCE_PROD ? (long long)(T)(acc * e) : p.mode == HB_REDUCE_MIN ? (e < acc ? e : acc) : (e > acc ? e : acc);
    }
    __shared__ long long sh[RT];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = RT / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            const long long a = sh[threadIdx.x], b = sh[threadIdx.x + d];
            sh[threadIdx.x] = p.mode == HB_REDUCE_SUM ? a + b : p.mode == HB_REDUCE_PROD ? (long long)(T)(a * b) : p.mode == HB_REDUCE_MIN ? (b < a ? b : a) : (b > a ? b : a);
        }
        __syncthreads();
    }
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        gpart[blockIdx.x] = sh[0];
        __threadfence();
        is_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        long long a = gpart[0];
        for (unsigned i = 1; i < gridDim.x; ++i) {
            const long long b = gpart[i];
            a = p.mode == HB_REDUCE_SUM ? a + b : p.mode == HB_REDUCE_PROD ? (long long)(T)(a * b) : p.mode == HB_REDUCE_MIN ? (b < a ? b : a) : (b > a ? b : a);
        }
        *static_cast<long long *>(p.result) = (long long)(T)a;  // result in the pixel type (data_t reduce(data_t, data_t))
        *p.ticket = 0;
    }
}

// per-device scratch: partials, ticket, result slot and a pinned host mirror (allocated once; the
// reference allocates and frees three buffers per call, runtime/hipacc_cu.tpp:331-391)
struct Scratch {
    MMS *partials = nullptr;
    long long *ipart = nullptr;
    unsigned *ticket = nullptr;
    MMS *result = nullptr;
    MMS *host = nullptr;
    int cap = 0;
};
static Scratch g_scratch[16];

static int get_scratch(Scratch **out, int blocks) {
    int dev = 0;
    cudaGetDevice(&dev);
    Scratch &s = g_scratch[dev & 15];
    if (s.cap < blocks) {
        if (s.partials) { cudaFree(s.partials); cudaFree(s.ipart); }
        int rc = check_cuda(cudaMalloc(&s.partials, sizeof(MMS) * blocks), "cudaMalloc(reduce partials)");
        rc |= check_cuda(cudaMalloc(&s.ipart, sizeof(long long) * blocks), "cudaMalloc(reduce partials)");
        if (rc) return rc;
        s.cap = blocks;
    }
    if (!s.ticket) {
        int rc = check_cuda(cudaMalloc(&s.ticket, sizeof(unsigned)), "cudaMalloc(ticket)");
        rc |= check_cuda(cudaMemset(s.ticket, 0, sizeof(unsigned)), "cudaMemset(ticket)");
        rc |= check_cuda(cudaMalloc(&s.result, sizeof(MMS)), "cudaMalloc(result)");
        rc |= check_cuda(cudaMallocHost(&s.host, sizeof(MMS)), "cudaMallocHost(result)");
        if (rc) return rc;
    }
    *out = &s;
    return HB_OK;
}

static int reduce_grid(long long units) {
    long long blocks = (units + RT - 1) / RT;
    const long long cap = (long long)sm_count() * 8;  // persistent: 8 CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : (int)blocks;
}

static int launch_mms(const hb_view &v, void *result_dev, cudaStream_t s) {
    const long long cpr = (v.width / 4 + CHUNK_V - 1) / CHUNK_V;
    const long long chunks = (cpr < 1 ? 1 : cpr) * v.height;
    const int blocks = (int)stream_grid(chunks, 16);   // measured best of 8 / 16 / 32 / 64 / one-shot (tools/stream_grid_sweep.sh)
    Scratch *sc = nullptr;
    int rc = get_scratch(&sc, blocks);
    if (rc) return rc;
    ReduceParams p{v.data, v.stride, v.width, v.height, v.offset_x, v.offset_y, sc->partials, sc->ticket, result_dev ? result_dev : sc->result, 0};
    reduce_mms_f32_kernel<<<blocks, RT, 0, s>>>(p);
    g_launches++;
    return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" int hb_reduce_minmaxsum_f32_async(const hb_view *in_, void *partials_device, void *stream) {
    HB_REQUIRE(in_ && partials_device, HB_ERR_INVALID, "hb_reduce_minmaxsum_f32_async: null argument");
    hb_view v = norm_view(*in_);
    HB_REQUIRE(view_ok(v) && v.dtype == HB_F32, HB_ERR_INVALID, "hb_reduce_minmaxsum_f32_async: needs a valid f32 view");
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_reduce_minmaxsum_f32_async");
    int rc = launch_mms(v, partials_device, s);
    if (rc) return rc;
    return scope.finish();
}

extern "C" int hb_reduce_minmaxsum_f32(const hb_view *in_, float result_host[3], void *stream) {
    HB_REQUIRE(in_ && result_host, HB_ERR_INVALID, "hb_reduce_minmaxsum_f32: null argument");
    hb_view v = norm_view(*in_);
    HB_REQUIRE(view_ok(v) && v.dtype == HB_F32, HB_ERR_INVALID, "hb_reduce_minmaxsum_f32: needs a valid f32 view");
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_reduce_minmaxsum_f32");
    int rc = launch_mms(v, nullptr, s);
    if (rc) return rc;
    rc = scope.finish();
    Scratch *sc = nullptr;
    get_scratch(&sc, 1);
    rc |= check_cuda(cudaMemcpyAsync(sc->host, sc->result, sizeof(MMS), cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(result)");
    rc |= check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize()");  // blocking like the reference
    result_host[0] = sc->host->mn;
    result_host[1] = sc->host->mx;
    result_host[2] = (float)sc->host->sum;
    return rc ? HB_ERR_CUDA : HB_OK;
}

extern "C" int hb_reduce(const hb_view *in_, int mode, void *result_host, void *stream) {
    HB_REQUIRE(in_ && result_host, HB_ERR_INVALID, "hb_reduce: null argument");
    HB_REQUIRE(mode >= HB_REDUCE_SUM && mode <= HB_REDUCE_PROD, HB_ERR_INVALID, "hb_reduce: bad mode");
    hb_view v = norm_view(*in_);
    HB_REQUIRE(view_ok(v), HB_ERR_INVALID, "hb_reduce: malformed view");
    cudaStream_t s = (cudaStream_t)stream;
    if (v.dtype == HB_F32) {
        HB_REQUIRE(mode != HB_REDUCE_PROD, HB_ERR_UNSUPPORTED, "hb_reduce: float PROD has no device kernel; no CPU fallback");
        float r[3];
        int rc = hb_reduce_minmaxsum_f32(&v, r, stream);
        *static_cast<float *>(result_host) = mode == HB_REDUCE_MIN ? r[0] : mode == HB_REDUCE_MAX ? r[1] : r[2];
        return rc;
    }
    const int blocks = reduce_grid((long long)v.width * v.height);
    Scratch *sc = nullptr;
    int rc = get_scratch(&sc, blocks);
    if (rc) return rc;
    ReduceParams p{v.data, v.stride, v.width, v.height, v.offset_x, v.offset_y, sc->partials, sc->ticket, sc->result, mode};
    OpScope scope(s, "hb_reduce");
    switch (v.dtype) {
    case HB_U8: reduce_int_kernel<uchar><<<blocks, RT, 0, s>>>(p, sc->ipart); break;
    case HB_S8: reduce_int_kernel<signed char><<<blocks, RT, 0, s>>>(p, sc->ipart); break;
    case HB_S16: reduce_int_kernel<short><<<blocks, RT, 0, s>>>(p, sc->ipart); break;
    case HB_U16: reduce_int_kernel<unsigned short><<<blocks, RT, 0, s>>>(p, sc->ipart); break;
    case HB_S32: reduce_int_kernel<int><<<blocks, RT, 0, s>>>(p, sc->ipart); break;
    case HB_U32: reduce_int_kernel<unsigned int><<<blocks, RT, 0, s>>>(p, sc->ipart); break;
    default: log_msg(2, "hb_reduce: dtype %d unsupported", v.dtype); return HB_ERR_UNSUPPORTED;
    }
    g_launches++;
    rc = scope.finish();
    rc |= check_cuda(cudaMemcpyAsync(sc->host, sc->result, sizeof(long long), cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(result)");
    rc |= check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize()");
    const long long r = *reinterpret_cast<long long *>(sc->host);
    switch (v.dtype) {
    case HB_U8: *static_cast<uchar *>(result_host) = (uchar)r; break;
    case HB_S8: *static_cast<signed char *>(result_host) = (signed char)r; break;
    case HB_S16: *static_cast<short *>(result_host) = (short)r; break;
    case HB_U16: *static_cast<unsigned short *>(result_host) = (unsigned short)r; break;
    case HB_S32: *static_cast<int *>(result_host) = (int)r; break;
    default: *static_cast<unsigned int *>(result_host) = (unsigned int)r; break;
    }
    return rc ? HB_ERR_CUDA : HB_OK;
}
