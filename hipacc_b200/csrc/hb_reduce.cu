// hb_reduce.cu -- global reductions (Kernel::reduce()/reduced_data(), dsl/kernel.hpp:121-161) for sm_100a.
//
// Replaces hipaccApplyReductionShared + hipacc_shared_reduction (runtime/hipacc_cu.tpp:312-408,
// runtime/hipacc_cu_red.hpp:140-346: one pixel per thread, warp-synchronous shared-memory tree,
// 3 cudaMalloc/cudaFree + a blocking memcpy per call).
//
// One pass over HBM for every pixel type: 16 CTAs per SM stride over row chunks with 16-byte streaming loads
// (4 independent loads in flight per thread), each thread keeps private accumulators, then warp shuffles -> one
// partial per CTA -> the last CTA to finish (atomic ticket) folds the partials in a FIXED order with all its
// threads, so results are deterministic for a given grid.
//   float images  : min, max and sum come out of the same pass (fused); float sums are float per thread and double
//                   across threads (DESIGN.md "Float SUM"); PROD has its own kernel with the same accumulation rule.
//   integer images: the fold runs in the pixel type's modular arithmetic (data_t reduce(data_t, data_t)): SUM and
//                   PROD wrap modulo 2^bits and are therefore independent of the order -- bit-exact; byte and
//                   16-bit sums use the integer dot-product instructions (IDP.4A / IDP.2A) on whole words.
// Scratch (partials, ticket, result slot, pinned host mirror) is allocated ONCE per (device, stream) at its maximum
// size and never reallocated, so concurrent reductions on different streams do not share a ticket and a captured
// graph never refers to freed memory; hb_graph_begin reserves the scratch of its stream before the capture starts.
//
// NaN pixels: MIN / MAX ignore them (fminf / fmaxf); an image of NaNs only gives +inf / -inf.  The DSL's serial fold
// `reduce(result, pixel)` with min(a,b) = a < b ? a : b (dsl/math_functions.hpp:349-351) instead RESTARTS at the
// pixel after a NaN, and the reference's OpenMP runtime does that per thread chunk (runtime/hipacc_cpu_red.hpp:19-68),
// i.e. the reference has no thread-count independent result for such images (tests/test_gpu_parity.py::test_reduce_nan).
#include "hb_common.cuh"
#include "hb_internal.h"

#include <cfloat>
#include <climits>
#include <cstring>
#include <map>
#include <mutex>
#include <type_traits>

namespace hb {

struct MMS {  // min, max, sum partial
    float mn, mx;
    double sum;
};

struct ReduceParams {
    const void *in;
    int stride, w, h, ox, oy;
    void *partials;         // gridDim.x entries (MMS, or 8-byte slots for the single-op kernels)
    unsigned *ticket;       // zero before launch, reset by the last CTA
    void *result;           // device: MMS (f32 fused) or 8-byte scalar slot
    int mode;
};

__device__ __forceinline__ MMS mms_combine(MMS a, MMS b) {
    MMS r;
    r.mn = fminf(a.mn, b.mn);
    r.mx = fmaxf(a.mx, b.mx);
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
constexpr int kReduceCtasPerSm = 16;

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

// how a row of `w` pixels of `es` bytes splits into scalar head, 16-byte vectors and scalar tail
struct RowSplit {
    int head_n, nvec, tail0, cpr;
};
__host__ __device__ inline RowSplit row_split(const void *in, int stride, int ox, int w, int es) {
    const int per = 16 / es;
    const uintptr_t base_addr = reinterpret_cast<uintptr_t>(in) + (size_t)ox * es;
    const bool aligned_rows = (((size_t)stride * es) % 16 == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0);
    const int head = aligned_rows ? (int)(((16 - (base_addr & 15)) & 15) / es) : 0;
    RowSplit r;
    r.head_n = head < w ? head : w;
    r.nvec = aligned_rows ? (w - r.head_n) / per : 0;
    r.tail0 = r.head_n + r.nvec * per;
    r.cpr = r.nvec > 0 ? (r.nvec + CHUNK_V - 1) / CHUNK_V : 1;
    return r;
}

__global__ void __launch_bounds__(RT) reduce_mms_f32_kernel(const __grid_constant__ ReduceParams p) {
    const float *in = static_cast<const float *>(p.in);
    const float INF = __int_as_float(0x7f800000);
    float mn = INF, mx = -INF;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;  // 4 independent float chains per thread

    const RowSplit rs = row_split(in, p.stride, p.ox, p.w, 4);
    const long long total = (long long)rs.cpr * p.h;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int y = (int)(u / rs.cpr), c = (int)(u - (long long)y * rs.cpr);
        const float *row = in + (size_t)(p.oy + y) * p.stride + p.ox;
        const float4 *vrow = reinterpret_cast<const float4 *>(row + rs.head_n);
        const int v0 = c * CHUNK_V + threadIdx.x;
        float4 v[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const int i = v0 + k * RT;
            v[k] = i < rs.nvec ? __ldcs(vrow + i) : make_float4(INF, INF, INF, INF);
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            if (v0 + k * RT < rs.nvec) {
                mn = fminf(mn, fminf(fminf(v[k].x, v[k].y), fminf(v[k].z, v[k].w)));
                mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
                s0 += v[k].x; s1 += v[k].y; s2 += v[k].z; s3 += v[k].w;
            }
        }
        if (c == 0) {  // scalar head / tail pixels of this row
            const int nscal = rs.head_n + (p.w - rs.tail0);
            for (int k = threadIdx.x; k < nscal; k += RT) {
                const float e = row[k < rs.head_n ? k : rs.tail0 + (k - rs.head_n)];
                mn = fminf(mn, e); mx = fmaxf(mx, e); s0 += e;
            }
        }
    }
    MMS v{mn, mx, ((double)s0 + (double)s1) + ((double)s2 + (double)s3)};
    v = block_fold(v);

    MMS *partials = static_cast<MMS *>(p.partials);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v;
        __threadfence();
        const unsigned t = atomicAdd(p.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fixed-order fold of the per-CTA partials: thread t takes partials t, t+RT, ... then block fold
        MMS a{INF, -INF, 0.0};
        for (unsigned i = threadIdx.x; i < gridDim.x; i += RT) a = mms_combine(a, partials[i]);
        __syncthreads();
        a = block_fold(a);
        if (threadIdx.x == 0) {
            *static_cast<MMS *>(p.result) = a;
            *p.ticket = 0;  // ready for the next call
        }
    }
}

// ------------------------------------------------------------------------------------------------
// single-op kernels: one accumulator type A per pixel type, combine() is associative and commutative
// ------------------------------------------------------------------------------------------------
template <typename A, int MODE> __device__ __forceinline__ A op_identity() {
    if (MODE == HB_REDUCE_SUM) return (A)0;
    if (MODE == HB_REDUCE_PROD) return (A)1;
    if (std::is_same<A, double>::value) return MODE == HB_REDUCE_MIN ? (A)INFINITY : (A)-INFINITY;
    if (std::is_same<A, int>::value) return MODE == HB_REDUCE_MIN ? (A)INT_MAX : (A)INT_MIN;
    return MODE == HB_REDUCE_MIN ? (A)UINT_MAX : (A)0;
}
template <typename A, int MODE> __device__ __forceinline__ A op_combine(A a, A b) {
    if (MODE == HB_REDUCE_SUM) return a + b;
    if (MODE == HB_REDUCE_PROD) return a * b;
    if (MODE == HB_REDUCE_MIN) return b < a ? b : a;
    return b > a ? b : a;
}
template <typename A> __device__ __forceinline__ A shfl_down_any(A v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

template <typename A, int MODE>
__device__ __forceinline__ A block_fold_op(A v) {
    __shared__ A warp_part[RT / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = op_combine<A, MODE>(v, shfl_down_any(v, d));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < RT / 32 ? warp_part[lane] : op_identity<A, MODE>();
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = op_combine<A, MODE>(v, shfl_down_any(v, d));
    }
    return v;
}

// per-CTA partial -> ticket -> the last CTA folds all partials with every thread (fixed order), writes the 8-byte slot
template <typename A, int MODE>
__device__ __forceinline__ void grid_fold(const ReduceParams &p, A v) {
    v = block_fold_op<A, MODE>(v);
    A *partials = static_cast<A *>(p.partials);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v;
        __threadfence();
        is_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        A a = op_identity<A, MODE>();
        for (unsigned i = threadIdx.x; i < gridDim.x; i += RT) a = op_combine<A, MODE>(a, partials[i]);
        __syncthreads();
        a = block_fold_op<A, MODE>(a);
        if (threadIdx.x == 0) {
            *static_cast<A *>(p.result) = a;
            *p.ticket = 0;
        }
    }
}

// element k of a 16-byte vector as the accumulator type (sign- or zero-extended)
template <typename T, typename A> __device__ __forceinline__ A elem_of(const uint4 &v, int k) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    constexpr int per_word = 4 / (int)sizeof(T);
    const unsigned word = w[k / per_word];
    if (sizeof(T) == 4) return (A)(T)word;
    const int sh = (k % per_word) * 8 * (int)sizeof(T);
    return (A)(T)(word >> sh);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(RT) reduce_int_kernel(const __grid_constant__ ReduceParams p) {
    typedef typename std::conditional<std::is_signed<T>::value, int, unsigned>::type A;
    constexpr int N = 16 / (int)sizeof(T);
    const T *in = static_cast<const T *>(p.in);
    A acc[UNR];
#pragma unroll
    for (int k = 0; k < UNR; ++k) acc[k] = op_identity<A, MODE>();

    const RowSplit rs = row_split(in, p.stride, p.ox, p.w, (int)sizeof(T));
    const long long total = (long long)rs.cpr * p.h;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int y = (int)(u / rs.cpr), c = (int)(u - (long long)y * rs.cpr);
        const T *row = in + (size_t)(p.oy + y) * p.stride + p.ox;
        const uint4 *vrow = reinterpret_cast<const uint4 *>(row + rs.head_n);
        const int v0 = c * CHUNK_V + threadIdx.x;
        uint4 v[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k)
            if (v0 + k * RT < rs.nvec) v[k] = __ldcs(vrow + v0 + k * RT);
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            if (v0 + k * RT >= rs.nvec) continue;
            if (MODE == HB_REDUCE_SUM && sizeof(T) == 1) {          // IDP.4A: four bytes of a word per instruction
                const unsigned w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    acc[k] = std::is_signed<T>::value ? (A)__dp4a((int)w[q], 0x01010101, (int)acc[k]) : (A)__dp4a(w[q], 0x01010101u, (unsigned)acc[k]);
            } else if (MODE == HB_REDUCE_SUM && sizeof(T) == 2) {   // IDP.2A: both halves of a word (a.lo * b.byte0 + a.hi * b.byte1)
                const unsigned w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    acc[k] = std::is_signed<T>::value ? (A)__dp2a_lo((int)w[q], 0x0101, (int)acc[k]) : (A)__dp2a_lo(w[q], 0x0101u, (unsigned)acc[k]);
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) acc[k] = op_combine<A, MODE>(acc[k], elem_of<T, A>(v[k], i));
            }
        }
        if (c == 0) {
            const int nscal = rs.head_n + (p.w - rs.tail0);
            for (int k = threadIdx.x; k < nscal; k += RT)
                acc[0] = op_combine<A, MODE>(acc[0], (A)row[k < rs.head_n ? k : rs.tail0 + (k - rs.head_n)]);
        }
    }
    const A a = op_combine<A, MODE>(op_combine<A, MODE>(acc[0], acc[1]), op_combine<A, MODE>(acc[2], acc[3]));
    grid_fold<A, MODE>(p, a);
}

// float PROD (no sample uses it; the DSL allows any binary reduce()): float chains per thread, double across
__global__ void __launch_bounds__(RT) reduce_prod_f32_kernel(const __grid_constant__ ReduceParams p) {
    const float *in = static_cast<const float *>(p.in);
    float a0 = 1.0f, a1 = 1.0f, a2 = 1.0f, a3 = 1.0f;
    const RowSplit rs = row_split(in, p.stride, p.ox, p.w, 4);
    const long long total = (long long)rs.cpr * p.h;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int y = (int)(u / rs.cpr), c = (int)(u - (long long)y * rs.cpr);
        const float *row = in + (size_t)(p.oy + y) * p.stride + p.ox;
        const float4 *vrow = reinterpret_cast<const float4 *>(row + rs.head_n);
        const int v0 = c * CHUNK_V + threadIdx.x;
        float4 v[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) v[k] = v0 + k * RT < rs.nvec ? __ldcs(vrow + v0 + k * RT) : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
#pragma unroll
        for (int k = 0; k < UNR; ++k) { a0 *= v[k].x; a1 *= v[k].y; a2 *= v[k].z; a3 *= v[k].w; }
        if (c == 0) {
            const int nscal = rs.head_n + (p.w - rs.tail0);
            for (int k = threadIdx.x; k < nscal; k += RT) a0 *= row[k < rs.head_n ? k : rs.tail0 + (k - rs.head_n)];
        }
    }
    grid_fold<double, HB_REDUCE_PROD>(p, ((double)a0 * (double)a1) * ((double)a2 * (double)a3));
}

// ------------------------------------------------------------------------------------------------
// scratch: one set per (device, stream), allocated once at its maximum size, never reallocated
// ------------------------------------------------------------------------------------------------
struct Scratch {
    void *partials = nullptr;   // max_blocks * 16 bytes
    unsigned *ticket = nullptr;
    MMS *result = nullptr;      // 16 bytes: MMS or an 8-byte scalar slot
    MMS *host = nullptr;        // pinned mirror of `result`
};
static std::mutex g_scratch_mutex;
static std::map<std::pair<int, cudaStream_t>, Scratch> g_scratch;

static int max_blocks() { return sm_count() * kReduceCtasPerSm; }

static int get_scratch(Scratch **out, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    auto it = g_scratch.find({dev, s});
    if (it == g_scratch.end()) {
        HB_REQUIRE(!stream_is_capturing(s), HB_ERR_INVALID,
                   "reduction scratch for this stream does not exist yet and cannot be allocated inside a stream capture: "
                   "begin the capture with hb_graph_begin (it reserves the scratch) or run one reduction on the stream first");
        Scratch sc;
        int rc = check_cuda(cudaMalloc(&sc.partials, (size_t)max_blocks() * sizeof(MMS)), "cudaMalloc(reduce partials)");
        rc |= check_cuda(cudaMalloc(&sc.ticket, sizeof(unsigned)), "cudaMalloc(ticket)");
        rc |= check_cuda(cudaMalloc(&sc.result, sizeof(MMS)), "cudaMalloc(result)");
        rc |= check_cuda(cudaMallocHost(&sc.host, sizeof(MMS)), "cudaMallocHost(result)");
        if (!rc) rc = check_cuda(cudaMemset(sc.ticket, 0, sizeof(unsigned)), "cudaMemset(ticket)");   // synchronous: done before any launch
        if (rc) return rc;
        it = g_scratch.emplace(std::make_pair(dev, s), sc).first;
    }
    *out = &it->second;   // std::map nodes are stable
    return HB_OK;
}

// called by hb_graph_begin before the capture starts
int reserve_reduce_scratch(cudaStream_t s) {
    Scratch *sc = nullptr;
    return get_scratch(&sc, s);
}

static int grid_for(const hb_view &v, int es) {
    const RowSplit rs = row_split(v.data, v.stride, v.offset_x, v.width, es);
    const long long chunks = (long long)rs.cpr * v.height;
    long long blocks = stream_grid(chunks, kReduceCtasPerSm);   // measured best of 8 / 16 / 32 / 64 / one-shot (tools/stream_grid_sweep.sh)
    if (blocks > max_blocks()) blocks = max_blocks();
    return (int)blocks;
}

static int launch_mms(const hb_view &v, void *result_dev, Scratch *sc, cudaStream_t s) {
    const int blocks = grid_for(v, 4);
    ReduceParams p{v.data, v.stride, v.width, v.height, v.offset_x, v.offset_y, sc->partials, sc->ticket, result_dev ? result_dev : sc->result, 0};
    reduce_mms_f32_kernel<<<blocks, RT, 0, s>>>(p);
    g_launches++;
    return HB_OK;
}

template <typename T>
static void launch_int(const ReduceParams &p, int blocks, cudaStream_t s) {
    switch (p.mode) {
    case HB_REDUCE_SUM: reduce_int_kernel<T, HB_REDUCE_SUM><<<blocks, RT, 0, s>>>(p); break;
    case HB_REDUCE_MIN: reduce_int_kernel<T, HB_REDUCE_MIN><<<blocks, RT, 0, s>>>(p); break;
    case HB_REDUCE_MAX: reduce_int_kernel<T, HB_REDUCE_MAX><<<blocks, RT, 0, s>>>(p); break;
    default: reduce_int_kernel<T, HB_REDUCE_PROD><<<blocks, RT, 0, s>>>(p); break;
    }
}

// integer images (any mode) and float PROD: one launch, the scalar is left at `result_dev` (32-bit accumulator for integer
// pixels, double for float PROD)
static int launch_generic(const hb_view &v, int mode, void *result_dev, Scratch *sc, cudaStream_t s, const char *who) {
    HB_REQUIRE(!is_x4(v.dtype), HB_ERR_UNSUPPORTED, "%s: vector pixels have no device reduction; no CPU fallback", who);
    const int blocks = grid_for(v, dtype_size(v.dtype));
    ReduceParams p{v.data, v.stride, v.width, v.height, v.offset_x, v.offset_y, sc->partials, sc->ticket, result_dev, mode};
    switch (v.dtype) {
    case HB_U8: launch_int<uchar>(p, blocks, s); break;
    case HB_S8: launch_int<signed char>(p, blocks, s); break;
    case HB_S16: launch_int<short>(p, blocks, s); break;
    case HB_U16: launch_int<unsigned short>(p, blocks, s); break;
    case HB_S32: launch_int<int>(p, blocks, s); break;
    case HB_U32: launch_int<unsigned int>(p, blocks, s); break;
    case HB_F32: reduce_prod_f32_kernel<<<blocks, RT, 0, s>>>(p); break;
    default: log_msg(2, "%s: dtype %d unsupported", who, v.dtype); return HB_ERR_UNSUPPORTED;
    }
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
    Scratch *sc = nullptr;
    int rc = get_scratch(&sc, s);
    if (rc) return rc;
    OpScope scope(s, "hb_reduce_minmaxsum_f32_async");
    rc = launch_mms(v, partials_device, sc, s);
    if (rc) return rc;
    return scope.finish();
}

extern "C" int hb_reduce_minmaxsum_f32(const hb_view *in_, float result_host[3], void *stream) {
    HB_REQUIRE(in_ && result_host, HB_ERR_INVALID, "hb_reduce_minmaxsum_f32: null argument");
    hb_view v = norm_view(*in_);
    HB_REQUIRE(view_ok(v) && v.dtype == HB_F32, HB_ERR_INVALID, "hb_reduce_minmaxsum_f32: needs a valid f32 view");
    cudaStream_t s = (cudaStream_t)stream;
    HB_REQUIRE(!stream_is_capturing(s), HB_ERR_INVALID, "hb_reduce_minmaxsum_f32 blocks and cannot be captured; use hb_reduce_minmaxsum_f32_async");
    Scratch *sc = nullptr;
    int rc = get_scratch(&sc, s);
    if (rc) return rc;
    OpScope scope(s, "hb_reduce_minmaxsum_f32");
    rc = launch_mms(v, nullptr, sc, s);
    if (rc) return rc;
    rc = scope.finish();
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
    if (v.dtype == HB_F32 && mode != HB_REDUCE_PROD) {
        float r[3];
        int rc = hb_reduce_minmaxsum_f32(&v, r, stream);
        *static_cast<float *>(result_host) = mode == HB_REDUCE_MIN ? r[0] : mode == HB_REDUCE_MAX ? r[1] : r[2];
        return rc;
    }
    HB_REQUIRE(!stream_is_capturing(s), HB_ERR_INVALID, "hb_reduce blocks and cannot be captured; use hb_reduce_async");
    Scratch *sc = nullptr;
    int rc = get_scratch(&sc, s);
    if (rc) return rc;
    OpScope scope(s, "hb_reduce");
    rc = launch_generic(v, mode, sc->result, sc, s, "hb_reduce");
    if (rc) return rc;
    rc = scope.finish();
    rc |= check_cuda(cudaMemcpyAsync(sc->host, sc->result, sizeof(MMS), cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(result)");
    rc |= check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize()");
    const unsigned r = *reinterpret_cast<const unsigned *>(sc->host);   // integer accumulators are 32-bit
    switch (v.dtype) {
    case HB_U8: *static_cast<uchar *>(result_host) = (uchar)r; break;
    case HB_S8: *static_cast<signed char *>(result_host) = (signed char)r; break;
    case HB_S16: *static_cast<short *>(result_host) = (short)r; break;
    case HB_U16: *static_cast<unsigned short *>(result_host) = (unsigned short)r; break;
    case HB_S32: *static_cast<int *>(result_host) = (int)r; break;
    case HB_U32: *static_cast<unsigned int *>(result_host) = r; break;
    default: *static_cast<float *>(result_host) = (float)*reinterpret_cast<const double *>(sc->host); break;
    }
    return rc ? HB_ERR_CUDA : HB_OK;
}

extern "C" int hb_reduce_async(const hb_view *in_, int mode, void *result_device, void *stream) {
    HB_REQUIRE(in_ && result_device, HB_ERR_INVALID, "hb_reduce_async: null argument");
    HB_REQUIRE(mode >= HB_REDUCE_SUM && mode <= HB_REDUCE_PROD, HB_ERR_INVALID, "hb_reduce_async: bad mode");
    hb_view v = norm_view(*in_);
    HB_REQUIRE(view_ok(v), HB_ERR_INVALID, "hb_reduce_async: malformed view");
    HB_REQUIRE(!(v.dtype == HB_F32 && mode != HB_REDUCE_PROD), HB_ERR_UNSUPPORTED,
               "hb_reduce_async: float MIN / MAX / SUM come fused from hb_reduce_minmaxsum_f32_async");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch *sc = nullptr;
    int rc = get_scratch(&sc, s);
    if (rc) return rc;
    OpScope scope(s, "hb_reduce_async");
    rc = launch_generic(v, mode, result_device, sc, s, "hb_reduce_async");
    if (rc) return rc;
    return scope.finish();
}
