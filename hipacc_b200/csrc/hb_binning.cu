// hb_binning.cu -- binning / histograms (Kernel::binning() + binned_data(), dsl/kernel.hpp:163-209) for sm_100a.
//
// Replaces hipaccApplyBinningSegmented and the generated segmented-binning kernel
// (runtime/hipacc_cu.tpp:410-464, runtime/hipacc_cu_red.hpp:527-641: per-warp hand-rolled locks in
// shared memory, a second merge kernel, cudaMalloc/cudaFree per call).
//
// One pass over HBM (4 B per float pixel, 1 B per uchar pixel): 16 CTAs per SM walk the row chunks with
// 16-byte streaming loads (4 independent loads in flight per thread); every warp owns a private copy of
// the bins in shared memory (native shared-memory atomics, no cross-warp contention); at the end the
// CTA folds its copies and adds the non-zero bins to the result with global atomics.  Integer addition
// commutes, so the result does not depend on the schedule.
#include "hb_common.cuh"
#include "hb_internal.h"

#include <cmath>
#include <cstring>
#include <mutex>

namespace hb {

struct BinParams {
    const void *in;
    int stride, w, h, ox, oy;
    unsigned *bins;
    int num_bins, copies;  // copies of the bins in shared memory (0: global atomics only)
    int index_kind, value_kind;
    float p0, rp0, nbf;  // divisor, its refined reciprocal, (float)num_bins
};

constexpr int HT = 256, HU = 4;

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <typename T> struct BinVec;
template <> struct BinVec<float> { typedef float4 V; static constexpr int N = 4; };
template <> struct BinVec<uchar> { typedef uint4 V; static constexpr int N = 16; };

__device__ __forceinline__ void unpack(const float4 &v, float (&e)[4]) { e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w; }
__device__ __forceinline__ void unpack(const uint4 &v, uchar (&e)[16]) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = (uchar)(w[i >> 2] >> (8 * (i & 3)));
}

// Bin index of the float value `v` = (uint)v.  The C conversion is defined for v in (-1, 2^32); the values that
// truncate into [0, num_bins) are v in (-1, num_bins) -- everything else is dropped (the emitted Put helper drops
// idx >= num_bins; v <= -1, +-inf and NaN are undefined in C: x86-64 wraps them to a huge index or to 0, the
// reference's CUDA backend saturates -- this library drops them).  Branch-free: truncation is an RZ-add of 2^23
// that leaves the integer part in the mantissa (num_bins <= 2^22), no conversion-pipe instruction.
__device__ __forceinline__ bool f2bin(float v, float nbf, unsigned &idx) {
    const float t = __fadd_rz(fmaxf(v, 0.0f), 8388608.0f);
    idx = (unsigned)(__float_as_int(t) - 0x4B000000);
    return v < nbf && v > -1.0f;   // false for NaN
}

// pixel / p0 * num_bins, every operation correctly rounded.  FASTDIV: a / b through the host-computed correctly
// rounded reciprocal r = RN(1/b): q = RN(a*r), e = RN(a - b*q) (exact), q' = RN(q + e*r) is the correctly rounded
// quotient (Markstein; the host enables it only for finite b in [1e-15, 1e15] whose mantissa is not all ones) --
// the FMA tail of div.rn.f32 without its MUFU.RCP.  |a| >= 1e30 cannot land in a bin (and could overflow the tail).
template <bool FASTDIV>
__device__ __forceinline__ bool scaled_index(const BinParams &p, float a, unsigned &idx) {
    float q;
    if (FASTDIV) {
        const float q0 = __fmul_rn(a, p.rp0);
        q = __fmaf_rn(__fmaf_rn(-p.p0, q0, a), p.rp0, q0);
    } else {
        q = __fdiv_rn(a, p.p0);
    }
    const float v = fabsf(a) < 1e30f ? __fmul_rn(q, p.nbf) : -2.0f;
    return f2bin(v, p.nbf, idx);
}

// Shared-memory path: every pixel issues exactly ONE red.shared.add and no branch.  A pixel that falls into no bin
// (v <= -1, v >= num_bins, +-inf, NaN) adds to a spare word behind the copy's bins (index num_bins) that the final fold
// ignores.  The |a| < 1e30 guard of scaled_index() is not needed here: such a pixel's quotient is >= 1e15 in magnitude,
// inf or NaN, all of which land in the spare word anyway.
__device__ __forceinline__ unsigned f2bin_spare(float v, float nbf) {
    float w = v > -1.0f ? v : nbf;            // v <= -1 and NaN -> spare
    w = fminf(fmaxf(w, 0.0f), nbf);           // (-1, 0) -> bin 0 like the C conversion; v >= num_bins, +inf -> spare
    return (unsigned)__float_as_int(__fadd_rz(w, 8388608.0f));   // 0x4B000000 + index: the caller's base address carries -4 * 0x4B000000
}
constexpr unsigned kF2BinBias = 0x4B000000u;
template <bool FASTDIV>
__device__ __forceinline__ unsigned scaled_index_spare(const BinParams &p, float a) {
    float q;
    if (FASTDIV) {
        const float q0 = __fmul_rn(a, p.rp0);
        q = __fmaf_rn(__fmaf_rn(-p.p0, q0, a), p.rp0, q0);
    } else {
        q = __fdiv_rn(a, p.p0);
    }
    return f2bin_spare(__fmul_rn(q, p.nbf), p.nbf);
}
// sh32 = shared-space byte address of the warp's copy of the bins
template <typename T, int INDEX, int VALUE, bool FASTDIV>
__device__ __forceinline__ void bin_put_shared(const BinParams &p, unsigned sh32, T e) {
    unsigned addr;   // 32-bit arithmetic wraps: (sh32 - 4 * bias) + 4 * (bias + index) = sh32 + 4 * index
    if (INDEX == HB_BIN_INDEX_SCALE) addr = (sh32 - 4u * kF2BinBias) + 4u * scaled_index_spare<FASTDIV>(p, (float)e);
    else if (DtypeOf<T>::v == HB_F32) addr = (sh32 - 4u * kF2BinBias) + 4u * f2bin_spare((float)e, p.nbf);
    else addr = sh32 + 4u * min((unsigned)e, (unsigned)p.num_bins);
    if (VALUE == HB_BIN_VALUE_ONE) {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
    } else {
        const unsigned val = DtypeOf<T>::v == HB_F32 ? __float2uint_rz((float)e) : (unsigned)e;
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(val) : "memory");
    }
}

template <typename T, int INDEX, int VALUE, bool FASTDIV>
__device__ __forceinline__ void bin_put(const BinParams &p, unsigned *sh, T e) {
    unsigned idx;
    bool ok;
    if (INDEX == HB_BIN_INDEX_SCALE) {
        ok = scaled_index<FASTDIV>(p, (float)e, idx);
    } else if (DtypeOf<T>::v == HB_F32) {
        ok = f2bin((float)e, p.nbf, idx);
    } else {
        idx = (unsigned)e;
        ok = idx < (unsigned)p.num_bins;
    }
    // value = (uint)pixel: float pixels outside [0, 2^32) are undefined in C; they saturate here
    const unsigned val = VALUE == HB_BIN_VALUE_ONE ? 1u : DtypeOf<T>::v == HB_F32 ? __float2uint_rz((float)e) : (unsigned)e;
    unsigned *dst = (sh ? sh : p.bins) + idx;
    if (ok) atomicAdd(dst, val);
}

template <typename T, int INDEX, int VALUE, bool FASTDIV, bool SH>
__global__ void __launch_bounds__(HT) binning_kernel(const __grid_constant__ BinParams p) {
    typedef typename BinVec<T>::V V;
    constexpr int N = BinVec<T>::N;
    extern __shared__ unsigned hsh[];
    const int bstride = p.num_bins + 1;   // the bins of one copy + its spare word
    if (SH) {
        for (int i = threadIdx.x; i < p.copies * bstride; i += HT) hsh[i] = 0u;
        __syncthreads();
    }
    const unsigned my32 = SH ? smem_addr(hsh + ((threadIdx.x >> 5) % p.copies) * bstride) : 0u;
    auto put = [&](T e) {
        if (SH) bin_put_shared<T, INDEX, VALUE, FASTDIV>(p, my32, e);
        else bin_put<T, INDEX, VALUE, FASTDIV>(p, nullptr, e);
    };

    const T *in = static_cast<const T *>(p.in);
    const uintptr_t base_addr = reinterpret_cast<uintptr_t>(in) + (size_t)p.ox * sizeof(T);
    const bool aligned_rows = ((size_t)p.stride * sizeof(T)) % 16 == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0;
    const int head = aligned_rows ? (int)(((16 - (base_addr & 15)) & 15) / sizeof(T)) : 0;  // scalar pixels before alignment
    const int head_n = head < p.w ? head : p.w;
    const int nvec = aligned_rows ? (p.w - head_n) / N : 0;
    const int tail0 = head_n + nvec * N;
    const int cpr = nvec > 0 ? (nvec + HT * HU - 1) / (HT * HU) : 1;
    const long long total = (long long)cpr * p.h;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int y = (int)(u / cpr), c = (int)(u - (long long)y * cpr);
        const T *row = in + (size_t)(p.oy + y) * p.stride + p.ox;
        const V *vrow = reinterpret_cast<const V *>(row + head_n);
        const int v0 = c * (HT * HU) + threadIdx.x;
        V v[HU];
#pragma unroll
        for (int k = 0; k < HU; ++k)
            if (v0 + k * HT < nvec) v[k] = __ldcs(vrow + v0 + k * HT);
#pragma unroll
        for (int k = 0; k < HU; ++k)
            if (v0 + k * HT < nvec) {
                T e[N];
                unpack(v[k], e);
#pragma unroll
                for (int i = 0; i < N; ++i) put(e[i]);
            }
        if (c == 0) {  // scalar head / tail pixels of this row
            const int nscal = head_n + (p.w - tail0);
            for (int k = threadIdx.x; k < nscal; k += HT) put(row[k < head_n ? k : tail0 + (k - head_n)]);
        }
    }
    if (SH) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.num_bins; i += HT) {
            unsigned a = 0;
            for (int k = 0; k < p.copies; ++k) a += hsh[k * bstride + i];
            if (a) atomicAdd(p.bins + i, a);
        }
    }
}

template <typename T, int INDEX, int VALUE>
static void launch_binning_kernel(const BinParams &p, bool fastdiv, int blocks, size_t smem, cudaStream_t s) {
    if (p.copies) {
        if (INDEX == HB_BIN_INDEX_SCALE && fastdiv) binning_kernel<T, INDEX, VALUE, true, true><<<blocks, HT, smem, s>>>(p);
        else binning_kernel<T, INDEX, VALUE, false, true><<<blocks, HT, smem, s>>>(p);
    } else {   // more bins than shared memory holds: global atomics
        if (INDEX == HB_BIN_INDEX_SCALE && fastdiv) binning_kernel<T, INDEX, VALUE, true, false><<<blocks, HT, 0, s>>>(p);
        else binning_kernel<T, INDEX, VALUE, false, false><<<blocks, HT, 0, s>>>(p);
    }
}
template <typename T>
static void dispatch_binning(const BinParams &p, bool fastdiv, int blocks, size_t smem, cudaStream_t s) {
    if (p.index_kind == HB_BIN_INDEX_SCALE) {
        if (p.value_kind == HB_BIN_VALUE_ONE) launch_binning_kernel<T, HB_BIN_INDEX_SCALE, HB_BIN_VALUE_ONE>(p, fastdiv, blocks, smem, s);
        else launch_binning_kernel<T, HB_BIN_INDEX_SCALE, HB_BIN_VALUE_PIXEL>(p, fastdiv, blocks, smem, s);
    } else {
        if (p.value_kind == HB_BIN_VALUE_ONE) launch_binning_kernel<T, HB_BIN_INDEX_PIXEL, HB_BIN_VALUE_ONE>(p, fastdiv, blocks, smem, s);
        else launch_binning_kernel<T, HB_BIN_INDEX_PIXEL, HB_BIN_VALUE_PIXEL>(p, fastdiv, blocks, smem, s);
    }
}

struct BinScratch {
    unsigned *dev = nullptr, *host = nullptr;
    int cap = 0;
};
static BinScratch g_bin[16];
static std::mutex g_bin_mutex;   // the blocking form owns the per-device scratch for the whole call

static int launch_binning(const hb_binning_desc *d, unsigned *bins_dev, cudaStream_t s, const char *who) {
    hb_view v = norm_view(d->in);
    HB_REQUIRE(view_ok(v) && (v.dtype == HB_F32 || v.dtype == HB_U8), HB_ERR_UNSUPPORTED, "%s: needs a valid f32 or u8 view; no CPU fallback", who);
    HB_REQUIRE(d->num_bins > 0 && d->num_bins <= (1 << 22), HB_ERR_INVALID, "%s: num_bins %d out of range", who, d->num_bins);
    HB_REQUIRE(d->index_kind == HB_BIN_INDEX_SCALE || d->index_kind == HB_BIN_INDEX_PIXEL, HB_ERR_INVALID, "%s: bad index kind", who);
    HB_REQUIRE(d->value_kind == HB_BIN_VALUE_ONE || d->value_kind == HB_BIN_VALUE_PIXEL, HB_ERR_INVALID, "%s: bad value kind", who);
    HB_REQUIRE(d->index_kind != HB_BIN_INDEX_SCALE || d->p0 != 0.0, HB_ERR_INVALID, "%s: p0 == 0", who);
    BinParams p;
    memset(&p, 0, sizeof(p));
    p.in = v.data; p.stride = v.stride; p.w = v.width; p.h = v.height; p.ox = v.offset_x; p.oy = v.offset_y;
    p.bins = bins_dev; p.num_bins = d->num_bins; p.index_kind = d->index_kind; p.value_kind = d->value_kind;
    p.p0 = (float)d->p0; p.nbf = (float)(unsigned)d->num_bins;
    bool fastdiv = false;
    if (d->index_kind == HB_BIN_INDEX_SCALE) {   // r = RN(1/p0), see scaled_index()
        const float b = p.p0;
        unsigned bits;
        memcpy(&bits, &b, sizeof(bits));
        fastdiv = std::isfinite(b) && std::fabs(b) >= 1e-15f && std::fabs(b) <= 1e15f && (bits & 0x7FFFFFu) != 0x7FFFFFu;
        p.rp0 = fastdiv ? 1.0f / b : 0.0f;
    }
    const int max_words = 48 * 1024 / 4;
    const int per_copy = d->num_bins + 1;   // + the spare word that takes the pixels of no bin
    p.copies = per_copy > max_words ? 0 : (max_words / per_copy < HT / 32 ? max_words / per_copy : HT / 32);
    const size_t smem = (size_t)p.copies * per_copy * sizeof(unsigned);
    const int npv = v.dtype == HB_F32 ? 4 : 16;
    const long long cpr = (v.width / npv + HT * HU - 1) / (HT * HU);
    const long long chunks = (cpr < 1 ? 1 : cpr) * v.height;
    const int blocks = (int)stream_grid(chunks, smem > 16 * 1024 ? 4 : 16);
    int rc = check_cuda(cudaMemsetAsync(bins_dev, 0, sizeof(unsigned) * d->num_bins, s), "cudaMemsetAsync(bins)");
    if (rc) return rc;
    if (v.dtype == HB_F32) dispatch_binning<float>(p, fastdiv, blocks, smem, s);
    else dispatch_binning<uchar>(p, fastdiv, blocks, smem, s);
    g_launches++;
    return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" int hb_binning_async(const hb_binning_desc *d, uint32_t *bins_device, void *stream) {
    HB_REQUIRE(d && bins_device, HB_ERR_INVALID, "hb_binning_async: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_binning_async");
    int rc = launch_binning(d, bins_device, s, "hb_binning_async");
    if (rc) return rc;
    return scope.finish();
}

extern "C" int hb_binning(const hb_binning_desc *d, uint32_t *bins_host, void *stream) {
    HB_REQUIRE(d && bins_host, HB_ERR_INVALID, "hb_binning: null argument");
    HB_REQUIRE(d->num_bins > 0 && d->num_bins <= (1 << 22), HB_ERR_INVALID, "hb_binning: num_bins %d out of range", d->num_bins);
    HB_REQUIRE(!stream_is_capturing((cudaStream_t)stream), HB_ERR_INVALID, "hb_binning blocks and cannot be captured; use hb_binning_async");
    std::lock_guard<std::mutex> lock(g_bin_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    BinScratch &sc = g_bin[dev & 15];
    if (sc.cap < d->num_bins) {
        if (sc.dev) { cudaFree(sc.dev); cudaFreeHost(sc.host); sc.dev = nullptr; sc.host = nullptr; sc.cap = 0; }
        int rc = check_cuda(cudaMalloc(&sc.dev, sizeof(unsigned) * d->num_bins), "cudaMalloc(bins)");
        rc |= check_cuda(cudaMallocHost(&sc.host, sizeof(unsigned) * d->num_bins), "cudaMallocHost(bins)");
        if (rc) return rc;
        sc.cap = d->num_bins;
    }
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_binning");
    int rc = launch_binning(d, sc.dev, s, "hb_binning");
    if (rc) return rc;
    rc = scope.finish();
    rc |= check_cuda(cudaMemcpyAsync(sc.host, sc.dev, sizeof(unsigned) * d->num_bins, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(bins)");
    rc |= check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize()");  // blocking like the reference
    memcpy(bins_host, sc.host, sizeof(unsigned) * d->num_bins);
    return rc ? HB_ERR_CUDA : HB_OK;
}
