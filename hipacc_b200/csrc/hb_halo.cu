// hb_halo.cu -- peer-to-peer halo exchange for row-strip sharding over NVLink / NVSwitch (sm_100a).
//
// The reference has no multi-device path; BASELINE.json asks for row strips across the
// GPUs of one box with halo rows exchanged peer to peer.  One process per GPU: every rank exports its strip buffer
// and a small control block through CUDA IPC, its neighbours map both, and ONE kernel launch per rank and exchange
//   1. tells the neighbours that the ghost rows they pushed last time have been consumed (stream order: all kernels
//      that read them were launched before this one),
//   2. waits until the neighbours have consumed what this rank pushed last time,
//   3. PUSHES its R top / bottom owned rows into the neighbours' ghost rows with 16-byte peer stores,
//   4. publishes them (system-scope fence + release store of the exchange number into the neighbour's control
//      block) and waits for the neighbours' rows to arrive in its own ghost rows.
// No host involvement, no NCCL launch, no staging copy: the exchange number lives in device memory, so the launch
// is CUDA-graph replayable.  A bounded spin (about two seconds) raises an error flag instead of hanging the GPU if
// a neighbour never shows up.
#include "hb_common.cuh"
#include "hb_internal.h"

#include <cuda.h>   // CUdeviceptr / CUresult types only
#include <cstring>

namespace hb {

// control block (ints): [0] exchanges completed, [1] data from the upper neighbour, [2] data from the lower,
//                       [3] ack from the upper, [4] ack from the lower, [5] timeout flag
constexpr int CTRL_INTS = 8;

struct HaloParams {
    unsigned char *buf;      // this rank's strip buffer (row 0 = first ghost row)
    size_t pitch, row_bytes;
    int gt, rows, R;
    unsigned char *up_dst;   // where my top R owned rows go: the upper neighbour's bottom ghost rows (peer memory)
    unsigned char *down_dst; // where my bottom R owned rows go: the lower neighbour's top ghost rows
    size_t up_pitch, down_pitch;
    int *ctrl, *up_ctrl, *down_ctrl;
};

__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool spin_until(int *flag, int want) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < want) {
        if (clock64() - t0 > 4000000000LL) return false;  // ~2 s at 1.9 GHz
    }
    return true;
}

// copy `rows` rows of `row_bytes` bytes into peer memory.  The CTAs of one exchange share the work: CTA `part` of
// `nparts` takes every nparts-th chunk of 16-byte vectors; each thread has 8 independent loads in flight before its
// first remote store (a dependent load -> store loop moves ~16 KB per microsecond and CTA, NVLink wants megabytes in flight).
__device__ __forceinline__ void push_rows(unsigned char *dst, size_t dpitch, const unsigned char *src, size_t spitch, size_t row_bytes, int rows,
                                          int part, int nparts) {
    if ((row_bytes | dpitch | spitch | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) % 16 == 0) {
        // flattened 32-bit vector index (a 64-bit division per vector would cost more than the copy)
        constexpr int U = 8;
        const unsigned vpr = (unsigned)(row_bytes / 16), total = vpr * (unsigned)rows;
        const unsigned step = blockDim.x * U;
        for (unsigned base = (unsigned)part * step; base < total; base += step * (unsigned)nparts) {
            uint4 v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const unsigned i = base + k * blockDim.x + threadIdx.x;
                if (i < total) {
                    const unsigned r = i / vpr, c = i - r * vpr;
                    v[k] = __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)r * spitch) + c);
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const unsigned i = base + k * blockDim.x + threadIdx.x;
                if (i < total) {
                    const unsigned r = i / vpr, c = i - r * vpr;
                    reinterpret_cast<uint4 *>(dst + (size_t)r * dpitch)[c] = v[k];
                }
            }
        }
    } else {
        const size_t total = row_bytes * rows;
        for (size_t i = (size_t)part * blockDim.x + threadIdx.x; i < total; i += (size_t)blockDim.x * nparts) {
            const size_t r = i / row_bytes, c = i - r * row_bytes;
            dst[r * dpitch + c] = src[r * spitch + c];
        }
    }
}

constexpr int kMaxHalo = 4;
struct HaloBatch {
    HaloParams p[kMaxHalo];
};

// one CTA per strip buffer of the batch (e.g. the Gaussian and the Laplacian level of a pyramid transition): the
// exchanges of a batch run concurrently in one launch
// A wait that times out (a neighbour never arrived) must not let the exchange carry on: pushing into ghost rows the
// neighbour may still be reading, or running the next operator on ghost rows that never arrived, would corrupt results
// silently.  The kernel then raises ctrl[5], skips the push / publish steps and POISONS the exchange counter (negative),
// so every later exchange on this control block returns immediately and hb_halo_status reports the failure.
constexpr int kPoisoned = -(1 << 30);
constexpr int HALO_THREADS = 512;

// grid = (buffers of the batch, parts): the `parts` CTAs of a buffer all wait for the neighbours' acks (read-only
// polling), push their share of the rows, and the LAST one to finish (ticket in ctrl[6]) publishes the data, waits for
// the neighbours' rows and advances the exchange counter -- every CTA has read the counter before that can happen.
__global__ void __launch_bounds__(HALO_THREADS) halo_exchange_kernel(const __grid_constant__ HaloBatch batch) {
    const HaloParams &p = batch.p[blockIdx.x];
    const int part = blockIdx.y, nparts = gridDim.y;
    __shared__ int ok;
    const int e = p.ctrl[0];
    if (e < 0) return;         // poisoned by an earlier timeout
    if (threadIdx.x == 0) {
        if (part == 0) {       // my ghost rows of exchange e-1 are consumed: the neighbours may overwrite them
            if (p.up_ctrl) st_release_sys(p.up_ctrl + 4, e);      // I am the upper neighbour's lower neighbour
            if (p.down_ctrl) st_release_sys(p.down_ctrl + 3, e);
        }
        bool good = true;
        if (p.up_ctrl) good = spin_until(p.ctrl + 3, e) && good;
        if (p.down_ctrl) good = spin_until(p.ctrl + 4, e) && good;
        ok = good ? 1 : 0;
    }
    __syncthreads();
    if (ok) {
        if (p.up_dst) push_rows(p.up_dst, p.up_pitch, p.buf + (size_t)p.gt * p.pitch, p.pitch, p.row_bytes, p.R, part, nparts);
        if (p.down_dst) push_rows(p.down_dst, p.down_pitch, p.buf + (size_t)(p.gt + p.rows - p.R) * p.pitch, p.pitch, p.row_bytes, p.R, part, nparts);
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!ok) atomicExch(p.ctrl + 5, 1);
        __threadfence();
        if (atomicAdd(reinterpret_cast<unsigned *>(p.ctrl + 6), 1u) != (unsigned)nparts - 1) return;   // not the last CTA of this buffer
        p.ctrl[6] = 0;
        __threadfence();
        bool good = atomicAdd(p.ctrl + 5, 0) == 0;
        if (good) {
            if (p.up_ctrl) st_release_sys(p.up_ctrl + 2, e + 1);   // data from its lower neighbour
            if (p.down_ctrl) st_release_sys(p.down_ctrl + 1, e + 1);
            if (p.up_ctrl) good = spin_until(p.ctrl + 1, e + 1) && good;
            if (p.down_ctrl) good = spin_until(p.ctrl + 2, e + 1) && good;
        }
        if (good) {
            p.ctrl[0] = e + 1;
        } else {
            p.ctrl[5] = 1;
            p.ctrl[0] = kPoisoned;
        }
        __threadfence();
    }
}

// ------------------------------------------------------------------------------------------------
// All-gather of row strips over peer memory: every rank owns rows [row0, row0 + rows) of an image that all ranks keep
// in FULL (the coarse pyramid levels, small enough to replicate) and pushes them into the same rows of every peer's
// copy.  One CTA per peer: CTA j handshakes with peer j alone (ack: "your rows of the previous gather are consumed",
// push, publish, wait for its rows), so the peers proceed independently; the last CTA to finish advances the
// gather counter.  Control block (ints): [0] gathers completed, [1] ticket, [5] timeout flag,
// [8 + slot] ack from rank `slot`, [24 + slot] data from rank `slot`, [40 + j] ticket of the CTAs that serve peer j.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 15;
constexpr int GATHER_CTRL_INTS = 64;   // [40 + j]: per-peer ticket
struct GatherParams {
    unsigned char *buf;
    size_t pitch, row_bytes;
    int row0, rows, my_slot, n_peers;
    int *ctrl;
    unsigned char *peer_buf[kMaxPeers];
    int *peer_ctrl[kMaxPeers];
    int peer_slot[kMaxPeers];
};

__global__ void __launch_bounds__(HALO_THREADS) allgather_rows_kernel(const __grid_constant__ GatherParams p) {
    const int j = blockIdx.x, part = blockIdx.y, nparts = gridDim.y;
    __shared__ int ok;
    const int e = p.ctrl[0];
    if (e < 0) return;   // poisoned by an earlier timeout
    int *pc = p.peer_ctrl[j];
    const int ps = p.peer_slot[j];
    if (threadIdx.x == 0) {
        if (part == 0) st_release_sys(pc + 8 + p.my_slot, e);   // what peer j pushed into my copy last time is consumed
        ok = spin_until(p.ctrl + 8 + ps, e) ? 1 : 0;            // and peer j has consumed what I pushed
    }
    __syncthreads();
    if (ok) {
        const size_t off = (size_t)p.row0 * p.pitch;
        push_rows(p.peer_buf[j] + off, p.pitch, p.buf + off, p.pitch, p.row_bytes, p.rows, part, nparts);
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!ok) atomicExch(p.ctrl + 5, 1);
        __threadfence();
        // per-peer ticket: the last CTA of peer j publishes to it and waits for its rows
        if (atomicAdd(reinterpret_cast<unsigned *>(p.ctrl + 40 + j), 1u) != (unsigned)nparts - 1) return;
        p.ctrl[40 + j] = 0;
        if (ok) {
            st_release_sys(pc + 24 + p.my_slot, e + 1);
            if (!spin_until(p.ctrl + 24 + ps, e + 1)) atomicExch(p.ctrl + 5, 1);
        }
        __threadfence();
        // global ticket over the peers: every CTA of the launch has read `e` and finished
        if (atomicAdd(reinterpret_cast<unsigned *>(p.ctrl + 1), 1u) == gridDim.x - 1) {
            p.ctrl[1] = 0;
            p.ctrl[0] = atomicAdd(p.ctrl + 5, 0) != 0 ? kPoisoned : e + 1;
            __threadfence();
        }
    }
}

}  // namespace hb

using namespace hb;

// base address of the allocation that contains `p` (cuMemGetAddressRange, resolved through cudart: no -lcuda)
static bool allocation_base(const void *p, const void **base) {
    typedef CUresult (*Fn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static Fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuMemGetAddressRange", &sym, 12000, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<Fn>(sym);
        else
            cudaGetLastError();
    }
    if (!fn) return false;
    CUdeviceptr b = 0;
    size_t sz = 0;
    if (fn(&b, &sz, reinterpret_cast<CUdeviceptr>(p)) != CUDA_SUCCESS) return false;
    *base = reinterpret_cast<const void *>(b);
    return true;
}

extern "C" int hb_ipc_export(const void *device_ptr, hb_ipc_mem *out) {
    HB_REQUIRE(device_ptr && out, HB_ERR_INVALID, "hb_ipc_export: null argument");
    // a CUDA IPC handle names a whole allocation: a pointer into the middle of one (e.g. a tensor carved out of a caching
    // allocator's block) would be opened at the block's base on the other side and the halo rows would land elsewhere
    const void *base = nullptr;
    HB_REQUIRE(!allocation_base(device_ptr, &base) || base == device_ptr, HB_ERR_INVALID,
               "hb_ipc_export: %p is not the base of its allocation (%p); allocate strip buffers with hb_image_create", device_ptr, base);
    static_assert(sizeof(cudaIpcMemHandle_t) == HB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    int rc = check_cuda(cudaIpcGetMemHandle(&h, const_cast<void *>(device_ptr)), "cudaIpcGetMemHandle()");
    if (rc) return rc;
    memcpy(out->handle, &h, sizeof(h));
    return HB_OK;
}

extern "C" int hb_ipc_open(const hb_ipc_mem *in, void **peer_ptr) {
    HB_REQUIRE(in && peer_ptr, HB_ERR_INVALID, "hb_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, in->handle, sizeof(h));
    return check_cuda(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle()");
}

extern "C" int hb_ipc_close(void *peer_ptr) {
    if (!peer_ptr) return HB_OK;
    return check_cuda(cudaIpcCloseMemHandle(peer_ptr), "cudaIpcCloseMemHandle()");
}

extern "C" int hb_halo_ctrl_create(void **ctrl) {
    HB_REQUIRE(ctrl, HB_ERR_INVALID, "hb_halo_ctrl_create: null argument");
    int rc = check_cuda(cudaMalloc(ctrl, GATHER_CTRL_INTS * sizeof(int)), "cudaMalloc(halo control block)");
    if (rc) return rc;
    rc = check_cuda(cudaMemset(*ctrl, 0, GATHER_CTRL_INTS * sizeof(int)), "cudaMemset(halo control block)");
    rc |= check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize()");
    return rc;
}

extern "C" int hb_halo_ctrl_destroy(void *ctrl) {
    if (!ctrl) return HB_OK;
    return check_cuda(cudaFree(ctrl), "cudaFree(halo control block)");
}

extern "C" int hb_halo_status(const void *ctrl, int *exchanges, int *timed_out) {
    HB_REQUIRE(ctrl, HB_ERR_INVALID, "hb_halo_status: null argument");
    int h[CTRL_INTS];
    int rc = check_cuda(cudaMemcpy(h, ctrl, sizeof(h), cudaMemcpyDeviceToHost), "cudaMemcpy(halo control block)");
    if (exchanges) *exchanges = h[0] < 0 ? -1 : h[0];   // -1: poisoned after a timeout
    if (timed_out) *timed_out = h[5];
    return rc;
}

// CTAs that share one push: one per 64 KiB, at most 32 (a 1-row halo stays a single CTA, the 5 MB level-0 halo of the
// sharded pyramid and the 0.5 MB strips of an all-gather spread over enough SMs to fill NVLink)
static int push_parts(size_t bytes) {
    const size_t n = (bytes + 65535) / 65536;
    return n < 1 ? 1 : n > 32 ? 32 : (int)n;
}

static int fill_halo_params(const hb_halo_desc *d, HaloParams &p) {
    HB_REQUIRE(d && d->buf && d->ctrl, HB_ERR_INVALID, "hb_halo_exchange: null buffer / control block");
    HB_REQUIRE(d->radius > 0 && d->rows >= d->radius && d->row_bytes > 0 && d->pitch_bytes >= d->row_bytes, HB_ERR_INVALID,
               "hb_halo_exchange: bad geometry (radius %d, rows %d)", d->radius, d->rows);
    HB_REQUIRE((d->up_buf == nullptr) == (d->up_ctrl == nullptr) && (d->down_buf == nullptr) == (d->down_ctrl == nullptr), HB_ERR_INVALID,
               "hb_halo_exchange: a neighbour needs both its buffer and its control block");
    HB_REQUIRE(!d->up_buf || d->ghost_top == d->radius, HB_ERR_INVALID, "hb_halo_exchange: ghost_top must equal the radius when an upper neighbour exists");
    memset(&p, 0, sizeof(p));
    p.buf = static_cast<unsigned char *>(d->buf); p.pitch = d->pitch_bytes; p.row_bytes = d->row_bytes;
    p.gt = d->ghost_top; p.rows = d->rows; p.R = d->radius;
    if (d->up_buf) {   // my top rows -> the upper neighbour's bottom ghost rows
        p.up_dst = static_cast<unsigned char *>(d->up_buf) + (size_t)(d->up_ghost_top + d->up_rows) * d->up_pitch_bytes;
        p.up_pitch = d->up_pitch_bytes; p.up_ctrl = static_cast<int *>(d->up_ctrl);
    }
    if (d->down_buf) {  // my bottom rows -> the lower neighbour's top ghost rows
        HB_REQUIRE(d->down_ghost_top >= d->radius, HB_ERR_INVALID, "hb_halo_exchange: the lower neighbour has no room for %d ghost rows", d->radius);
        p.down_dst = static_cast<unsigned char *>(d->down_buf) + (size_t)(d->down_ghost_top - d->radius) * d->down_pitch_bytes;
        p.down_pitch = d->down_pitch_bytes; p.down_ctrl = static_cast<int *>(d->down_ctrl);
    }
    p.ctrl = static_cast<int *>(d->ctrl);
    return HB_OK;
}

extern "C" int hb_halo_exchange_batch(const hb_halo_desc *const *descs, int n, void *stream) {
    HB_REQUIRE(descs && n >= 1 && n <= kMaxHalo, HB_ERR_INVALID, "hb_halo_exchange_batch: 1..%d descriptors", kMaxHalo);
    HaloBatch b;
    memset(&b, 0, sizeof(b));
    for (int i = 0; i < n; ++i) {
        const int rc = fill_halo_params(descs[i], b.p[i]);
        if (rc) return rc;
    }
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_halo_exchange");
    size_t bytes = 0;
    for (int i = 0; i < n; ++i) bytes = bytes > (size_t)b.p[i].R * b.p[i].row_bytes ? bytes : (size_t)b.p[i].R * b.p[i].row_bytes;
    halo_exchange_kernel<<<dim3(n, push_parts(bytes)), HALO_THREADS, 0, s>>>(b);
    g_launches++;
    return scope.finish();
}

extern "C" int hb_halo_exchange(const hb_halo_desc *d, void *stream) { return hb_halo_exchange_batch(&d, 1, stream); }

extern "C" int hb_allgather_rows(const hb_gather_desc *d, void *stream) {
    HB_REQUIRE(d && d->buf && d->ctrl, HB_ERR_INVALID, "hb_allgather_rows: null buffer / control block");
    HB_REQUIRE(d->n_peers >= 1 && d->n_peers <= kMaxPeers && d->n_peers <= HB_MAX_PEERS, HB_ERR_INVALID, "hb_allgather_rows: 1..%d peers", kMaxPeers);
    HB_REQUIRE(d->rows > 0 && d->row0 >= 0 && d->row_bytes > 0 && d->pitch_bytes >= d->row_bytes, HB_ERR_INVALID, "hb_allgather_rows: bad geometry");
    HB_REQUIRE(d->my_slot >= 0 && d->my_slot < 16, HB_ERR_INVALID, "hb_allgather_rows: slot out of range");
    GatherParams p;
    memset(&p, 0, sizeof(p));
    p.buf = static_cast<unsigned char *>(d->buf); p.pitch = d->pitch_bytes; p.row_bytes = d->row_bytes;
    p.row0 = d->row0; p.rows = d->rows; p.my_slot = d->my_slot; p.n_peers = d->n_peers;
    p.ctrl = static_cast<int *>(d->ctrl);
    for (int j = 0; j < d->n_peers; ++j) {
        HB_REQUIRE(d->peer_buf[j] && d->peer_ctrl[j] && d->peer_slot[j] >= 0 && d->peer_slot[j] < 16 && d->peer_slot[j] != d->my_slot, HB_ERR_INVALID,
                   "hb_allgather_rows: peer %d incomplete", j);
        p.peer_buf[j] = static_cast<unsigned char *>(d->peer_buf[j]);
        p.peer_ctrl[j] = static_cast<int *>(d->peer_ctrl[j]);
        p.peer_slot[j] = d->peer_slot[j];
    }
    cudaStream_t s = (cudaStream_t)stream;
    OpScope scope(s, "hb_allgather_rows");
    allgather_rows_kernel<<<dim3(d->n_peers, push_parts((size_t)d->rows * d->row_bytes)), HALO_THREADS, 0, s>>>(p);
    g_launches++;
    return scope.finish();
}
