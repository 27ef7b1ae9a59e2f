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

__device__ __forceinline__ void push_rows(unsigned char *dst, size_t dpitch, const unsigned char *src, size_t spitch, size_t row_bytes, int rows) {
    if ((row_bytes | dpitch | spitch | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) % 16 == 0) {
        const size_t vpr = row_bytes / 16;
        for (size_t i = threadIdx.x; i < vpr * rows; i += blockDim.x) {
            const size_t r = i / vpr, c = i - r * vpr;
            reinterpret_cast<uint4 *>(dst + r * dpitch)[c] = reinterpret_cast<const uint4 *>(src + r * spitch)[c];
        }
    } else {
        for (size_t i = threadIdx.x; i < row_bytes * rows; i += blockDim.x) {
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

__global__ void __launch_bounds__(1024) halo_exchange_kernel(const __grid_constant__ HaloBatch batch) {
    const HaloParams &p = batch.p[blockIdx.x];
    __shared__ int ok;
    const int e = p.ctrl[0];   // read by every thread before thread 0 advances it (barriers below)
    if (e < 0) return;         // poisoned by an earlier timeout
    if (threadIdx.x == 0) {
        // my ghost rows of exchange e-1 are consumed: the neighbours may overwrite them
        if (p.up_ctrl) st_release_sys(p.up_ctrl + 4, e);      // I am the upper neighbour's lower neighbour
        if (p.down_ctrl) st_release_sys(p.down_ctrl + 3, e);
        bool good = true;
        if (p.up_ctrl) good = spin_until(p.ctrl + 3, e) && good;
        if (p.down_ctrl) good = spin_until(p.ctrl + 4, e) && good;
        ok = good ? 1 : 0;
    }
    __syncthreads();
    if (ok) {
        if (p.up_dst) push_rows(p.up_dst, p.up_pitch, p.buf + (size_t)p.gt * p.pitch, p.pitch, p.row_bytes, p.R);
        if (p.down_dst) push_rows(p.down_dst, p.down_pitch, p.buf + (size_t)(p.gt + p.rows - p.R) * p.pitch, p.pitch, p.row_bytes, p.R);
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool good = ok != 0;
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
    int rc = check_cuda(cudaMalloc(ctrl, CTRL_INTS * sizeof(int)), "cudaMalloc(halo control block)");
    if (rc) return rc;
    rc = check_cuda(cudaMemset(*ctrl, 0, CTRL_INTS * sizeof(int)), "cudaMemset(halo control block)");
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
    halo_exchange_kernel<<<n, 1024, 0, s>>>(b);
    g_launches++;
    return scope.finish();
}

extern "C" int hb_halo_exchange(const hb_halo_desc *d, void *stream) { return hb_halo_exchange_batch(&d, 1, stream); }
