// hb_tma.cuh -- Tensor Memory Accelerator (TMA) tile loads + mbarrier helpers for sm_100a.
//
// The local-operator kernels stage "output tile + halo" boxes of the input image into shared memory.
// Interior tiles (box entirely inside the accessor's boundary window) are fetched by ONE thread with a
// single cp.async.bulk.tensor.2d (SASS: UTMALDG) that completes on an mbarrier, so the other threads
// spend no instructions on staging and the copy of tile i+1 overlaps the arithmetic of tile i.
// Border tiles are staged by all threads through the boundary-mode remap (hb_common.cuh).
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudart, no -lcuda)
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (TMA unit)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes / reads of a buffer happen-before a later async-proxy (TMA) write to it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HB_DONE_%=;\n"
        "bra HB_WAIT_%=;\n"
        "HB_DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}

// one box of the tensor described by `map` -> shared memory; (x, y) = element coordinates of the box
// origin (may be negative / beyond the extent: out-of-bounds elements are zero-filled)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

// ---- host side -------------------------------------------------------------------------------------
// Encode a 2-D tiled tensor map over an image (row pitch in bytes must be a multiple of 16 and the base
// 16-byte aligned; box_w * elem_size must be a multiple of 16, box dims <= 256).  Returns false when the
// image cannot be described (caller falls back to the all-threads loader -- still on the device).
bool make_tile_map(CUtensorMap *out, const void *base, int dtype, int img_w, int img_h, int stride_px, int box_w, int box_h);
bool tma_addressable(const void *base, int dtype, int stride_px);

}  // namespace hb
