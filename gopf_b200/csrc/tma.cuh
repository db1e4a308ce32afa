// Blackwell bulk-copy (TMA) and mbarrier primitives for the warp-specialised line kernels
// (tma_kernels.cuh), plus the host-side tensor-map encoder.
//
// Device side: thin wrappers over the PTX instructions (SASS: UTMALDG / UTMASTG / UBLKCP, SYNCS).
// One elected thread arms an mbarrier with the byte count of a tile and issues the bulk copy; the
// copy engine completes the transaction on that barrier while every warp of the CTA keeps computing.
// Stores go the other way from shared memory (`bulk_group` completion, waited for by the issuing
// thread only when the buffer is needed again).
//
// Host side: cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint (no link-time
// dependency on libcuda: the library still loads on a box without a driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace gopf {
namespace tma {

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the barrier initialisation visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// orders this thread's generic-proxy shared-memory accesses before later async-proxy ones (bulk stores
// reading, bulk loads overwriting the same buffer)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}
// A wait that never completes would hang the device until the driver's watchdog (or the caller's patience) ends
// it.  Every spin loop of these kernels therefore gives up after GOPF_TMA_WAIT_LIMIT_NS (4 s: four orders of
// magnitude above any legitimate wait), reports where, and traps: the launch fails with an error instead.
#define GOPF_TMA_WAIT_LIMIT_NS 4000000000ULL
__device__ __noinline__ void wait_timed_out(int where, unsigned a, unsigned b) {
    printf("gopf: copy-engine kernel wait %d timed out (block %d thread %d, %u %u)\n", where, (int)blockIdx.x, (int)threadIdx.x, a, b);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, int where = 0) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > GOPF_TMA_WAIT_LIMIT_NS) wait_timed_out(where, parity, spins);
    }
}

// ---- tensor (tiled) copies, rank 4: coordinates innermost first -------------------------------------
__device__ __forceinline__ void load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void store_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];\n" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
                 "r"(c3)
                 : "memory");
}
// rank 5 (the blocked k-space layout: column, row_low, row_high, slab_low, slab_high)
__device__ __forceinline__ void load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                        uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void store_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];\n" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global [%0, {%1, %2, %3, %4, %5}];\n" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void prefetch_descriptor(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

// ---- linear bulk copies (contiguous lines; no tensor map) -------------------------------------------
__device__ __forceinline__ void load_1d(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void store_1d(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// all of this thread's committed bulk stores have finished READING shared memory (buffer reusable)
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// ... have completed (writes visible); needed before kernel exit only implicitly
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// named barrier over one consumer group (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}
#endif  // __CUDACC__

// ---- host: fp64 tensor map (rank 2..5) over a complex128 array ------------------------------------------------
// dims / strides in complex cells (innermost first; stride[0] is implicitly 1 cell), box in cells.  The map
// is encoded over doubles (2 per cell) because there is no 16-byte element type.  l2_promotion: 0 none,
// 1 64 B, 2 128 B, 3 256 B (the granularity at which L2 fills from DRAM: narrow row segments of adjacent
// tiles then share one DRAM burst).
// swizzle: 0 none, 1 32 B, 2 64 B, 3 128 B (16-byte chunks XOR-ed with shared-memory address bits 7..: the
// inner box must not exceed the swizzle span).
inline cudaError_t encode_c128(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                               const unsigned long long* strides_cells, const unsigned* box, int l2_promotion,
                               int swizzle = 0) {
    if (rank < 2 || rank > 5) return cudaErrorInvalidValue;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (!p || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        fn = reinterpret_cast<EncodeFn>(p);
    }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) {
        gd[i] = i == 0 ? dims[0] * 2 : dims[i];
        bx[i] = i == 0 ? box[0] * 2 : box[i];
        if (i > 0) gs[i - 1] = strides_cells[i - 1] * 16;
    }
    CUtensorMapL2promotion promo = l2_promotion == 3   ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                   : l2_promotion == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                   : l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                       : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    const CUtensorMapSwizzle sw = swizzle == 3   ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace tma
}  // namespace gopf
