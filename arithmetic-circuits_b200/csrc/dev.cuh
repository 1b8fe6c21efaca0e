// Device-side helpers shared by the kernels: mbarrier / TMA bulk-copy PTX wrappers, warp utilities.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fr.cuh"

namespace acg {

// ---- mbarrier + bulk async copy (TMA, SASS: UBLKCP / SYNCS) -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// Same copy for data that is read exactly once per pass (the tile stream): L2 evict-first, so that the stream
// does not push the witness -- which every tile re-reads through windows and far gathers -- out of L2.
#ifndef ACG_STREAM_EVICT_FIRST
#define ACG_STREAM_EVICT_FIRST 1
#endif
__device__ __forceinline__ void tma_load_1d_stream(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
#if ACG_STREAM_EVICT_FIRST
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
#else
    tma_load_1d(smem_dst, gmem_src, bytes, bar);
#endif
}
// .. and for the little that the NEXT launch reads before anything else (the first tile of every CTA's run): L2
// evict-last, so that back-to-back checks of a resident system start from L2 instead of DRAM
__device__ __forceinline__ void tma_load_1d_evict_last(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
// pull [p, p + bytes) towards L2 (bytes % 16 == 0, p 16-byte aligned); no completion to wait for
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// the same for read-once data: the prefetched lines are marked evict-first
#ifndef ACG_PREFETCH_EVICT_FIRST
#define ACG_PREFETCH_EVICT_FIRST 0
#endif
__device__ __forceinline__ void prefetch_l2_bulk_stream(const void* p, uint32_t bytes) {
#if ACG_PREFETCH_EVICT_FIRST
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(pol)
                 : "memory");
#else
    prefetch_l2_bulk(p, bytes);
#endif
}
// bulk copy of witness data (re-read by many tiles): L2 evict-last
#ifndef ACG_WITNESS_EVICT_LAST
#define ACG_WITNESS_EVICT_LAST 0
#endif
__device__ __forceinline__ void tma_load_1d_keep(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
#if ACG_WITNESS_EVICT_LAST
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
#else
    tma_load_1d(smem_dst, gmem_src, bytes, bar);
#endif
}
// 16-byte asynchronous global -> shared copy that bypasses L1 (SASS LDGSTS.BYPASS), and the wait for all of this
// thread's copies
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;" ::: "memory");
}
// one witness element (32 bytes) through the read-only path
__device__ __forceinline__ fr_t ld_witness(const fr_t* p) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
// order this thread's generic-proxy shared-memory accesses before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// programmatic dependent launch (sm_90+): let the next kernel of the stream start launching / wait for the previous
// kernel of the stream to have completed and flushed (both are no-ops for an ordinary launch)
__device__ __forceinline__ void griddep_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint32_t lane_id() {
    return threadIdx.x & 31u;
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// 64-bit atomics on the {n_violations, first_bad_row} result pair
__device__ __forceinline__ void report_bad_rows(unsigned long long* result, uint32_t bad_ballot, uint64_t warp_row0) {
    // called by one lane per warp with the ballot of violated rows (lane i <-> row warp_row0 + i)
    atomicAdd(&result[0], (unsigned long long)__popc(bad_ballot));
    atomicMin(&result[1], (unsigned long long)(warp_row0 + (uint64_t)(__ffs(bad_ballot) - 1)));
}

}  // namespace acg
