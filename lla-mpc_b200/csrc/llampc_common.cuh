// Shared helpers: error mapping, packed keys, warp/block min, mbarrier + 1-D TMA bulk copy (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/llampc_b200.h"

namespace llampc {

typedef unsigned long long u64;

#define LLAMPC_CUDA_TRY(expr)                         \
    do {                                              \
        cudaError_t _e = (expr);                      \
        if (_e != cudaSuccess) return (int)_e;        \
    } while (0)

__device__ __forceinline__ u64 pack_key(float err, unsigned idx) {
    // non-negative finite floats order like their bit patterns; NaN (0x7fc00000) sorts above +inf.
    return ((u64)__float_as_uint(err) << 32) | (u64)idx;
}

__device__ __forceinline__ u64 u64_min(u64 a, u64 b) { return a < b ? a : b; }

__device__ __forceinline__ u64 warp_min_u64(u64 k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k = u64_min(k, __shfl_xor_sync(0xffffffffu, k, o));
    return k;
}

// min over a CTA of NWARPS warps; result valid in every thread.  `sbuf` holds NWARPS+1 keys.
template <int NWARPS>
__device__ __forceinline__ u64 block_min_u64(u64 k, u64* sbuf) {
    k = warp_min_u64(k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                       // sbuf may still be read from a previous round
    if (lane == 0) sbuf[warp] = k;
    __syncthreads();
    u64 r = sbuf[0];
#pragma unroll
    for (int i = 1; i < NWARPS; ++i) r = u64_min(r, sbuf[i]);
    return r;
}

__device__ __forceinline__ u64 u64_max(u64 a, u64 b) { return a > b ? a : b; }

// block min for the selection kernels: the second stage is done by warp 0 alone (32 partials, one per lane),
// so a round costs two barriers and ~60 instructions per thread instead of a 32-way scan in every thread.
template <int NWARPS>
__device__ __forceinline__ u64 block_min_u64_w0(u64 k, u64* sbuf) {
    k = warp_min_u64(k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sbuf[warp] = k;
    __syncthreads();
    if (warp == 0) {
        u64 r = lane < NWARPS ? sbuf[lane] : ~0ull;
        r = warp_min_u64(r);
        if (lane == 0) sbuf[NWARPS] = r;
    }
    __syncthreads();
    const u64 r = sbuf[NWARPS];
    __syncthreads();                       // sbuf is rewritten by the next round
    return r;
}

// ascending bitonic sort of one key per lane (15 compare-exchange stages of 2 shuffles each)
__device__ __forceinline__ u64 warp_sort_u64(u64 key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const u64 other = __shfl_xor_sync(0xffffffffu, key, j);
            const bool up = (lane & k) == 0 || k == 32;
            const bool lower = (lane & j) == 0;
            key = (lower == up) ? u64_min(key, other) : u64_max(key, other);
        }
    }
    return key;
}

// a: this warp's ascending keys (lane i = i-th smallest); b_rev: the other sorted list read in REVERSED lane
// order.  Returns the 32 smallest of the union, ascending (bitonic merge, 5 stages).
__device__ __forceinline__ u64 warp_merge_low32(u64 a, u64 b_rev, int lane) {
    u64 m = u64_min(a, b_rev);
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const u64 other = __shfl_xor_sync(0xffffffffu, m, j);
        m = ((lane & j) == 0) ? u64_min(m, other) : u64_max(m, other);
    }
    return m;
}

// ---- mbarrier + cp.async.bulk (TMA 1-D bulk copy global -> shared; SASS: UBLKCP) ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

}  // namespace llampc
