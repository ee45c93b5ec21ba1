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

// K-way merge of n_lists ascending lists of LLAMPC_LIST_LEN keys into the K smallest keys overall (K <= LLAMPC_LIST_LEN):
// K rounds of "block-min over the list heads"; each thread owns up to MERGE_LPT lists and keeps their head and
// next key in registers so that the load of a popped list's successor is off the critical path.
// out[0] = *best_key (then re-armed to ~0 for the next tick), out[1..K] = ascending top-K.  sbuf: THREADS/32+1 keys.
constexpr int MERGE_LPT = 8;

template <int THREADS>
__device__ __forceinline__ void merge_lists_device(const u64* __restrict__ lists, int n_lists, int K,
                                                   u64* __restrict__ best_key, u64* __restrict__ out, u64* sbuf) {
    u64 head[MERGE_LPT], next[MERGE_LPT];
    int pos[MERGE_LPT];
#pragma unroll
    for (int j = 0; j < MERGE_LPT; ++j) {
        const int l = threadIdx.x + j * THREADS;
        head[j] = ~0ull; next[j] = ~0ull; pos[j] = 1;
        if (l < n_lists) {
            head[j] = __ldcg(lists + (size_t)l * LLAMPC_LIST_LEN);
            next[j] = __ldcg(lists + (size_t)l * LLAMPC_LIST_LEN + 1);
        }
    }
    if (threadIdx.x == 0 && best_key) {
        out[0] = __ldcg(best_key);
        *best_key = ~0ull;
    }
    for (int r = 0; r < K; ++r) {
        u64 mine = head[0];
#pragma unroll
        for (int j = 1; j < MERGE_LPT; ++j) mine = u64_min(mine, head[j]);
        const u64 sel = block_min_u64_w0<THREADS / 32>(mine, sbuf);
        if (threadIdx.x == 0) out[1 + r] = sel;
        if (sel == ~0ull || mine != sel) continue;
#pragma unroll
        for (int j = 0; j < MERGE_LPT; ++j) {
            if (head[j] == sel) {                  // pop: successor becomes the head, prefetch the one after
                head[j] = next[j];
                pos[j] += 1;
                const int l = threadIdx.x + j * THREADS;
                next[j] = pos[j] < LLAMPC_LIST_LEN ? __ldcg(lists + (size_t)l * LLAMPC_LIST_LEN + pos[j]) : ~0ull;
            }
        }
    }
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
