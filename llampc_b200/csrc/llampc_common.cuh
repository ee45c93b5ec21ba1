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

// warp-wide minimum of a packed key with the hardware integer reduction (REDUX): high word first, then the low word
// among the lanes that hold the winning high word -- two reductions instead of a five-step shuffle ladder.
__device__ __forceinline__ u64 warp_min_key(u64 k) {
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((u64)mh << 32) | ml;
}

// Global top-K (K <= LLAMPC_LIST_LEN) of n_lists ascending lists of LLAMPC_LIST_LEN keys, without a memory load
// inside any selection round.  Key fact: a key that is not the head of its list can only be in the global top-K if
// the head of that list is too, so only the K lists with the smallest heads can contribute.
//   A1  every warp selects the K smallest heads of the lists it owns (heads in registers, K REDUX rounds);
//   A2  warp 0 merges the per-warp selections -> the K lists with the globally smallest heads;
//   B   warp 0 loads those <= K lists (one coalesced wave of loads) and K-way merges them from shared memory.
// out[0] = arg-min key (= out[1]), out[1..K] = ascending top-K, out[K+1..LIST_LEN] = ~0.
// Capacity: THREADS * MERGE_LPT lists.  Shared memory: MergeSmem<THREADS>.
constexpr int MERGE_LPT = 8;

template <int THREADS>
struct MergeSmem {
    u64 wkey[THREADS / 32][LLAMPC_LIST_LEN];
    int wid[THREADS / 32][LLAMPC_LIST_LEN];
    int sel_list[LLAMPC_LIST_LEN];
    u64 rows[LLAMPC_LIST_LEN][LLAMPC_LIST_LEN];
};

template <int THREADS>
__device__ __forceinline__ void merge_lists_device(const u64* __restrict__ lists, int n_lists, int K,
                                                   u64* __restrict__ out, MergeSmem<THREADS>& sm) {
    constexpr int NW = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 head[MERGE_LPT];
#pragma unroll
    for (int j = 0; j < MERGE_LPT; ++j) {
        const int l = threadIdx.x + j * THREADS;
        head[j] = l < n_lists ? __ldcg(lists + (size_t)l * LLAMPC_LIST_LEN) : ~0ull;
    }
    // A1: per-warp K smallest heads
    for (int r = 0; r < LLAMPC_LIST_LEN; ++r) {
        u64 sel = ~0ull;
        if (r < K) {
            u64 mine = head[0];
#pragma unroll
            for (int j = 1; j < MERGE_LPT; ++j) mine = u64_min(mine, head[j]);
            sel = warp_min_key(mine);
            if (sel != ~0ull && mine == sel) {
#pragma unroll
                for (int j = 0; j < MERGE_LPT; ++j)
                    if (head[j] == sel) { head[j] = ~0ull; sm.wid[warp][r] = threadIdx.x + j * THREADS; }
            }
        }
        if (lane == 0) sm.wkey[warp][r] = sel;
    }
    __syncthreads();
    if (warp != 0) return;
    // A2: the K lists with the globally smallest heads
    {
        int p = 0;
        u64 h = lane < NW ? sm.wkey[lane][0] : ~0ull;
        for (int r = 0; r < LLAMPC_LIST_LEN; ++r) {
            u64 sel = ~0ull;
            if (r < K) sel = warp_min_key(h);
            if (sel == ~0ull) {
                if (lane == 0) sm.sel_list[r] = -1;
            } else if (h == sel) {
                sm.sel_list[r] = sm.wid[lane][p];
                ++p;
                h = p < LLAMPC_LIST_LEN ? sm.wkey[lane][p] : ~0ull;
            }
        }
    }
    __syncwarp();
    // B: fetch the selected lists, then merge them from shared memory
    for (int i = lane; i < LLAMPC_LIST_LEN * LLAMPC_LIST_LEN; i += 32) {
        const int l = sm.sel_list[i / LLAMPC_LIST_LEN];
        (&sm.rows[0][0])[i] = l >= 0 ? __ldcg(lists + (size_t)l * LLAMPC_LIST_LEN + (i % LLAMPC_LIST_LEN)) : ~0ull;
    }
    __syncwarp();
    {
        int p = 0;
        u64 h = lane < LLAMPC_LIST_LEN ? sm.rows[lane][0] : ~0ull;
        for (int r = 0; r < LLAMPC_LIST_LEN; ++r) {
            const u64 sel = r < K ? warp_min_key(h) : ~0ull;
            if (lane == 0) {
                out[1 + r] = sel;
                if (r == 0) out[0] = sel;                          // arg-min key = head of the top-K
            }
            if (sel != ~0ull && h == sel) {
                ++p;
                h = p < LLAMPC_LIST_LEN ? sm.rows[lane][p] : ~0ull;
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

// wait for the phase with the given parity; safe to call from several places of one kernel (unique labels)
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

}  // namespace llampc
