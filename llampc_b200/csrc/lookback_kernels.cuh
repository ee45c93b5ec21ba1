// Internal launcher interface between the look-back translation units (one .cu per kernel family so that the 40-odd
// template instantiations compile in parallel) and the C ABI in lookback.cu.
#pragma once
#include "llampc_launch.cuh"
#include "llampc_model.cuh"
#include "lookback_select.cuh"

namespace llampc {

// Everything a window-recompute launch needs (K1 / K1p).
struct LbArgs {
    const float4* bank; int N, Npad;
    const float* hist; int W; long hist_stride_floats; int n_vehicles;
    StepSize z;
    float* avg_err;              // [n_vehicles][N] or NULL
    u64* cta_lists;              // [n_vehicles][grid.x][LIST_LEN] (grid path with the last-CTA merge) or NULL
    int idx_offset;
    NewRow nr;                   // newest history row riding in the kernel parameters (slot < 0: none)
    FusedMerge fm;               // last-CTA merge per vehicle (K = 0: off)
    PeerXchg px;                 // NVLink min-loc (world = 0: off)
    TreeMerge tm;                // in-kernel tree merge, single history (K = 0: off)
    bool pdl;                    // programmatic dependent launch (K1p, tree mode)
    bool wide;                   // K1p: the full-range slip-angle form of the step (LLAMPC_LB_FLAG_WIDE)
};

// candidates per CTA of the grid kernels for a window split of `sy`
static inline int k1_cands_per_cta(bool packed, int sy) { return (packed ? 2 * LB_THREADS : LB_THREADS) / sy; }

int launch_k1_scalar(const LbArgs& a, int sy, bool geom, bool mufu, cudaStream_t st);      // lookback_k1.cu
int launch_k1_packed(const LbArgs& a, int sy, bool geom, bool mufu, cudaStream_t st);      // lookback_k1p.cu

// K1e: packed window kernel on equal warp shares, single history, merge tree (lookback_k1e.cu)
int lookback_equal_plan(int N, int W, int* grid, int* block, size_t* bytes, TreeLayout* lay);
int lookback_equal_launch(const float4* bank, int N, int Npad, const float* hist, int W, const StepSize& z, float* avg_err,
                          int idx_offset, bool geom, bool mufu, int K, void* workspace, unsigned long long workspace_bytes,
                          u64* out, const NewRow& nr, const PeerXchg& px, cudaStream_t st);

// K1r / K1v: rolling window (lookback_rolling.cu)
constexpr int RV_THREADS = 256;
constexpr int RV_WARPS = RV_THREADS / 32;
constexpr int RV_CPP = RV_THREADS * 4;             // candidates per pass
constexpr int RV_MAX_PASSES = 2;
constexpr int RV_MAX_N = RV_CPP * RV_MAX_PASSES;   // 2,048
int launch_k1r(const float4* bank, int N, int Npad, int W, StepSize z, const NewRow& nr, const float* hist, int n_vehicles,
               float* err_ring, float* avg_err, u64* cta_lists, int idx_offset, int emit, const FusedMerge& fm, bool geom,
               bool mufu, cudaStream_t st);
int launch_k1v(const float4* bank, int N, int Npad, int W, StepSize z, int slot, const float* hist, int n_vehicles,
               float* err_ring, float* avg_err, int idx_offset, int emit, int K, u64* out, bool geom, bool mufu,
               cudaStream_t st);

// Top-K of a vehicle's keys held in shared memory, by FILTERING instead of sorting (K1v):
//   1  every group of 32 keys left its minimum in s_group (ng <= 64 groups);
//   2  warp 0 sorts the minima; T = the K-th smallest.  At least K keys are <= T (those minima themselves), so the
//      top-K is a subset of {key <= T} -- about K .. 3K keys of the N;
//   3  the survivors are compacted with a shared-memory counter and warp 0 sorts them 32 at a time (sort + bitonic
//      merge into the running 32 smallest).  Keys are unique (index in the low word), so the result does not depend
//      on the compaction order.
// Call with the whole CTA after a __syncthreads() that made s_key / s_group visible.  o: [LIST_LEN + 1] keys
// (o[0] = arg-min key, o[1..K] = ascending top-K).
template <int THREADS>
__device__ __forceinline__ void cta_topk_filter(const u64* s_key, int n_keys, const u64* s_group, int ng, int K, u64* s_cand,
                                                u64* s_thr, int* s_count, u64* __restrict__ o) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {                                               // threshold = K-th smallest group minimum
        u64 a = lane < ng ? s_group[lane] : ~0ull;
        a = warp_sort_u64(a, lane);
        if (ng > 32) {
            u64 b = lane + 32 < ng ? s_group[lane + 32] : ~0ull;
            b = warp_sort_u64(b, lane);
            a = warp_merge_low32(a, __shfl_sync(0xffffffffu, b, 31 - lane), lane);
        }
        const u64 t = __shfl_sync(0xffffffffu, a, K - 1);          // ~0 when fewer than K groups hold a key: keep all
        if (lane == 0) { *s_thr = t; *s_count = 0; }
    }
    __syncthreads();
    const u64 T = *s_thr;
    for (int i = tid; i < n_keys; i += THREADS) {
        const u64 k = s_key[i];
        if (k <= T && k != ~0ull) s_cand[atomicAdd(s_count, 1)] = k;
    }
    __syncthreads();
    if (warp != 0) return;
    const int n = *s_count;
    u64 run = ~0ull;
    for (int c = 0; c < n; c += 32) {
        u64 k = c + lane < n ? s_cand[c + lane] : ~0ull;
        k = warp_sort_u64(k, lane);
        run = c == 0 ? k : warp_merge_low32(run, __shfl_sync(0xffffffffu, k, 31 - lane), lane);
    }
    if (lane == 0) o[0] = run;
    if (lane < LLAMPC_LIST_LEN) o[1 + lane] = lane < K ? run : ~0ull;
}

}  // namespace llampc
