// K1r / K1v: the reference's rolling error-window bookkeeping on the device (rt.py:349-358).  sm_100a.
#include "lookback_kernels.cuh"
#include "llampc_packed.cuh"

namespace llampc {

// ---------------------------------------------------------------------------------------------------
// K1r rolling window (the reference's own bookkeeping, rt.py:349-358): only the newest transition is integrated
// (one RK4 step per candidate), its error replaces ring column `slot` of err_ring (np.roll + write of the last
// column), and the window mean is re-summed from the ring -- N steps per tick instead of N*W steps.
// emit = 0 while the window is filling (columns stored, no decision).
//
// Ring layout and summation order (shared with K1v, so both give bit-identical means): err_ring [LLAMPC_RING_ROWS(W)][Npad]
// = W error columns followed by ceil(W / 4) rows of PARTIAL SUMS, partial_j = ((e_4j + e_4j+1) + e_4j+2) + e_4j+3, and
// the window sum is partial_0 + partial_1 + ... in that order.  A tick changes one column, hence one partial: K1v reads
// the 3 other columns of that group and the other partials (7 rows instead of 19 at W = 20); K1r, whose tick is
// latency-bound anyway, re-adds every column in the same order and keeps the partial row up to date.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int ring_groups(int W) { return (W + 3) >> 2; }

template <bool GEOM_SHARED, bool MUFU_SIN>
__global__ void __launch_bounds__(LB_THREADS)
lookback_rolling_kernel(const float4* __restrict__ bank, int N, int Npad, int W, StepSize z, NewRow nr,
                        const float* __restrict__ hist, float* __restrict__ err_ring, float* __restrict__ avg_err,
                        u64* __restrict__ cta_lists, int idx_offset, int emit, FusedMerge fm) {
    __shared__ u64 skeys[LB_THREADS];
    const int tid = threadIdx.x;
    const int v = blockIdx.y;                      // vehicle (Monte-Carlo layout); 0 for a single loop
    const int cand = blockIdx.x * LB_THREADS + tid;
    const bool valid = cand < N;
    const int ci = valid ? cand : N - 1;
    const Cand p = load_cand(bank, Npad, ci);
    HistRow r;
    if (hist) {                                    // rows of many vehicles: ring slot `slot` of hist [V][W][20]
        const float4* hr = reinterpret_cast<const float4*>(hist + ((size_t)v * W + nr.slot) * LLAMPC_HIST_ROW);
        r.q0 = __ldg(hr); r.q1 = __ldg(hr + 1); r.q2 = __ldg(hr + 2); r.q3 = __ldg(hr + 3); r.q4 = __ldg(hr + 4);
    } else {                                       // single loop: the row rides in the kernel parameters
        r.q0 = make_float4(nr.v[0], nr.v[1], nr.v[2], nr.v[3]);
        r.q1 = make_float4(nr.v[4], nr.v[5], nr.v[6], nr.v[7]);
        r.q2 = make_float4(nr.v[8], nr.v[9], nr.v[10], nr.v[11]);
        r.q3 = make_float4(nr.v[12], nr.v[13], nr.v[14], nr.v[15]);
        r.q4 = make_float4(nr.v[16], nr.v[17], nr.v[18], nr.v[19]);
    }
    err_ring += (size_t)v * (W + ring_groups(W)) * Npad;
    bool ok;
    float e = lookback_step_fast<GEOM_SHARED, MUFU_SIN>(p, r, z, ok);
    if (!ok) e = lookback_step<GEOM_SHARED, MUFU_SIN>(p, r, z);
    e *= 0.25f;                                    // errors of rt.py:349 (mean over the 4 scored states)
    if (valid) err_ring[(size_t)nr.slot * Npad + cand] = e;
    // window sum in the canonical grouped order; the group of the new column also refreshes its partial-sum row
    float sum = 0.0f;
    const float* col = err_ring + ci;
    const int gnew = nr.slot >> 2;
    for (int g0 = emit ? 0 : gnew; g0 < (emit ? ring_groups(W) : gnew + 1); ++g0) {
        float vq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int w = 4 * g0 + j;
            vq[j] = (w < W && w != nr.slot) ? __ldcg(col + (size_t)w * Npad) : 0.0f;
        }
        float part = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int w = 4 * g0 + j;
            if (w < W) part = j == 0 ? (w == nr.slot ? e : vq[0]) : part + (w == nr.slot ? e : vq[j]);
        }
        if (g0 == gnew && valid) err_ring[(size_t)(W + gnew) * Npad + cand] = part;
        sum = g0 == 0 ? part : sum + part;
    }
    if (!emit) return;                             // uniform
    const float err = sum / (float)W;
    u64 key = ~0ull;
    if (valid) {
        if (avg_err) avg_err[(size_t)v * N + cand] = err;
        key = pack_key(err, (unsigned)(idx_offset + cand));
    }
    cta_select_emit<LB_THREADS / 32>(key, skeys, v, cta_lists);
    if (fm.K > 0) {                                // last CTA of the vehicle merges its lists (one launch per tick)
        __shared__ bool is_last;
        __shared__ MergeSmem<LB_THREADS> msm;
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = atomicAdd(fm.ticket + v, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        merge_lists_device<LB_THREADS>(cta_lists + (size_t)v * gridDim.x * LLAMPC_LIST_LEN, gridDim.x, fm.K,
                                       fm.out + (size_t)v * (LLAMPC_LIST_LEN + 1), msm);
        if (tid == 0) fm.ticket[v] = 0;
    }
}

// ---------------------------------------------------------------------------------------------------
// K1v rolling window, ONE CTA PER VEHICLE (Monte-Carlo layout: thousands of vehicles, a bank of <= 2,048 candidates
// each).  Same arithmetic per candidate and same summation order as K1r; what changes is everything around the step
// (ncu on K1r at 4,096 x 1,024 x 20: ~1,200 instructions per candidate-tick, ~800 of them selection).
//   * Every thread owns FOUR ADJACENT candidates: ring columns and partial sums move as LDG.128 (the strided row
//     addresses cannot be immediates), the new column is read back by the thread that stored it.
//   * The window sum uses the partial-sum rows of the ring (see K1r): the 3 other columns of the new column's group and
//     the other ceil(W/4) - 1 partials -- 7 LDG.128 per thread instead of 19 at W = 20, 150 MB of ring traffic per tick
//     instead of 335 MB at 4,096 x 1,024: 103.5 -> 99 us there, 76 -> 64 us at W = 50 (the kernel is bound by the issue
//     slots of the 397-instruction scalar step, not by HBM).
//   * The step stays scalar (64 registers, 4 CTAs per SM = 32 warps).  The packed f32x2 step (LLAMPC_RV_PACKED=1: 118-126
//     registers, 2 CTAs per SM) was measured again with the partial sums in place: 126 us -- sixteen warps do not cover the
//     latency of the ring reads.
//   * The vehicle's keys stay in shared memory and are FILTERED, not sorted (cta_topk_filter, lookback_kernels.cuh):
//     every group of 32 keys (warp x candidate slot) leaves its minimum (two REDUX), <= 64 minima per vehicle; the K-th
//     smallest minimum is a threshold that at least K keys pass; the survivors (K .. 3K keys) are sorted by warp 0.
//   * The ring lines a warp will read are prefetched into L2 before its RK4 steps.
// Needs Npad % 4 == 0 and a 16-byte aligned ring (the entry point falls back to K1r otherwise).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2_4lines(const void* p, int n_lines) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    if (n_lines > 1) asm volatile("prefetch.global.L2 [%0+128];" ::"l"(p));
    if (n_lines > 2) asm volatile("prefetch.global.L2 [%0+256];" ::"l"(p));
    if (n_lines > 3) asm volatile("prefetch.global.L2 [%0+384];" ::"l"(p));
}

#ifndef LLAMPC_RV_PACKED
#define LLAMPC_RV_PACKED 0
#endif
#ifndef LLAMPC_RV_MIN_BLOCKS
#define LLAMPC_RV_MIN_BLOCKS (LLAMPC_RV_PACKED ? 2 : 4)
#endif
template <bool GEOM_SHARED, bool MUFU_SIN>
__global__ void __launch_bounds__(RV_THREADS, LLAMPC_RV_MIN_BLOCKS)
lookback_rolling_vehicle_kernel(const float4* __restrict__ bank, int N, int Npad, int W, StepSize z, int slot,
                                const float* __restrict__ hist, float* __restrict__ err_ring,
                                float* __restrict__ avg_err, int idx_offset, int emit, int K, u64* __restrict__ out) {
    __shared__ float4 srow[5];
    __shared__ u64 s_key[RV_MAX_N];
    __shared__ u64 s_cand[RV_MAX_N];
    __shared__ u64 s_group[RV_MAX_PASSES * RV_WARPS * 4];
    __shared__ u64 s_thr;
    __shared__ int s_count;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v = blockIdx.x;
    const int passes = (N + RV_CPP - 1) / RV_CPP;
    const int G = ring_groups(W), gnew = slot >> 2;
    if (tid < 5) srow[tid] = __ldg(reinterpret_cast<const float4*>(hist + ((size_t)v * W + slot) * LLAMPC_HIST_ROW) + tid);
    err_ring += (size_t)v * (W + G) * Npad;
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < passes; ++j) {
        const int wbase = j * RV_CPP + warp * 128;                 // first candidate of this warp in this pass
        const int c0 = wbase + lane * 4;                           // this thread's candidates c0 .. c0 + 3
        if (wbase < N) {                                           // the rows this warp reads after its RK4 steps -> L2
            const int n_lines = min(4, (Npad - wbase) >> 5);
            const int n_rows = emit ? 3 + G : 3;                   // the group's other columns (+ the partial rows)
            if (lane < n_rows) {
                int row = lane < 3 ? 4 * gnew + lane + (4 * gnew + lane >= slot ? 1 : 0) : W + (lane - 3);
                if (lane >= 3 || row < min(W, 4 * gnew + 4)) prefetch_l2_4lines(err_ring + (size_t)row * Npad + wbase, n_lines);
            }
        }
        u64 key[4] = {~0ull, ~0ull, ~0ull, ~0ull};
        if (c0 < N) {
            HistRow r;
            r.q0 = srow[0]; r.q1 = srow[1]; r.q2 = srow[2]; r.q3 = srow[3]; r.q4 = srow[4];
            float* mine = err_ring + (size_t)slot * Npad + c0;
#if !LLAMPC_RV_PACKED
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {                          // one candidate at a time: 64 registers, 4 CTAs per SM
                if (c0 + q >= N) break;
                const Cand p = load_cand(bank, Npad, c0 + q);
                bool ok;
                float e = lookback_step_fast<GEOM_SHARED, MUFU_SIN>(p, r, z, ok);
                if (!ok) e = lookback_step<GEOM_SHARED, MUFU_SIN>(p, r, z);
                __stcg(mine + q, 0.25f * e);                       // errors of rt.py:349 (mean over the 4 scored states)
            }
#else
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {                          // two packed pairs (c0, c0 + 1), (c0 + 2, c0 + 3)
                const int i0 = min(c0 + 2 * q, N - 1), i1 = min(c0 + 2 * q + 1, N - 1);
                const Cand2 p = load_cand2(bank, Npad, i0, i1);
                bool ok0, ok1;
                const F2 e = lookback_step_fast2<GEOM_SHARED, MUFU_SIN>(p, r, z, ok0, ok1);
                float e0, e1;
                up(e, e0, e1);
                if (!ok0) e0 = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, i0, srow, z);
                if (!ok1) e1 = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, i1, srow, z);
                __stcg(reinterpret_cast<float2*>(mine + 2 * q), make_float2(0.25f * e0, 0.25f * e1));
            }
#endif
            // the new column is read back by the thread that just stored it (program order; columns N .. Npad - 1 are padding)
            const float4 enew = __ldcg(reinterpret_cast<const float4*>(mine));
            // partial sum of the new column's group, canonical order ((e_4g + e_4g+1) + e_4g+2) + e_4g+3
            float4 part = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int w = 4 * gnew + i;
                if (w < W) {
                    const float4 c = w == slot ? enew : __ldcg(reinterpret_cast<const float4*>(err_ring + (size_t)w * Npad + c0));
                    if (i == 0) part = c;
                    else { part.x += c.x; part.y += c.y; part.z += c.z; part.w += c.w; }
                }
            }
            __stcg(reinterpret_cast<float4*>(err_ring + (size_t)(W + gnew) * Npad + c0), part);
            if (emit) {
                float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                for (int g0 = 0; g0 < G; ++g0) {                   // partial_0 + partial_1 + ... in order
                    const float4 c = g0 == gnew ? part : __ldcg(reinterpret_cast<const float4*>(err_ring + (size_t)(W + g0) * Npad + c0));
                    if (g0 == 0) sum = c;
                    else { sum.x += c.x; sum.y += c.y; sum.z += c.z; sum.w += c.w; }
                }
                const float fw = (float)W;
                const float err[4] = {sum.x / fw, sum.y / fw, sum.z / fw, sum.w / fw};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (c0 + q < N) {
                        if (avg_err) avg_err[(size_t)v * N + c0 + q] = err[q];
                        key[q] = pack_key(err[q], (unsigned)(idx_offset + c0 + q));
                    }
            }
        }
        if (!emit) continue;                                       // uniform
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            s_key[j * RV_CPP + q * RV_THREADS + tid] = key[q];     // any order: the filter below scans them all
            const u64 gmin = warp_min_key(key[q]);
            if (lane == 0) s_group[(j * RV_WARPS + warp) * 4 + q] = gmin;
        }
    }
    if (!emit) return;
    __syncthreads();
    cta_topk_filter<RV_THREADS>(s_key, passes * RV_CPP, s_group, passes * RV_WARPS * 4, K, s_cand, &s_thr, &s_count,
                                out + (size_t)v * (LLAMPC_LIST_LEN + 1));
}

int launch_k1r(const float4* bank, int N, int Npad, int W, StepSize z, const NewRow& nr, const float* hist, int n_vehicles,
               float* err_ring, float* avg_err, u64* cta_lists, int idx_offset, int emit, const FusedMerge& fm, bool geom,
               bool mufu, cudaStream_t st) {
    const dim3 grid((N + LB_THREADS - 1) / LB_THREADS, n_vehicles);
    auto kern = mufu ? (geom ? lookback_rolling_kernel<true, true> : lookback_rolling_kernel<false, true>)
                     : (geom ? lookback_rolling_kernel<true, false> : lookback_rolling_kernel<false, false>);
    return issue(kern, grid, dim3(LB_THREADS), 0, st, bank, N, Npad, W, z, nr, hist, err_ring, avg_err, cta_lists, idx_offset,
                 emit, fm);
}

int launch_k1v(const float4* bank, int N, int Npad, int W, StepSize z, int slot, const float* hist, int n_vehicles,
               float* err_ring, float* avg_err, int idx_offset, int emit, int K, u64* out, bool geom, bool mufu,
               cudaStream_t st) {
    auto kern = mufu ? (geom ? lookback_rolling_vehicle_kernel<true, true> : lookback_rolling_vehicle_kernel<false, true>)
                     : (geom ? lookback_rolling_vehicle_kernel<true, false> : lookback_rolling_vehicle_kernel<false, false>);
    return issue(kern, dim3(n_vehicles), dim3(RV_THREADS), 0, st, bank, N, Npad, W, z, slot, hist, err_ring, avg_err, idx_offset,
                 emit, K, out);
}

}  // namespace llampc
