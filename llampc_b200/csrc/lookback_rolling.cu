// K1r / K1v: the reference's rolling error-window bookkeeping on the device (rt.py:349-358).  sm_100a.
#include "lookback_kernels.cuh"

namespace llampc {

// ---------------------------------------------------------------------------------------------------
// K1r rolling window (the reference's own bookkeeping, rt.py:349-358): only the newest transition is integrated
// (one RK4 step per candidate), its error replaces ring column `slot` of err_ring [W][Npad] (np.roll + write of
// the last column), and the window mean is re-summed from the ring -- N steps and N*W*4 bytes per tick instead
// of N*W steps.  emit = 0 while the window is filling (columns stored, no decision).
// ---------------------------------------------------------------------------------------------------
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
    err_ring += (size_t)v * W * Npad;
    bool ok;
    float e = lookback_step_fast<GEOM_SHARED, MUFU_SIN>(p, r, z, ok);
    if (!ok) e = lookback_step<GEOM_SHARED, MUFU_SIN>(p, r, z);
    e *= 0.25f;                                    // errors of rt.py:349 (mean over the 4 scored states)
    if (valid) err_ring[(size_t)nr.slot * Npad + cand] = e;
    if (!emit) return;                             // uniform
    // window re-sum in ring order (deterministic); the loads of 8 columns are issued before the first add so that
    // enough bytes are in flight per SM for HBM (the ring of 4,096 vehicles is 335 MB: this kernel is HBM-bound there)
    float sum = 0.0f;
    const float* col = err_ring + ci;
    int w = 0;
    for (; w + 8 <= W; w += 8) {
        float vq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) vq[j] = __ldcg(col + (size_t)(w + j) * Npad);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += (w + j == nr.slot) ? e : vq[j];
    }
    for (; w < W; ++w) sum += (w == nr.slot) ? e : __ldcg(col + (size_t)w * Npad);
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
// each).  Same arithmetic as K1r (one RK4 step per candidate, ring column `slot` replaced, window re-summed in ring
// order: scores bit-identical to K1r); what changes is everything around the step.  ncu on K1r at 4,096 x 1,024 x 20:
// ~1,200 instructions per candidate-tick, ~800 of them selection (a 15-stage register bitonic network per warp, three
// merges per CTA, a list round trip through L2, a last-CTA merge per vehicle: ALU pipe 48 %, FMA 27 %, HBM 19 %).
//   * Every thread owns FOUR ADJACENT candidates: the ring is re-summed with LDG.128 (5 loads and 10 address
//     instructions per candidate instead of 20 and 40), the new column is read back by the thread that stored it.
//   * The vehicle's keys stay in shared memory and are FILTERED, not sorted (cta_topk_filter, lookback_kernels.cuh):
//     every group of 32 keys (warp x candidate slot) leaves its minimum (two REDUX), <= 64 minima per vehicle; the K-th
//     smallest minimum is a threshold that at least K keys pass; the survivors (K .. 3K keys) are sorted by warp 0.
//   * The ring lines a warp will re-sum are prefetched into L2 before its RK4 steps (lane w asks for row w).
// Needs Npad % 4 == 0 and a 16-byte aligned ring (the entry point falls back to K1r otherwise).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2_4lines(const void* p, int n_lines) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    if (n_lines > 1) asm volatile("prefetch.global.L2 [%0+128];" ::"l"(p));
    if (n_lines > 2) asm volatile("prefetch.global.L2 [%0+256];" ::"l"(p));
    if (n_lines > 3) asm volatile("prefetch.global.L2 [%0+384];" ::"l"(p));
}

#ifndef LLAMPC_RV_MIN_BLOCKS
#define LLAMPC_RV_MIN_BLOCKS 4
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
    if (tid < 5) srow[tid] = __ldg(reinterpret_cast<const float4*>(hist + ((size_t)v * W + slot) * LLAMPC_HIST_ROW) + tid);
    err_ring += (size_t)v * W * Npad;
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < passes; ++j) {
        const int wbase = j * RV_CPP + warp * 128;                 // first candidate of this warp in this pass
        const int c0 = wbase + lane * 4;                           // this thread's candidates c0 .. c0 + 3
        if (emit && wbase < N) {
            const int n_lines = min(4, (Npad - wbase) >> 5);
            for (int w = lane; w < W; w += 32)
                if (w != slot) prefetch_l2_4lines(err_ring + (size_t)w * Npad + wbase, n_lines);
        }
        u64 key[4] = {~0ull, ~0ull, ~0ull, ~0ull};
        if (c0 < N) {
            float* mine = err_ring + (size_t)slot * Npad + c0;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                if (c0 + q >= N) break;
                const Cand p = load_cand(bank, Npad, c0 + q);
                HistRow r;
                r.q0 = srow[0]; r.q1 = srow[1]; r.q2 = srow[2]; r.q3 = srow[3]; r.q4 = srow[4];
                bool ok;
                float e = lookback_step_fast<GEOM_SHARED, MUFU_SIN>(p, r, z, ok);
                if (!ok) e = lookback_step<GEOM_SHARED, MUFU_SIN>(p, r, z);
                __stcg(mine + q, 0.25f * e);                       // errors of rt.py:349 (mean over the 4 scored states)
            }
            if (emit) {
                // window re-sum in ring order, as K1r; the new column is read back by the thread that just stored it
                // (program order), so the loop is loads and adds only.  Columns N .. Npad - 1 are padding.
                float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                const float* col = err_ring + c0;
                int w = 0;
                for (; w + 4 <= W; w += 4, col += (size_t)4 * Npad) {
                    float4 vq[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) vq[i] = __ldcg(reinterpret_cast<const float4*>(col + (size_t)i * Npad));
#pragma unroll
                    for (int i = 0; i < 4; ++i) { sum.x += vq[i].x; sum.y += vq[i].y; sum.z += vq[i].z; sum.w += vq[i].w; }
                }
                for (; w < W; ++w, col += Npad) {
                    const float4 vq = __ldcg(reinterpret_cast<const float4*>(col));
                    sum.x += vq.x; sum.y += vq.y; sum.z += vq.z; sum.w += vq.w;
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
