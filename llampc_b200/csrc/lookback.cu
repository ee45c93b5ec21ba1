// K1 look-back window kernel, K4 top-K, fp64 finalist re-score, one-step batch kernels.  sm_100a.
//
// Reference behaviour being replaced: evaluate_models_vectorized (llampc/mpc/evaluate_models_vectorized.py:4-23)
// + the scoring / selection block of run_nmpc_orca_llampc_rt.py:347-360.
#include "llampc_common.cuh"
#include "llampc_model.cuh"
#include "llampc_model_f64.cuh"
#include "llampc_packed.cuh"
#include "lookback_select.cuh"
#include <math.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

namespace llampc {

constexpr int NUM_SMS = 148;

// ---------------------------------------------------------------------------------------------------
// Launch helper.  Normally issue() is kern<<<...>>>(...).  While llampc_lookback_tick collects (g_collect set), the
// launch is recorded instead -- function, shape and a copy of every argument -- and the tick replays its two kernels
// (scoring + fp64 re-score) as ONE CUDA graph whose kernel nodes are re-parameterised every tick: measured with
// tools/ubench/graph_launch.cu, 2.0 us of host enqueue time instead of 7.6 us and 3.9 us less from enqueue to the
// host-visible result.
// ---------------------------------------------------------------------------------------------------
struct PendingLaunch {
    void* func; dim3 grid, block; size_t smem; void* args[24]; int n_args; size_t used;
    alignas(16) unsigned char store[2048];
};
constexpr int TICK_GRAPH_MAX_NODES = 2;
static thread_local PendingLaunch* g_collect = nullptr;
static thread_local int g_collect_n = 0;

template <class T>
static inline void pending_push(PendingLaunch& pl, const T& v) {
    const size_t off = (pl.used + alignof(T) - 1) & ~(alignof(T) - 1);
    memcpy(pl.store + off, &v, sizeof(T));
    pl.args[pl.n_args++] = pl.store + off;
    pl.used = off + sizeof(T);
}

template <class... KArgs, class... Args>
static int issue(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static_assert(sizeof...(KArgs) == sizeof...(Args), "argument count");
    if (g_collect && g_collect_n < TICK_GRAPH_MAX_NODES) {
        PendingLaunch& pl = g_collect[g_collect_n++];
        pl.func = reinterpret_cast<void*>(kern);
        pl.grid = grid; pl.block = block; pl.smem = smem; pl.n_args = 0; pl.used = 0;
        (pending_push<KArgs>(pl, static_cast<KArgs>(args)), ...);
        return 0;
    }
    kern<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
    return (int)cudaGetLastError();
}

struct TickGraph {
    cudaGraph_t graph; cudaGraphExec_t exec; cudaGraphNode_t node[TICK_GRAPH_MAX_NODES];
    void* func[TICK_GRAPH_MAX_NODES]; int n_nodes;
};

static void tick_graph_destroy(TickGraph* tg) {
    if (!tg) return;
    if (tg->exec) cudaGraphExecDestroy(tg->exec);
    if (tg->graph) cudaGraphDestroy(tg->graph);
    tg->exec = nullptr; tg->graph = nullptr; tg->n_nodes = 0;
}

// replays the collected launches as one graph on `st` (created on first use, re-parameterised afterwards)
static int tick_graph_launch(TickGraph* tg, PendingLaunch* pl, int n, cudaStream_t st) {
    cudaKernelNodeParams kp[TICK_GRAPH_MAX_NODES];
    bool same = tg->exec != nullptr && tg->n_nodes == n;
    for (int i = 0; i < n; ++i) {
        memset(&kp[i], 0, sizeof(kp[i]));
        kp[i].func = pl[i].func; kp[i].gridDim = pl[i].grid; kp[i].blockDim = pl[i].block;
        kp[i].sharedMemBytes = (unsigned)pl[i].smem; kp[i].kernelParams = pl[i].args;
        same = same && tg->func[i] == pl[i].func;
    }
    if (same) {
        for (int i = 0; i < n && same; ++i)
            if (cudaGraphExecKernelNodeSetParams(tg->exec, tg->node[i], &kp[i]) != cudaSuccess) { (void)cudaGetLastError(); same = false; }
    }
    if (!same) {
        tick_graph_destroy(tg);
        LLAMPC_CUDA_TRY(cudaGraphCreate(&tg->graph, 0));
        for (int i = 0; i < n; ++i) {
            LLAMPC_CUDA_TRY(cudaGraphAddKernelNode(&tg->node[i], tg->graph, i ? &tg->node[i - 1] : nullptr, i ? 1 : 0, &kp[i]));
            tg->func[i] = pl[i].func;
        }
        LLAMPC_CUDA_TRY(cudaGraphInstantiate(&tg->exec, tg->graph, 0));
        tg->n_nodes = n;
    }
    return (int)cudaGraphLaunch(tg->exec, st);
}

// ---------------------------------------------------------------------------------------------------
// K1.  grid = (ceil(N / (128/SY)), n_vehicles); block = 128 threads = (128/SY candidates) x (SY window splits).
// The W history rows (80 B each) are staged once per CTA with one TMA bulk copy; every warp then reads
// the same row at the same time (shared-memory broadcast).  Thread (c, sy) integrates window rows
// sy, sy+SY, ...; partial sums are combined in a fixed order so the result is run-to-run deterministic.
// (CTA-level selection, NewRow, FusedMerge, PeerXchg: lookback_select.cuh)
// ---------------------------------------------------------------------------------------------------
template <int SY, bool GEOM_SHARED, bool MUFU_SIN>
__global__ void __launch_bounds__(LB_THREADS, LLAMPC_LB_MIN_BLOCKS)
lookback_window_kernel(const float4* __restrict__ bank, int N, int Npad, const float* __restrict__ hist, int W,
                       long hist_stride_floats, StepSize z, float* __restrict__ avg_err, u64* __restrict__ best_key,
                       u64* __restrict__ cta_lists, int idx_offset, NewRow nr, FusedMerge fm, PeerXchg px, TreeMerge tm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ u64 skeys[LB_THREADS];
    float4* srow = reinterpret_cast<float4*>(smem_raw);
    float* spart = reinterpret_cast<float*>(smem_raw + (size_t)W * (LLAMPC_HIST_ROW * 4));

    constexpr int CPB = LB_THREADS / SY;
    const int tid = threadIdx.x;
    const int v = blockIdx.y;
    const unsigned bytes = (unsigned)W * (LLAMPC_HIST_ROW * 4);

    if (tid == 0) {                                // one thread arms the barrier and starts the bulk copy right away
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, bytes);
        tma_bulk_g2s(srow, hist + (size_t)v * hist_stride_floats, bytes, &mbar);
    }
    __syncthreads();                               // the initialised barrier is visible to the waiting threads

    const int c = tid % CPB, sy = tid / CPB;
    const int cand = blockIdx.x * CPB + c;
    const bool valid = cand < N;
    const Cand p = load_cand(bank, Npad, valid ? cand : N - 1);   // overlaps the bulk copy

    mbar_wait(&mbar, 0);
    if (nr.slot >= 0) {                            // uniform over the grid
        if (tid < LLAMPC_HIST_ROW / 4) {
            const float4 q = make_float4(nr.v[4 * tid], nr.v[4 * tid + 1], nr.v[4 * tid + 2], nr.v[4 * tid + 3]);
            srow[nr.slot * 5 + tid] = q;
            if (blockIdx.x == 0 && blockIdx.y == 0)
                reinterpret_cast<float4*>(const_cast<float*>(hist))[nr.slot * 5 + tid] = q;
        }
        __syncthreads();
    }

    float acc = 0.0f;
    for (int w = sy; w < W; w += SY) {
        HistRow r;
        r.q0 = srow[w * 5 + 0];
        r.q1 = srow[w * 5 + 1];
        r.q2 = srow[w * 5 + 2];
        r.q3 = srow[w * 5 + 3];
        r.q4 = srow[w * 5 + 4];
        bool ok;
        float e = lookback_step_fast<GEOM_SHARED, MUFU_SIN>(p, r, z, ok);
        if (!ok) e = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, valid ? cand : N - 1, srow + w * 5, z);
        acc += e;
    }

    if (SY > 1) {
        spart[sy * CPB + c] = acc;
        __syncthreads();
        if (sy == 0) {
#pragma unroll
            for (int j = 1; j < SY; ++j) acc += spart[j * CPB + c];
        }
    }
    // errors = mean over the 4 scored states (rt.py:349); avg = mean over the window (rt.py:357)
    const float err = acc * (0.25f / (float)W);
    u64 key = ~0ull;
    if (sy == 0 && valid) {
        if (avg_err) avg_err[(size_t)v * N + cand] = err;
        key = pack_key(err, (unsigned)(idx_offset + cand));
    }
    if (tm.K > 0) {                                // uniform over the grid: tree finish (single history), one launch per tick
        __shared__ u64 mrows[BAL_FAN][BAL_ROW_PAD];
        key = cta_select32<(CPB >= 32 ? CPB / 32 : 1)>(key, skeys);
        if (tid < 32) tree_merge(key, tid, (int)blockIdx.x, (int)gridDim.x, tm.K, tm.ws, mrows, tm.out, px);
        return;
    }
    cta_select_emit<(CPB >= 32 ? CPB / 32 : 1)>(key, skeys, v, best_key, cta_lists);   // CPB < 32: part of warp 0 holds keys
    if (fm.K > 0) {                                // uniform over the grid
        __shared__ bool is_last;
        __shared__ MergeSmem<LB_THREADS> msm;
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = atomicAdd(fm.ticket + v, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        u64* outv = fm.out + (size_t)v * (LLAMPC_LIST_LEN + 1);
        merge_lists_device<LB_THREADS>(cta_lists + (size_t)v * gridDim.x * LLAMPC_LIST_LEN, gridDim.x, fm.K,
                                       best_key ? best_key + v : nullptr, outv, msm);
        if (tid == 0) fm.ticket[v] = 0;            // ready for the next launch on the same stream
        if (px.world > 1 && tid < 32) {            // warp 0: min-loc across the GPUs of the box, in this launch
            __syncwarp();
            const u64 mine = __shfl_sync(0xffffffffu, tid == 0 ? *reinterpret_cast<volatile u64*>(outv) : 0ull, 0);
            const u64 g = peer_minloc(px, mine, tid);
            if (tid == 0) outv[0] = g;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K1p.  K1 with TWO candidates per thread in packed f32x2 arithmetic (llampc_packed.cuh): FFMA2 / FMUL2 / FADD2 carry
// both candidates through one issue slot, the history-row values are broadcast scalar operands.  Same tiling as K1
// with twice the candidates per CTA: block = 128 threads = (128/SY candidate pairs) x (SY window splits); thread
// (c, sy) owns candidates base + c and base + 128/SY + c (both bank loads stay coalesced).
// ---------------------------------------------------------------------------------------------------
#ifndef LLAMPC_LB2_MIN_BLOCKS
#define LLAMPC_LB2_MIN_BLOCKS 4
#endif
template <int SY, bool GEOM_SHARED, bool MUFU_SIN>
__global__ void __launch_bounds__(LB_THREADS, LLAMPC_LB2_MIN_BLOCKS)
lookback_window2_kernel(const float4* __restrict__ bank, int N, int Npad, const float* __restrict__ hist, int W,
                        long hist_stride_floats, StepSize z, float* __restrict__ avg_err, u64* __restrict__ best_key,
                        u64* __restrict__ cta_lists, int idx_offset, NewRow nr, FusedMerge fm, PeerXchg px, TreeMerge tm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ u64 skeys[LB_THREADS];
    float4* srow = reinterpret_cast<float4*>(smem_raw);
    float* spart = reinterpret_cast<float*>(smem_raw + (size_t)W * (LLAMPC_HIST_ROW * 4));

    constexpr int CPB = LB_THREADS / SY;           // candidate pairs per CTA
    const int tid = threadIdx.x;
    const int v = blockIdx.y;
    const unsigned bytes = (unsigned)W * (LLAMPC_HIST_ROW * 4);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, bytes);
        tma_bulk_g2s(srow, hist + (size_t)v * hist_stride_floats, bytes, &mbar);
    }
    __syncthreads();

    const int c = tid % CPB, sy = tid / CPB;
    const int cand0 = blockIdx.x * (2 * CPB) + c, cand1 = cand0 + CPB;
    const bool valid0 = cand0 < N, valid1 = cand1 < N;
    const int i0 = valid0 ? cand0 : N - 1, i1 = valid1 ? cand1 : N - 1;
    const Cand2 p = load_cand2(bank, Npad, i0, i1);              // overlaps the bulk copy

    mbar_wait(&mbar, 0);
    if (nr.slot >= 0) {                            // uniform over the grid
        if (tid < LLAMPC_HIST_ROW / 4) {
            const float4 q = make_float4(nr.v[4 * tid], nr.v[4 * tid + 1], nr.v[4 * tid + 2], nr.v[4 * tid + 3]);
            srow[nr.slot * 5 + tid] = q;
            if (blockIdx.x == 0 && blockIdx.y == 0)
                reinterpret_cast<float4*>(const_cast<float*>(hist))[nr.slot * 5 + tid] = q;
        }
        __syncthreads();
    }

    float acc0 = 0.0f, acc1 = 0.0f;
    for (int w = sy; w < W; w += SY) {
        HistRow r;
        r.q0 = srow[w * 5 + 0];
        r.q1 = srow[w * 5 + 1];
        r.q2 = srow[w * 5 + 2];
        r.q3 = srow[w * 5 + 3];
        r.q4 = srow[w * 5 + 4];
        bool ok0, ok1;
        const F2 e = lookback_step_fast2<GEOM_SHARED, MUFU_SIN>(p, r, z, ok0, ok1);
        float e0, e1;
        up(e, e0, e1);
        if (!ok0) e0 = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, i0, srow + w * 5, z);
        if (!ok1) e1 = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, i1, srow + w * 5, z);
        acc0 += e0;
        acc1 += e1;
    }

    if (SY > 1) {
        spart[(sy * 2) * CPB + c] = acc0;
        spart[(sy * 2 + 1) * CPB + c] = acc1;
        __syncthreads();
        if (sy == 0) {
#pragma unroll
            for (int j = 1; j < SY; ++j) {
                acc0 += spart[(j * 2) * CPB + c];
                acc1 += spart[(j * 2 + 1) * CPB + c];
            }
        }
    }
    // errors = mean over the 4 scored states (rt.py:349); avg = mean over the window (rt.py:357)
    const float scale = 0.25f / (float)W;
    const float err0 = acc0 * scale, err1 = acc1 * scale;
    u64 k0 = ~0ull, k1 = ~0ull;
    if (sy == 0) {
        if (valid0) {
            if (avg_err) avg_err[(size_t)v * N + cand0] = err0;
            k0 = pack_key(err0, (unsigned)(idx_offset + cand0));
        }
        if (valid1) {
            if (avg_err) avg_err[(size_t)v * N + cand1] = err1;
            k1 = pack_key(err1, (unsigned)(idx_offset + cand1));
        }
    }
    constexpr int KW = CPB >= 32 ? CPB / 32 : 1;   // CPB < 32: part of warp 0 holds keys
    u64 key = ~0ull;
    if ((tid >> 5) < KW) {                         // two keys per lane -> the 32 smallest of the warp's 64
        const int lane = tid & 31;
        k0 = warp_sort_u64(k0, lane);
        k1 = warp_sort_u64(k1, lane);
        key = warp_merge_low32(k0, __shfl_sync(0xffffffffu, k1, 31 - lane), lane);
    }
    if (tm.K > 0) {                                // uniform over the grid: tree finish (single history), one launch per tick
        __shared__ u64 mrows[BAL_FAN][BAL_ROW_PAD];
        key = cta_select32<KW, true>(key, skeys);
        if (tid < 32) tree_merge(key, tid, (int)blockIdx.x, (int)gridDim.x, tm.K, tm.ws, mrows, tm.out, px);
        return;
    }
    cta_select_emit<KW, true>(key, skeys, v, best_key, cta_lists);
    if (fm.K > 0) {                                // uniform over the grid
        __shared__ bool is_last;
        __shared__ MergeSmem<LB_THREADS> msm;
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = atomicAdd(fm.ticket + v, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        u64* outv = fm.out + (size_t)v * (LLAMPC_LIST_LEN + 1);
        merge_lists_device<LB_THREADS>(cta_lists + (size_t)v * gridDim.x * LLAMPC_LIST_LEN, gridDim.x, fm.K,
                                       best_key ? best_key + v : nullptr, outv, msm);
        if (tid == 0) fm.ticket[v] = 0;
        if (px.world > 1 && tid < 32) {
            __syncwarp();
            const u64 mine = __shfl_sync(0xffffffffu, tid == 0 ? *reinterpret_cast<volatile u64*>(outv) : 0ull, 0);
            const u64 g = peer_minloc(px, mine, tid);
            if (tid == 0) outv[0] = g;
        }
    }
}

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
                        u64* __restrict__ best_key, u64* __restrict__ cta_lists, int idx_offset, int emit, FusedMerge fm) {
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
    cta_select_emit<LB_THREADS / 32>(key, skeys, v, best_key, cta_lists);
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
                                       best_key ? best_key + v : nullptr, fm.out + (size_t)v * (LLAMPC_LIST_LEN + 1), msm);
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
//   * The vehicle's keys stay in shared memory and are FILTERED, not sorted:
//       1  every group of 32 keys (warp x candidate slot) leaves its minimum (two REDUX): <= 64 minima per vehicle;
//       2  warp 0 sorts the minima; T = the K-th smallest.  At least K keys are <= T (those minima themselves), so
//          the top-K is a subset of {key <= T} -- about K .. 3K keys of the N;
//       3  the survivors are compacted with a shared-memory counter and warp 0 sorts them 32 at a time (sort +
//          bitonic merge into the running 32 smallest).  Keys are unique (index in the low word), so the result
//          does not depend on the compaction order.
//   * The ring lines a warp will re-sum are prefetched into L2 before its RK4 steps (lane w asks for row w).
// Needs Npad % 4 == 0 and a 16-byte aligned ring (the entry point falls back to K1r otherwise).
// ---------------------------------------------------------------------------------------------------
constexpr int RV_THREADS = 256;
constexpr int RV_WARPS = RV_THREADS / 32;
constexpr int RV_CPP = RV_THREADS * 4;             // candidates per pass
constexpr int RV_MAX_PASSES = 2;
constexpr int RV_MAX_N = RV_CPP * RV_MAX_PASSES;   // 2,048

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
    if (warp == 0) {                                               // threshold = K-th smallest group minimum
        const int ng = passes * RV_WARPS * 4;
        u64 a = lane < ng ? s_group[lane] : ~0ull;
        a = warp_sort_u64(a, lane);
        if (ng > 32) {
            u64 b = lane + 32 < ng ? s_group[lane + 32] : ~0ull;
            b = warp_sort_u64(b, lane);
            a = warp_merge_low32(a, __shfl_sync(0xffffffffu, b, 31 - lane), lane);
        }
        const u64 t = __shfl_sync(0xffffffffu, a, K - 1);          // ~0 when fewer than K groups hold a key: keep all
        if (lane == 0) { s_thr = t; s_count = 0; }
    }
    __syncthreads();
    const u64 T = s_thr;
    for (int i = tid; i < passes * RV_CPP; i += RV_THREADS) {
        const u64 k = s_key[i];
        if (k <= T && k != ~0ull) s_cand[atomicAdd(&s_count, 1)] = k;
    }
    __syncthreads();
    if (warp != 0) return;
    const int n = s_count;
    u64 run = ~0ull;
    for (int c = 0; c < n; c += 32) {
        u64 k = c + lane < n ? s_cand[c + lane] : ~0ull;
        k = warp_sort_u64(k, lane);
        run = c == 0 ? k : warp_merge_low32(run, __shfl_sync(0xffffffffu, k, 31 - lane), lane);
    }
    u64* o = out + (size_t)v * (LLAMPC_LIST_LEN + 1);              // out[0] = arg-min key, out[1..K] = ascending top-K
    if (lane == 0) o[0] = run;
    if (lane < K) o[1 + lane] = run;
}

// ---------------------------------------------------------------------------------------------------
// K4' stand-alone merge of the per-CTA sorted lists written by K1 / K1r into the global top-K (K <= LLAMPC_LIST_LEN),
// one CTA per vehicle; the algorithm is merge_lists_device (llampc_common.cuh).  Used when a launch produces more
// than 1,024 lists (otherwise the last CTA of K1 runs the same routine itself); optionally carries the NVLink
// min-loc exchange.  out[0] = *best_key (then re-armed to ~0 for the next tick), out[1..K] = ascending top-K.
// ---------------------------------------------------------------------------------------------------
template <int MERGE_THREADS>
__global__ void __launch_bounds__(MERGE_THREADS)
topk_merge_lists_kernel(const u64* __restrict__ lists, int n_lists, int K, u64* __restrict__ best_key,
                        u64* __restrict__ out, PeerXchg px) {
    __shared__ MergeSmem<MERGE_THREADS> sm;
    const int v = blockIdx.x;                      // vehicle
    u64* outv = out + (size_t)v * (LLAMPC_LIST_LEN + 1);
    merge_lists_device<MERGE_THREADS>(lists + (size_t)v * n_lists * LLAMPC_LIST_LEN, n_lists, K,
                                      best_key ? best_key + v : nullptr, outv, sm);
    if (px.world > 1 && threadIdx.x < 32) {        // warp 0: min-loc across the GPUs of the box (single vehicle)
        __syncwarp();
        const u64 mine = __shfl_sync(0xffffffffu, threadIdx.x == 0 ? *reinterpret_cast<volatile u64*>(outv) : 0ull, 0);
        const u64 g = peer_minloc(px, mine, threadIdx.x);
        if (threadIdx.x == 0) outv[0] = g;
    }
}

__global__ void fill_keys_kernel(u64* keys, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ~0ull;
}

// ---------------------------------------------------------------------------------------------------
// K4 top-K.  Every CTA extracts the K smallest keys of its slice by K rounds of "block-min of the keys
// strictly above the previous winner" (keys are unique: the low word is the index); the last CTA to
// finish (atomic ticket) merges the per-CTA lists the same way.  Ascending output.
// ---------------------------------------------------------------------------------------------------
constexpr int TK_THREADS = 256;

template <class KeyAt>
__device__ __forceinline__ void select_k_smallest(KeyAt key_at, int n_elems, int K, u64* out, u64* sbuf) {
    // thread-local minimum over its strided elements that are > lower (or all of them on the first round)
    auto scan = [&](bool first, u64 lower) {
        u64 m = ~0ull;
        for (int e = threadIdx.x; e < n_elems; e += TK_THREADS) {
            u64 k = key_at(e);
            if ((first || k > lower) && k < m) m = k;
        }
        return m;
    };
    u64 mine = scan(true, 0);
    for (int r = 0; r < K; ++r) {
        u64 sel = block_min_u64<TK_THREADS / 32>(mine, sbuf);
        if (threadIdx.x == 0) out[r] = sel;
        if (sel == ~0ull) continue;               // fewer than K elements: pad with ~0
        if (mine == sel) mine = scan(false, sel);
    }
}

__global__ void __launch_bounds__(TK_THREADS)
topk_kernel(const float* __restrict__ err, int N, int idx_offset, int K, int per_cta, u64* __restrict__ scratch,
            unsigned* __restrict__ counter, u64* __restrict__ out) {
    __shared__ u64 sbuf[TK_THREADS / 32 + 1];
    __shared__ bool is_last;
    const int base = blockIdx.x * per_cta;
    const int n = min(per_cta, N - base);
    select_k_smallest(
        [&](int e) { return pack_key(__ldg(err + base + e), (unsigned)(idx_offset + base + e)); }, n, K,
        scratch + (size_t)blockIdx.x * K, sbuf);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int total = gridDim.x * K;
    select_k_smallest([&](int e) { return __ldcg(scratch + e); }, total, K, out, sbuf);
    if (threadIdx.x == 0) *counter = 0;           // ready for the next launch on the same stream
}

// ---------------------------------------------------------------------------------------------------
// fp64 re-score of finalists: one warp per finalist, lanes stride over the window.
// ---------------------------------------------------------------------------------------------------
struct NewRow64 { double v[LLAMPC_HIST64_ROW]; int slot; };

// Optional zero-copy hand-off of the tick result: the last re-score block to retire copies `words` result words to
// mapped pinned host memory and then publishes `seq` in the word after them; the host polls that word instead of
// paying for a D2H copy launch plus a stream synchronisation.
struct FinalCopy { const u64* src; volatile u64* dst_host; unsigned* ticket; int words; u64 seq; };

// Optional multi-GPU finalist all-gather over NVLink peer memory, carried by the last re-score block: every rank
// owns a symmetric buffer [2 parities][world][wpr] (wpr = 2 Kt + 1: Kt keys, Kt fp64 scores, sequence word); the block
// stores its finalists into slot `rank` of every peer, waits until its own buffer holds all `world` contributions of
// this tick, and hands the whole set (world * 2 Kt words after the local arg-min key) to the host.
struct PeerGather { u64* const* peers; int world; int rank; unsigned seq; int kt; };
constexpr u64 PEER_FAIL_BIT = 1ull << 63;         // set in the published sequence word when a peer timed out

constexpr int RF_THREADS = 128;                   // 64 window rows at a time x 2 lanes (front / rear tyre) per row

__global__ void __launch_bounds__(RF_THREADS)
refine_f64_kernel(const double* __restrict__ bank64, int N, double* __restrict__ hist64, int W, double h,
                  const u64* __restrict__ keys, int idx_offset, double* __restrict__ out, NewRow64 nr, FinalCopy fc,
                  PeerGather pg) {
    __shared__ double spart[RF_THREADS / 32];
    __shared__ bool last_block;
    __shared__ int peer_fail;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    if (nr.slot >= 0 && f == 0 && tid < LLAMPC_HIST64_ROW) hist64[(size_t)nr.slot * LLAMPC_HIST64_ROW + tid] = nr.v[tid];
    const long long ci = (long long)(unsigned)(keys[f] & 0xffffffffull) - idx_offset;
    double result = __longlong_as_double(0x7ff8000000000000ll);       // padded key (~0) or foreign shard -> NaN
    if (ci >= 0 && ci < N) {                       // uniform over the block
        Params64 p;
        double* pp = reinterpret_cast<double*>(&p);
#pragma unroll
        for (int j = 0; j < LLAMPC_NPARAM; ++j) pp[j] = bank64[(size_t)j * N + ci];
        double acc = 0.0;
        const bool rear = tid & 1;                 // lane pair (2k, 2k+1) shares window row w
        for (int w0 = 0; w0 < W; w0 += RF_THREADS / 2) {      // uniform trip count: the pair shuffles need full warps
            const int w = w0 + (tid >> 1);
            const int wc = w < W ? w : W - 1;
            double r[LLAMPC_HIST64_ROW];
#pragma unroll
            for (int i = 0; i < LLAMPC_HIST64_ROW; ++i)
                r[i] = (wc == nr.slot) ? nr.v[i] : hist64[(size_t)wc * LLAMPC_HIST64_ROW + i];
            double y1[6];
            rk4_step64_pair(p, r, r[6], r[7], h, rear, y1);
            double e = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { double d = y1[i] - r[8 + i]; e += d * d; }
            if (w < W && !rear) acc += e / 4;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) spart[tid >> 5] = acc;
        __syncthreads();
        result = ((spart[0] + spart[1]) + (spart[2] + spart[3])) / W;
    }
    if (tid == 0) out[f] = result;
    if (fc.dst_host) {                             // uniform over the grid
        __threadfence();
        __syncthreads();
        if (tid == 0) last_block = atomicAdd(fc.ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!last_block) return;
        __threadfence();
        int words = fc.words;
        if (pg.world > 1) {
            // fc.src = [arg-min key | Kt keys | Kt scores]; exchange the 2 Kt finalist words with every peer
            const int wpr = 2 * pg.kt + 1, parity = pg.seq & 1;
            const size_t my_slot = ((size_t)parity * pg.world + pg.rank) * wpr;
            for (int i = tid; i < 2 * pg.kt * pg.world; i += RF_THREADS) {
                const int q = i / (2 * pg.kt), j = i % (2 * pg.kt);
                reinterpret_cast<volatile u64*>(pg.peers[q])[my_slot + j] = __ldcg(fc.src + 1 + j);
            }
            __threadfence_system();
            __syncthreads();
            if (tid == 0) peer_fail = 0;
            __syncthreads();
            if (tid < pg.world) {
                reinterpret_cast<volatile u64*>(pg.peers[tid])[my_slot + 2 * pg.kt] = (u64)pg.seq;
                volatile u64* own = pg.peers[pg.rank] + ((size_t)parity * pg.world + tid) * wpr;
                const long long t0 = clock64();
                while (own[2 * pg.kt] != (u64)pg.seq) {
                    if (clock64() - t0 > 2000000000ll) { peer_fail = 1; break; }   // ~1 s: the peer never arrived
                    __nanosleep(64);
                }
            }
            __threadfence_system();
            __syncthreads();
            fc.dst_host[0] = __ldcg(fc.src);
            volatile u64* own = pg.peers[pg.rank] + (size_t)parity * pg.world * wpr;
            const bool failed = peer_fail != 0;                    // poison: keys ~0, scores NaN, and the flag below
            for (int i = tid; i < 2 * pg.kt * pg.world; i += RF_THREADS) {
                const int q = i / (2 * pg.kt), j = i % (2 * pg.kt);
                fc.dst_host[1 + i] = failed ? ~0ull : own[(size_t)q * wpr + j];
            }
            words = 1 + 2 * pg.kt * pg.world;
        } else {
            for (int i = tid; i < words; i += RF_THREADS) fc.dst_host[i] = __ldcg(fc.src + i);
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            *fc.ticket = 0;
            // a peer that never arrived is reported in the sequence word (llampc_lookback_finish -> LLAMPC_E_PEER)
            fc.dst_host[words] = (pg.world > 1 && peer_fail) ? (fc.seq | PEER_FAIL_BIT) : fc.seq;
            __threadfence_system();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// One RK4 step / one RHS evaluation for N (model, state, input) triples, f64 in/out, fp32 increments.
// ---------------------------------------------------------------------------------------------------
template <int MODE>   // 0: RK4 step, 1: right-hand side, 2: forces and slip angles
__global__ void __launch_bounds__(128)
onestep_kernel(const float4* __restrict__ bank, int N, int Npad, const double* __restrict__ x64, int x_shared,
               const double* __restrict__ u64v, int u_shared, float h, double* __restrict__ out, int out_cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const Cand p = load_cand(bank, Npad, i);
    const double* x = x64 + (x_shared ? 0 : (size_t)i * 6);
    const double* u = u64v + (u_shared ? 0 : (size_t)i * 2);
    double s0d, c0d, sdd, cdd;
    sincos(x[2], &s0d, &c0d);
    sincos(u[1], &sdd, &cdd);
    Ctl ctl;
    ctl.pwm = (float)u[0]; ctl.delta = (float)u[1]; ctl.sd = (float)sdd; ctl.cd = (float)cdd;
    const float s0 = (float)s0d, c0 = (float)c0d, vx = (float)x[3], vy = (float)x[4], w = (float)x[5];
    if (MODE == 2) {                               // calc_forces_batch(..., return_slip=True), dynamic.py:117-154
        float af, ar;
        slip_angles(p, ctl.delta, vx, vy, w, af, ar);
        double* o = out + (size_t)i * 5;
        o[0] = (double)pacejka<false>(p.Bf, p.Cf, p.Df, af);
        o[1] = (double)drive_force(p, ctl.pwm, vx);
        o[2] = (double)pacejka<false>(p.Br, p.Cr, p.Dr, ar);
        o[3] = (double)af; o[4] = (double)ar;
    } else if (MODE == 1) {
        Deriv a = accel<false>(p, ctl, vx, vy, w);
        double* o = out + (size_t)i * 6;
        o[0] = (double)fmaf(vx, c0, -vy * s0);
        o[1] = (double)fmaf(vx, s0, vy * c0);
        o[2] = x[5];
        o[3] = (double)a.vx; o[4] = (double)a.vy; o[5] = (double)a.w;
    } else {
        float inc[6];
        rk4_increment<false>(p, ctl, s0, c0, vx, vy, w, h, inc);
        double* o = out + (size_t)i * out_cols;
        for (int j = 0; j < out_cols; ++j) o[j] = x[j] + (double)inc[j];
    }
}

}  // namespace llampc

using namespace llampc;

// ===================================================================================================
// C ABI
// ===================================================================================================
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// K1p (two candidates per thread, packed f32x2) is the default for banks that can fill the GPU with pair-threads;
// small banks keep one candidate per thread (twice the threads: measured 12.4 us against 14.2 us per tick at
// 1,024 x 20).  LLAMPC_K1_PACKED=0 / 1 forces the scalar / packed kernel for every size.
constexpr int K1P_MIN_CANDIDATES = 8192;
static bool k1_packed(int N) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LLAMPC_K1_PACKED");
        v = !e ? 2 : (e[0] == '0' ? 0 : 1);
    }
    return v == 2 ? N >= K1P_MIN_CANDIDATES : v != 0;
}
static inline int k1_cands_per_cta(int N) { return k1_packed(N) ? 2 * LB_THREADS : LB_THREADS; }

static int choose_split(int N, int W) {
    // SM time ~ (CTAs on the busiest SM) x (rows per thread) while the FMA pipe is the limiter.
    int best = 1;
    long best_cost = -1;
    for (int sy = 1; sy <= 16; sy *= 2) {          // 8 and 16 only pay off for banks too small to fill the GPU
        if (sy > W) break;
        long ctas = ((long)N * sy + k1_cands_per_cta(N) - 1) / k1_cands_per_cta(N);
        long cost = ((ctas + NUM_SMS - 1) / NUM_SMS) * ((W + sy - 1) / sy);
        // a finer split must win by > 3 %: it doubles the number of per-CTA lists the top-K merge has to read
        if (best_cost < 0 || cost * 100 < best_cost * 97) { best_cost = cost; best = sy; }
    }
    return best;
}

template <int SY, bool GEOM, bool MUFU>
static int launch_lookback(const float* bank, int N, int Npad, const float* hist, int W, int n_vehicles,
                           int hist_stride_rows, double Ts, float* avg_err, u64* best_key, u64* cta_lists,
                           int idx_offset, const NewRow& nr, const FusedMerge& fm, const PeerXchg& px, const TreeMerge& tm,
                           cudaStream_t st) {
    const bool packed = k1_packed(N);
    auto kern = packed ? lookback_window2_kernel<SY, GEOM, MUFU> : lookback_window_kernel<SY, GEOM, MUFU>;
    const size_t smem = (size_t)W * (LLAMPC_HIST_ROW * 4) + (SY > 1 ? LB_THREADS * 4 * (packed ? 2 : 1) : 0);
    if (smem > 48 * 1024) {
        static bool raised[2] = {false, false};  // per kernel (scalar / packed); idempotent attribute, benign if two threads race
        if (!raised[packed]) {
            LLAMPC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            raised[packed] = true;
        }
    }
    const int CPB = (packed ? 2 * LB_THREADS : LB_THREADS) / SY;
    dim3 grid((N + CPB - 1) / CPB, n_vehicles);
    return issue(kern, grid, dim3(LB_THREADS), smem, st, reinterpret_cast<const float4*>(bank), N, Npad, hist, W,
                 (long)hist_stride_rows * LLAMPC_HIST_ROW, make_step(Ts), avg_err, best_key, cta_lists, idx_offset, nr, fm, px,
                 tm);
}

static int lookback_window_impl(const float* bank, int N, int Npad, const float* hist, int W, int n_vehicles,
                                int hist_stride_rows, double Ts, float* avg_err, llampc_key_t* best_key,
                                llampc_key_t* cta_lists, int idx_offset, int geom_shared, int split,
                                const NewRow& nr, const FusedMerge& fm, llampc_stream_t stream,
                                const PeerXchg& px = PeerXchg{nullptr, 0, 0, 0},
                                const TreeMerge& tm = TreeMerge{{nullptr, nullptr, nullptr, nullptr, nullptr}, nullptr, 0}) {
    if (!bank || !hist || (!best_key && !cta_lists && !avg_err && tm.K <= 0) || N <= 0 || Npad < N || n_vehicles <= 0 || hist_stride_rows < W) return LLAMPC_E_ARG;
    if (W <= 0 || W > LLAMPC_MAX_W || n_vehicles > 65535) return LLAMPC_E_RANGE;
    if (!aligned16(bank) || !aligned16(hist) || (hist_stride_rows * LLAMPC_HIST_ROW * 4) % 16) return LLAMPC_E_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int mufu = 0;
    if (split >= 32) { mufu = 1; split -= 32; }   // bit 5: MUFU.SIN tyre sine (LookBack(fast_sin=True))
    if (split == 0) split = choose_split(N, W);
    if (split > W) split = 1;
#define LB_CASE(SYV)                                                                                                  \
    case SYV:                                                                                                         \
        if (mufu) return geom_shared ? launch_lookback<SYV, true, true>(bank, N, Npad, hist, W, n_vehicles, hist_stride_rows, Ts, avg_err, best_key, cta_lists, idx_offset, nr, fm, px, tm, st)   \
                                     : launch_lookback<SYV, false, true>(bank, N, Npad, hist, W, n_vehicles, hist_stride_rows, Ts, avg_err, best_key, cta_lists, idx_offset, nr, fm, px, tm, st); \
        return geom_shared ? launch_lookback<SYV, true, false>(bank, N, Npad, hist, W, n_vehicles, hist_stride_rows, Ts, avg_err, best_key, cta_lists, idx_offset, nr, fm, px, tm, st)            \
                           : launch_lookback<SYV, false, false>(bank, N, Npad, hist, W, n_vehicles, hist_stride_rows, Ts, avg_err, best_key, cta_lists, idx_offset, nr, fm, px, tm, st);
    switch (split) {
        LB_CASE(1)
        LB_CASE(2)
        LB_CASE(4)
        LB_CASE(8)
        LB_CASE(16)
        default: return LLAMPC_E_ARG;
    }
#undef LB_CASE
}

extern "C" int llampc_lookback_window_f32(const float* bank, int N, int Npad, const float* hist, int W,
                                          int n_vehicles, int hist_stride_rows, double Ts, float* avg_err,
                                          llampc_key_t* best_key, llampc_key_t* cta_lists, int idx_offset,
                                          int geom_shared, int split, llampc_stream_t stream) {
    NewRow nr;
    nr.slot = -1;
    FusedMerge fm = {nullptr, nullptr, 0};
    return lookback_window_impl(bank, N, Npad, hist, W, n_vehicles, hist_stride_rows, Ts, avg_err, best_key, cta_lists,
                                idx_offset, geom_shared, split, nr, fm, stream);
}

extern "C" int llampc_lookback_window_topk_f32(const float* bank, int N, int Npad, const float* hist, int W,
                                               int n_vehicles, int hist_stride_rows, double Ts, float* avg_err,
                                               llampc_key_t* best_key, llampc_key_t* cta_lists, int idx_offset,
                                               int geom_shared, int split, int K, unsigned* ticket, llampc_key_t* out,
                                               llampc_stream_t stream) {
    if (!cta_lists || !out || !ticket || !best_key) return LLAMPC_E_ARG;
    if (K <= 0 || K > LLAMPC_LIST_LEN) return LLAMPC_E_RANGE;
    const int n_lists = llampc_lookback_num_lists(N, W, split);
    if (n_lists <= 0) return LLAMPC_E_ARG;
    NewRow nr;
    nr.slot = -1;
    if (n_lists > LB_THREADS * MERGE_LPT) {          // too many lists for one CTA: K1, then the stand-alone merge kernel
        FusedMerge none = {nullptr, nullptr, 0};
        int rc = lookback_window_impl(bank, N, Npad, hist, W, n_vehicles, hist_stride_rows, Ts, avg_err, best_key, cta_lists,
                                      idx_offset, geom_shared, split, nr, none, stream);
        if (rc) return rc;
        return llampc_topk_merge_lists(cta_lists, n_lists, n_vehicles, K, best_key, out, stream);
    }
    FusedMerge fm = {ticket, out, K};
    return lookback_window_impl(bank, N, Npad, hist, W, n_vehicles, hist_stride_rows, Ts, avg_err, best_key, cta_lists,
                                idx_offset, geom_shared, split, nr, fm, stream);
}

static int lookback_rolling_impl(const float* bank, int N, int Npad, const float* row32_h, const float* hist, int n_vehicles,
                                 int slot, int W, double Ts, float* err_ring, float* avg_err, llampc_key_t* best_key,
                                 llampc_key_t* cta_lists, int idx_offset, int geom_shared, int emit, const FusedMerge& fm,
                                 llampc_stream_t stream) {
    if (!bank || (!row32_h && !hist) || !err_ring || N <= 0 || Npad < N || slot < 0 || slot >= W || n_vehicles <= 0)
        return LLAMPC_E_ARG;
    if (W <= 0 || W > LLAMPC_MAX_W || n_vehicles > 65535) return LLAMPC_E_RANGE;
    if (!aligned16(bank) || (hist && !aligned16(hist))) return LLAMPC_E_ALIGN;
    NewRow nr;
    if (!hist)
        for (int i = 0; i < LLAMPC_HIST_ROW; ++i) nr.v[i] = row32_h[i];
    nr.slot = slot;
    const dim3 grid((N + LB_THREADS - 1) / LB_THREADS, n_vehicles);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the single loop keeps the polynomial tyre sine (its tick is latency-bound anyway); the Monte-Carlo layout runs the
    // SFU sine of the default K1 tick (MUFU.SIN: scores within 2.2e-5 of the oracle)
    // (bit 1 of geom_shared asks for the polynomial sine there as well: wide banks, see include/llampc_b200.h)
    const bool geom = geom_shared & 1, mufu = hist && !(geom_shared & 2);
    auto kern = mufu ? (geom ? lookback_rolling_kernel<true, true> : lookback_rolling_kernel<false, true>)
                     : (geom ? lookback_rolling_kernel<true, false> : lookback_rolling_kernel<false, false>);
    return issue(kern, grid, dim3(LB_THREADS), 0, st, reinterpret_cast<const float4*>(bank), N, Npad, W, make_step(Ts), nr,
                 hist, err_ring, avg_err, best_key, cta_lists, idx_offset, emit, fm);
}

static int merge_lists_impl(const llampc_key_t* cta_lists, int n_lists, int n_vehicles, int K, llampc_key_t* best_key,
                            llampc_key_t* out, const PeerXchg& px, llampc_stream_t stream);

extern "C" int llampc_lookback_window_topk_peer_f32(const float* bank, int N, int Npad, const float* hist, int W,
                                                    int hist_stride_rows, double Ts, float* avg_err,
                                                    llampc_key_t* best_key, llampc_key_t* cta_lists, int idx_offset,
                                                    int geom_shared, int split, int K, unsigned* ticket,
                                                    llampc_key_t* out, llampc_key_t* const* peer_bufs, int world,
                                                    int rank, unsigned seq, llampc_stream_t stream) {
    if (!cta_lists || !out || !ticket || !best_key || !peer_bufs) return LLAMPC_E_ARG;
    if (K <= 0 || K > LLAMPC_LIST_LEN || world < 2 || world > 32 || rank < 0 || rank >= world) return LLAMPC_E_RANGE;
    const int n_lists = llampc_lookback_num_lists(N, W, split);
    if (n_lists <= 0) return LLAMPC_E_ARG;
    NewRow nr;
    nr.slot = -1;
    PeerXchg px = {peer_bufs, world, rank, seq};
    if (n_lists > LB_THREADS * MERGE_LPT) {          // big shard: K1, then the merge kernel carries the exchange
        FusedMerge none = {nullptr, nullptr, 0};
        int rc = lookback_window_impl(bank, N, Npad, hist, W, 1, hist_stride_rows, Ts, avg_err, best_key, cta_lists,
                                      idx_offset, geom_shared, split, nr, none, stream);
        if (rc) return rc;
        return merge_lists_impl(cta_lists, n_lists, 1, K, best_key, out, px, stream);
    }
    FusedMerge fm = {ticket, out, K};
    return lookback_window_impl(bank, N, Npad, hist, W, 1, hist_stride_rows, Ts, avg_err, best_key, cta_lists, idx_offset,
                                geom_shared, split, nr, fm, stream, px);
}

extern "C" int llampc_lookback_rolling_f32(const float* bank, int N, int Npad, const float* row32_h, int slot, int W,
                                           double Ts, float* err_ring, float* avg_err, llampc_key_t* best_key,
                                           llampc_key_t* cta_lists, int idx_offset, int geom_shared, int emit,
                                           llampc_stream_t stream) {
    FusedMerge none = {nullptr, nullptr, 0};
    return lookback_rolling_impl(bank, N, Npad, row32_h, nullptr, 1, slot, W, Ts, err_ring, avg_err, best_key, cta_lists,
                                 idx_offset, geom_shared, emit, none, stream);
}

extern "C" int llampc_lookback_rolling_multi_f32(const float* bank, int N, int Npad, const float* hist, int n_vehicles,
                                                 int slot, int W, double Ts, float* err_ring, float* avg_err,
                                                 llampc_key_t* best_key, llampc_key_t* cta_lists, int idx_offset,
                                                 int geom_shared, int emit, int K, unsigned* ticket, llampc_key_t* out,
                                                 llampc_stream_t stream) {
    if (!hist) return LLAMPC_E_ARG;
    if (emit && (K < 0 || K > LLAMPC_LIST_LEN)) return LLAMPC_E_RANGE;
    // banks of <= 2,048 candidates: one CTA per vehicle, top-K by threshold filter in shared memory (K1v; scores and
    // keys identical to the K1r path below; best_key / cta_lists / ticket are not touched).  LLAMPC_K1R_CTA=0 keeps K1r.
    const char* sw = getenv("LLAMPC_K1R_CTA");                   // read per call: the parity test flips it
    if (N > 0 && N <= RV_MAX_N && Npad % 4 == 0 && aligned16(err_ring) && (!emit || (K > 0 && out)) && !(sw && sw[0] == '0')) {
        if (!bank || !err_ring || Npad < N || slot < 0 || slot >= W || n_vehicles <= 0) return LLAMPC_E_ARG;
        if (W <= 0 || W > LLAMPC_MAX_W || n_vehicles > 65535) return LLAMPC_E_RANGE;
        if (!aligned16(bank) || !aligned16(hist)) return LLAMPC_E_ALIGN;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const bool geom = geom_shared & 1, strict = geom_shared & 2;
        auto kern = geom ? (strict ? lookback_rolling_vehicle_kernel<true, false> : lookback_rolling_vehicle_kernel<true, true>)
                         : (strict ? lookback_rolling_vehicle_kernel<false, false> : lookback_rolling_vehicle_kernel<false, true>);
        return issue(kern, dim3(n_vehicles), dim3(RV_THREADS), 0, st, reinterpret_cast<const float4*>(bank), N, Npad, W,
                     make_step(Ts), slot, hist, err_ring, avg_err, idx_offset, emit, K, out);
    }
    const int n_lists = (N + LB_THREADS - 1) / LB_THREADS;
    const bool in_kernel = emit && K > 0 && ticket && out && cta_lists && best_key && n_lists <= LB_THREADS * MERGE_LPT;
    FusedMerge fm = {in_kernel ? ticket : nullptr, in_kernel ? out : nullptr, in_kernel ? K : 0};
    int rc = lookback_rolling_impl(bank, N, Npad, nullptr, hist, n_vehicles, slot, W, Ts, err_ring, avg_err, best_key,
                                   cta_lists, idx_offset, geom_shared, emit, fm, stream);
    if (rc || !emit || K == 0 || in_kernel) return rc;
    if (!cta_lists || !out) return LLAMPC_E_ARG;
    return llampc_topk_merge_lists(cta_lists, n_lists, n_vehicles, K, best_key, out, stream);
}

extern "C" int llampc_lookback_num_lists(int N, int W, int split) {
    if (N <= 0 || W <= 0) return LLAMPC_E_ARG;
    if (split >= 32) split -= 32;
    if (split == 0) split = choose_split(N, W);
    if (split > W) split = 1;
    if (split != 1 && split != 2 && split != 4 && split != 8 && split != 16) return LLAMPC_E_ARG;
    const int cpb = k1_cands_per_cta(N) / split;
    return (N + cpb - 1) / cpb;
}

static int merge_lists_impl(const llampc_key_t* cta_lists, int n_lists, int n_vehicles, int K, llampc_key_t* best_key,
                            llampc_key_t* out, const PeerXchg& px, llampc_stream_t stream) {
    if (!cta_lists || !out || n_lists <= 0 || n_vehicles <= 0) return LLAMPC_E_ARG;
    if (K < 0 || K > LLAMPC_LIST_LEN || n_lists > 1024 * MERGE_LPT) return LLAMPC_E_RANGE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_lists <= 128 * MERGE_LPT)
        topk_merge_lists_kernel<128><<<n_vehicles, 128, 0, st>>>(cta_lists, n_lists, K, best_key, out, px);
    else if (n_lists <= 256 * MERGE_LPT)
        topk_merge_lists_kernel<256><<<n_vehicles, 256, 0, st>>>(cta_lists, n_lists, K, best_key, out, px);
    else
        topk_merge_lists_kernel<1024><<<n_vehicles, 1024, 0, st>>>(cta_lists, n_lists, K, best_key, out, px);
    return (int)cudaGetLastError();
}

extern "C" int llampc_topk_merge_lists(const llampc_key_t* cta_lists, int n_lists, int n_vehicles, int K,
                                       llampc_key_t* best_key, llampc_key_t* out, llampc_stream_t stream) {
    return merge_lists_impl(cta_lists, n_lists, n_vehicles, K, best_key, out, PeerXchg{nullptr, 0, 0, 0}, stream);
}

extern "C" int llampc_fill_keys(llampc_key_t* keys, int n, llampc_stream_t stream) {
    if (!keys || n <= 0) return LLAMPC_E_ARG;
    fill_keys_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, n);
    return (int)cudaGetLastError();
}

static int topk_per_cta(int N) {
    int per = 4096;
    while ((long)((N + per - 1) / per) > 256) per *= 2;      // at most 256 CTAs -> the merge sees <= 256*K keys
    return per;
}

extern "C" int llampc_topk_scratch_ctas(int N) {
    if (N <= 0) return LLAMPC_E_ARG;
    int per = topk_per_cta(N);
    return (N + per - 1) / per;
}

extern "C" int llampc_topk_f32(const float* err, int N, int idx_offset, int K, llampc_key_t* scratch,
                               unsigned* counter, llampc_key_t* out_keys, llampc_stream_t stream) {
    if (!err || !scratch || !counter || !out_keys || N <= 0) return LLAMPC_E_ARG;
    if (K <= 0 || K > LLAMPC_MAX_K) return LLAMPC_E_RANGE;
    const int per = topk_per_cta(N);
    const int ctas = (N + per - 1) / per;
    topk_kernel<<<ctas, TK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(err, N, idx_offset, K, per, scratch,
                                                                            counter, out_keys);
    return (int)cudaGetLastError();
}

extern "C" int llampc_refine_f64(const double* bank64, int N, const double* hist64, int W, double Ts,
                                 const llampc_key_t* keys, int n_fin, int idx_offset, double* out_err64,
                                 llampc_stream_t stream) {
    if (!bank64 || !hist64 || !keys || !out_err64 || N <= 0 || W <= 0 || n_fin <= 0) return LLAMPC_E_ARG;
    NewRow64 nr;
    nr.slot = -1;
    FinalCopy fc = {nullptr, nullptr, nullptr, 0, 0};
    PeerGather pg = {nullptr, 0, 0, 0, 0};
    refine_f64_kernel<<<n_fin, RF_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(bank64, N, const_cast<double*>(hist64), W, Ts,
                                                                           keys, idx_offset, out_err64, nr, fc, pg);
    return (int)cudaGetLastError();
}

extern "C" int llampc_rk4_batch_f32(const float* bank, int N, int Npad, const double* x64, int x_shared,
                                    const double* u64v, int u_shared, double Ts, double* out64, int out_cols,
                                    llampc_stream_t stream) {
    if (!bank || !x64 || !u64v || !out64 || N <= 0 || Npad < N) return LLAMPC_E_ARG;
    if (out_cols < 1 || out_cols > 6) return LLAMPC_E_RANGE;
    if (!aligned16(bank)) return LLAMPC_E_ALIGN;
    onestep_kernel<0><<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(bank), N, Npad, x64, x_shared, u64v, u_shared, (float)Ts, out64, out_cols);
    return (int)cudaGetLastError();
}

extern "C" int llampc_rhs_batch_f32(const float* bank, int N, int Npad, const double* x64, int x_shared,
                                    const double* u64v, int u_shared, double* out64, llampc_stream_t stream) {
    if (!bank || !x64 || !u64v || !out64 || N <= 0 || Npad < N) return LLAMPC_E_ARG;
    if (!aligned16(bank)) return LLAMPC_E_ALIGN;
    onestep_kernel<1><<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(bank), N, Npad, x64, x_shared, u64v, u_shared, 0.0f, out64, 6);
    return (int)cudaGetLastError();
}

extern "C" int llampc_forces_batch_f32(const float* bank, int N, int Npad, const double* x64, int x_shared,
                                       const double* u64v, int u_shared, double* out64, llampc_stream_t stream) {
    if (!bank || !x64 || !u64v || !out64 || N <= 0 || Npad < N) return LLAMPC_E_ARG;
    if (!aligned16(bank)) return LLAMPC_E_ALIGN;
    onestep_kernel<2><<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(bank), N, Npad, x64, x_shared, u64v, u_shared, 0.0f, out64, 5);
    return (int)cudaGetLastError();
}

// One-launch tick with the tree finish: K1b (persistent warp tasks) when K1's tiling cannot fill the SMs (fewer CTAs
// than SMs: small banks, or few candidates with a long window), else K1 with the tree merge.
static int lookback_tree_dispatch(const float* bank, int N, int Npad, const float* hist, int W, double Ts, float* avg_err,
                                  int idx_offset, int geom_shared, int split, int K, void* workspace,
                                  unsigned long long workspace_bytes, llampc_key_t* out, const NewRow& nr,
                                  const PeerXchg& px, llampc_stream_t stream) {
    if (!workspace || !out) return LLAMPC_E_ARG;
    if (K <= 0 || K > LLAMPC_LIST_LEN) return LLAMPC_E_RANGE;
    const int n_lists = llampc_lookback_num_lists(N, W, split);
    if (n_lists <= 0) return n_lists ? n_lists : LLAMPC_E_ARG;
    // K1b pays off only where K1 leaves SMs idle AND every K1 thread has a long serial walk (measured on B200: 4,096
    // candidates x 1,024 rows 115 us against 133 us; everywhere else K1 + tree is equal or faster)
    int sy = split & 31;
    if (sy == 0) sy = choose_split(N, W);
    if (sy > W) sy = 1;
    const char* force = getenv("LLAMPC_TREE_KERNEL");            // experiments: "k1" / "k1b"
    const bool k1b = force ? (force[0] == 'k' && force[1] == '1' && force[2] == 'b')
                           : (n_lists < device_sms() && (W + sy - 1) / sy >= 64);
    if (k1b)
        return lookback_balanced_launch(bank, N, Npad, hist, W, Ts, avg_err, idx_offset, geom_shared, (split & 32) != 0, K,
                                        workspace, workspace_bytes, out, nr, px, static_cast<cudaStream_t>(stream));
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return LLAMPC_E_ALIGN;
    const TreeLayout lay = tree_layout(N, 0);
    if (workspace_bytes < lay.bytes) return LLAMPC_E_ARG;
    TreeMerge tm = {tree_workspace(static_cast<unsigned char*>(workspace), lay), out, K};
    FusedMerge none = {nullptr, nullptr, 0};
    return lookback_window_impl(bank, N, Npad, hist, W, 1, W, Ts, avg_err, nullptr, nullptr, idx_offset, geom_shared, split,
                                nr, none, stream, px, tm);
}

extern "C" long long llampc_lookback_balanced_workspace_bytes(int N, int W) {
    if (N <= 0 || W <= 0) return LLAMPC_E_ARG;
    if (W > LLAMPC_MAX_W) return LLAMPC_E_RANGE;
    return lookback_balanced_workspace_bytes(N, W);
}

extern "C" int llampc_lookback_window_balanced_f32(const float* bank, int N, int Npad, const float* hist, int W, double Ts,
                                                   float* avg_err, int idx_offset, int geom_shared, int fast_sin, int K,
                                                   void* workspace, unsigned long long workspace_bytes, llampc_key_t* out,
                                                   llampc_key_t* const* peer_bufs, int world, int rank, unsigned seq,
                                                   llampc_stream_t stream) {
    NewRow nr;
    nr.slot = -1;
    PeerXchg px = {nullptr, 0, 0, 0};
    if (peer_bufs) {
        if (world < 2 || world > 32 || rank < 0 || rank >= world) return LLAMPC_E_RANGE;
        px.peers = peer_bufs; px.world = world; px.rank = rank; px.seq = seq;
    }
    return lookback_tree_dispatch(bank, N, Npad, hist, W, Ts, avg_err, idx_offset, geom_shared, fast_sin ? 32 : 0, K,
                                  workspace, workspace_bytes, out, nr, px, stream);
}

extern "C" int llampc_lookback_finish(llampc_tick_t* t, llampc_stream_t stream);

extern "C" int llampc_lookback_tick(llampc_tick_t* t, llampc_stream_t stream) {
    if (!t || !t->bank || !t->hist || !t->best_key || !t->result || !t->result_h) return LLAMPC_E_ARG;
    if (t->slot < 0 || t->slot >= t->W || t->K < 0 || t->n_refine < 0) return LLAMPC_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kt = t->K > t->n_refine ? t->K : t->n_refine;
    if (Kt > LLAMPC_MAX_K) return LLAMPC_E_RANGE;
    const bool fused = t->cta_lists != nullptr && Kt <= LLAMPC_LIST_LEN;
    llampc_key_t* keys = t->result;                                  // [0] best, [1..Kt] finalists
    double* errs = reinterpret_cast<double*>(t->result + 1 + Kt);    // [Kt] fp64 scores
    NewRow nr;
    nr.slot = -1;
    if (t->row32_h) {                                                // the row rides in the kernel parameters
        for (int i = 0; i < LLAMPC_HIST_ROW; ++i) nr.v[i] = t->row32_h[i];
        nr.slot = t->slot;
    }
    // ---- the two kernels of a tick (scoring + fp64 re-score) are replayed as one re-parameterised CUDA graph when the
    // tick is a one-launch scoring kernel followed by the re-score (LLAMPC_TICK_GRAPH=0: plain stream launches)
    static int graph_env = -1;
    if (graph_env < 0) { const char* e = getenv("LLAMPC_TICK_GRAPH"); graph_env = (e && e[0] == '0') ? 0 : 1; }
    const int n_lists_roll = (t->N + LB_THREADS - 1) / LB_THREADS;
    const bool roll_one_launch = t->rolling == 1 && t->ticket != nullptr && Kt > 0 && n_lists_roll <= LB_THREADS * MERGE_LPT &&
                                 t->err_ring && t->row32_h && fused;
    const bool tree_one_launch = !t->rolling && fused && t->workspace && Kt > 0;
    PendingLaunch pending[TICK_GRAPH_MAX_NODES];
    struct CollectGuard {                                            // never leave the collector armed on an error return
        bool on;
        ~CollectGuard() { if (on) { g_collect = nullptr; g_collect_n = 0; } }
    } guard = {false};
    if (graph_env && Kt > 0 && t->n_refine > 0 && (roll_one_launch || tree_one_launch)) {
        if (!t->graph_state) t->graph_state = calloc(1, sizeof(TickGraph));
        if (t->graph_state) {
            g_collect = pending;
            g_collect_n = 0;
            guard.on = true;
        }
    }
    int rc;
    if (t->rolling) {
        // the reference's rolling bookkeeping: one new error column + ring re-sum, then the list merge
        if (!t->err_ring || !t->row32_h || !fused) return LLAMPC_E_ARG;
        const int n_lists_r = (t->N + LB_THREADS - 1) / LB_THREADS;
        const bool in_kernel_r = t->rolling == 1 && t->ticket != nullptr && Kt > 0 && n_lists_r <= LB_THREADS * MERGE_LPT;
        FusedMerge fmr = {in_kernel_r ? t->ticket : nullptr, in_kernel_r ? keys : nullptr, in_kernel_r ? Kt : 0};
        rc = lookback_rolling_impl(t->bank, t->N, t->Npad, t->row32_h, nullptr, 1, t->slot, t->W, t->Ts, t->err_ring,
                                   t->avg_err, t->best_key, t->cta_lists, t->idx_offset, t->geom_shared,
                                   t->rolling > 1 ? 0 : 1, fmr, stream);
        if (rc) return rc;
        if (t->rolling > 1) {                                        // window still filling: column stored, no decision
            if (t->n_refine > 0 && t->row64_h && t->hist64)
                LLAMPC_CUDA_TRY(cudaMemcpyAsync(t->hist64 + (size_t)t->slot * LLAMPC_HIST64_ROW, t->row64_h,
                                                LLAMPC_HIST64_ROW * sizeof(double), cudaMemcpyHostToDevice, st));
            return 0;
        }
        if (!in_kernel_r) {
            rc = llampc_topk_merge_lists(t->cta_lists, n_lists_r, 1, Kt, t->best_key, keys, stream);
            if (rc) return rc;
        }
    } else if (fused && t->workspace && Kt > 0) {
        // K1 (or K1b for small grids) with the tree merge, one launch; writes keys[0..LIST_LEN] itself (best_key stays armed)
        rc = lookback_tree_dispatch(t->bank, t->N, t->Npad, t->hist, t->W, t->Ts, t->avg_err, t->idx_offset,
                                    t->geom_shared, t->split, Kt, t->workspace, t->workspace_bytes, keys, nr,
                                    PeerXchg{nullptr, 0, 0, 0}, stream);
        if (rc) return rc;
    } else if (fused) {
        // K1 (block arg-min + per-CTA sorted lists) -> list merge (also moves best_key to keys[0] and re-arms it)
        const int n_lists = llampc_lookback_num_lists(t->N, t->W, t->split);
        const bool in_kernel = t->ticket != nullptr && Kt > 0 && n_lists <= LB_THREADS * MERGE_LPT;
        FusedMerge fm = {in_kernel ? t->ticket : nullptr, in_kernel ? keys : nullptr, in_kernel ? Kt : 0};
        rc = lookback_window_impl(t->bank, t->N, t->Npad, t->hist, t->W, 1, t->W, t->Ts, t->avg_err, t->best_key,
                                  t->cta_lists, t->idx_offset, t->geom_shared, t->split, nr, fm, stream);
        if (rc) return rc;
        if (!in_kernel) {
            rc = llampc_topk_merge_lists(t->cta_lists, n_lists, 1, Kt, t->best_key, keys, stream);
            if (rc) return rc;
        }
    } else {
        if (!t->avg_err) return LLAMPC_E_ARG;
        rc = llampc_fill_keys(t->best_key, 1, stream);
        if (rc) return rc;
        FusedMerge none = {nullptr, nullptr, 0};
        rc = lookback_window_impl(t->bank, t->N, t->Npad, t->hist, t->W, 1, t->W, t->Ts, t->avg_err, t->best_key, nullptr,
                                  t->idx_offset, t->geom_shared, t->split, nr, none, stream);
        if (rc) return rc;
        LLAMPC_CUDA_TRY(cudaMemcpyAsync(keys, t->best_key, sizeof(llampc_key_t), cudaMemcpyDeviceToDevice, st));
        if (Kt > 0) {
            if (!t->topk_scratch || !t->topk_counter) return LLAMPC_E_ARG;
            rc = llampc_topk_f32(t->avg_err, t->N, t->idx_offset, Kt, t->topk_scratch, t->topk_counter, keys + 1, stream);
            if (rc) return rc;
        }
    }
    const bool refine = Kt > 0 && t->n_refine > 0;
    t->pending_seq = 0;
    t->pending_words = 0;
    if (refine) {
        if (!t->bank64 || !t->hist64) return LLAMPC_E_ARG;
        NewRow64 nr64;
        nr64.slot = -1;
        if (t->row64_h) {
            for (int i = 0; i < LLAMPC_HIST64_ROW; ++i) nr64.v[i] = t->row64_h[i];
            nr64.slot = t->slot;
        }
        // zero-copy hand-off: needs the ticket word after the K1 ticket and a host buffer with one spare word
        const bool gather = t->peer_world > 1 && t->peer_bufs != nullptr;
        const int words = gather ? 1 + 2 * Kt * t->peer_world : 1 + 2 * Kt;
        FinalCopy fc = {nullptr, nullptr, nullptr, 0, 0};
        PeerGather pg = {nullptr, 0, 0, 0, 0};
        if (gather) {
            if (!(t->zero_copy && t->ticket)) return LLAMPC_E_ARG;     // the gather rides on the zero-copy hand-off
            pg.peers = t->peer_bufs; pg.world = t->peer_world; pg.rank = t->peer_rank; pg.seq = t->peer_seq; pg.kt = Kt;
        }
        if (t->zero_copy && t->ticket) {
            void* dptr = t->mapped_for == t->result_h ? t->mapped_dev : nullptr;
            if (!dptr && cudaHostGetDevicePointer(&dptr, t->result_h, 0) == cudaSuccess && dptr) {
                t->mapped_dev = dptr;
                t->mapped_for = t->result_h;
            }
            if (dptr) {
                static unsigned long long seq_counter = 1;
                fc.src = t->result;
                fc.dst_host = static_cast<volatile u64*>(dptr);
                fc.ticket = t->ticket + 1;
                fc.words = 1 + 2 * Kt;
                fc.seq = ++seq_counter;
                reinterpret_cast<volatile llampc_key_t*>(t->result_h)[words] = 0;
                t->pending_seq = fc.seq;
                t->pending_words = words;
            } else {
                (void)cudaGetLastError();                            // result_h is not mapped: use the copy path
                if (gather) return LLAMPC_E_ARG;
            }
        }
        rc = issue(refine_f64_kernel, dim3(Kt), dim3(RF_THREADS), 0, st, t->bank64, t->N, t->hist64, t->W, t->Ts,
                   static_cast<const u64*>(keys + 1), t->idx_offset, errs, nr64, fc, pg);
        if (rc) return rc;
    }
    if (guard.on) {                                                  // replay what was collected as one graph
        const int n = g_collect_n;
        g_collect = nullptr;
        g_collect_n = 0;
        guard.on = false;
        if (n > 0) {
            rc = tick_graph_launch(static_cast<TickGraph*>(t->graph_state), pending, n, st);
            if (rc) return rc;
        }
    }
    if (t->pending_seq == 0)
        LLAMPC_CUDA_TRY(cudaMemcpyAsync(t->result_h, t->result, (size_t)(1 + Kt + (refine ? Kt : 0)) * 8,
                                        cudaMemcpyDeviceToHost, st));
    if (!t->sync) return 0;                                          // the caller finishes with llampc_lookback_finish
    return llampc_lookback_finish(t, stream);
}

// Second half of a tick: waits for the result (polls the zero-copy sequence word, or synchronises the stream) and
// orders the finalists on the host.  llampc_lookback_tick calls it itself when t->sync != 0; with sync = 0 the caller
// may do unrelated host work (the NMPC solve of the next tick) between the two calls.
extern "C" int llampc_lookback_finish(llampc_tick_t* t, llampc_stream_t stream) {
    if (!t || !t->result_h) return LLAMPC_E_ARG;
    if (t->rolling > 1) return 0;                                    // a filling tick produced no result
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kt = t->K > t->n_refine ? t->K : t->n_refine;
    const bool refine = Kt > 0 && t->n_refine > 0;
    const bool gather = refine && t->peer_world > 1 && t->peer_bufs != nullptr;
    if (t->pending_seq != 0) {
        volatile llampc_key_t* flag = reinterpret_cast<volatile llampc_key_t*>(t->result_h) + t->pending_words;
        long spins = 0;
        while ((*flag & ~PEER_FAIL_BIT) != t->pending_seq) {
            if (++spins > 20000000L) {                               // tens of ms: something is wrong, stop spinning
                LLAMPC_CUDA_TRY(cudaStreamSynchronize(st));
                if ((*flag & ~PEER_FAIL_BIT) != t->pending_seq) return (int)cudaErrorUnknown;
                break;
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        __asm__ __volatile__("" ::: "memory");                         // the result words are read after the flag
        t->pending_seq = 0;
        if (*flag & PEER_FAIL_BIT) return LLAMPC_E_PEER;             // a peer never delivered its finalists: no decision
        if (gather) {
            // host: pick the Kt best of the world * Kt finalists (fp64 score, ties by index; NaN / padding last)
            // and compact them into the single-GPU layout [arg-min | Kt keys | Kt scores]
            const int total = Kt * t->peer_world;
            llampc_key_t* raw = t->result_h + 1;
            llampc_key_t ak[LLAMPC_MAX_K * 32];
            double ae[LLAMPC_MAX_K * 32];
            if (total > LLAMPC_MAX_K * 32) return LLAMPC_E_RANGE;
            for (int q = 0; q < t->peer_world; ++q)
                for (int j = 0; j < Kt; ++j) {
                    ak[q * Kt + j] = raw[(size_t)q * 2 * Kt + j];
                    memcpy(&ae[q * Kt + j], &raw[(size_t)q * 2 * Kt + Kt + j], 8);
                }
            llampc_key_t* hk2 = t->result_h + 1;
            double* he2 = reinterpret_cast<double*>(t->result_h + 1 + Kt);
            for (int i = 0; i < Kt; ++i) {                           // partial selection sort
                int best = -1;
                for (int j = i; j < total; ++j) {
                    if (ak[j] == ~0ull || ae[j] != ae[j]) continue;
                    if (best < 0 || ae[j] < ae[best] ||
                        (ae[j] == ae[best] && (unsigned)(ak[j] & 0xffffffffull) < (unsigned)(ak[best] & 0xffffffffull)))
                        best = j;
                }
                if (best < 0) { for (int r = i; r < Kt; ++r) { hk2[r] = ~0ull; he2[r] = NAN; } break; }
                const llampc_key_t tk = ak[best]; const double te = ae[best];
                ak[best] = ak[i]; ae[best] = ae[i];
                ak[i] = tk; ae[i] = te;
                hk2[i] = tk; he2[i] = te;
            }
        }
    } else {
        LLAMPC_CUDA_TRY(cudaStreamSynchronize(st));
    }
    // order the finalists on the host: by fp64 score (ties: lower index), NaN / padded entries last
    llampc_key_t* hk = t->result_h + 1;
    double* he = reinterpret_cast<double*>(t->result_h + 1 + Kt);
    if (!refine) {
        for (int i = 0; i < Kt; ++i) {
            const unsigned bits = (unsigned)(hk[i] >> 32);
            float f;
            memcpy(&f, &bits, 4);
            he[i] = hk[i] == ~0ull ? NAN : (double)f;
        }
    }
    for (int i = 1; i < Kt; ++i) {                                   // insertion sort, Kt <= 64
        const llampc_key_t k = hk[i];
        const double e = he[i];
        const unsigned idx = (unsigned)(k & 0xffffffffull);
        int j = i - 1;
        while (j >= 0) {
            const bool after = (he[j] != he[j]) ? (e == e) : (e == e && (e < he[j] || (e == he[j] && idx < (unsigned)(hk[j] & 0xffffffffull))));
            if (!after) break;
            hk[j + 1] = hk[j];
            he[j + 1] = he[j];
            --j;
        }
        hk[j + 1] = k;
        he[j + 1] = e;
    }
    return 0;
}

// Decodes the ordered finalists left in t->result_h by llampc_lookback_finish into plain index / score arrays.
extern "C" int llampc_lookback_decode(const llampc_tick_t* t, long long* idx_out, double* score_out, int* n_valid) {
    if (!t || !idx_out || !score_out || !n_valid || !t->result_h) return LLAMPC_E_ARG;
    *n_valid = 0;
    if (t->rolling > 1) return 0;                                    // window still filling
    const int Kt = t->K > t->n_refine ? t->K : t->n_refine;
    const llampc_key_t* hk = t->result_h + 1;
    const double* he = reinterpret_cast<const double*>(t->result_h + 1 + Kt);
    if (Kt == 0) {                                                   // arg-min only
        const unsigned bits = (unsigned)(t->result_h[0] >> 32);
        float f;
        memcpy(&f, &bits, 4);
        idx_out[0] = (long long)(t->result_h[0] & 0xffffffffull);
        score_out[0] = (double)f;
        *n_valid = 1;
        return 0;
    }
    int n = 0;
    for (int i = 0; i < Kt; ++i) {
        if (hk[i] == ~0ull || he[i] != he[i]) break;                 // ordered: NaN / padding last
        idx_out[n] = (long long)(hk[i] & 0xffffffffull);
        score_out[n] = he[i];
        ++n;
    }
    *n_valid = n;
    return 0;
}

// One FFI crossing per MPC tick for a host that holds the transition as three fp64 vectors: packs the history
// row(s) into the caller's scratch (t->row32_h / t->row64_h must point to writable host buffers), runs the tick
// (sync forced) and decodes the ordered finalists into plain index / score arrays.
extern "C" int llampc_lookback_push(llampc_tick_t* t, const double* x_k, const double* u_k, const double* x_k1,
                                    double lf_shared, double lr_shared, long long* idx_out, double* score_out,
                                    int* n_valid, llampc_stream_t stream) {
    if (!t || !x_k || !u_k || !x_k1 || !idx_out || !score_out || !n_valid || !t->row32_h) return LLAMPC_E_ARG;
    int rc = llampc_hist_row_pack_h(x_k, u_k, x_k1, t->Ts, lf_shared, lr_shared, const_cast<float*>(t->row32_h),
                                    const_cast<double*>(t->row64_h));
    if (rc) return rc;
    const int sync_was = t->sync;
    t->sync = 1;
    rc = llampc_lookback_tick(t, stream);
    t->sync = sync_was;
    if (rc) return rc;
    return llampc_lookback_decode(t, idx_out, score_out, n_valid);
}

// Frees what llampc_lookback_tick attached to the struct (the CUDA graph of the tick).  Safe to call more than once.
extern "C" int llampc_lookback_tick_release(llampc_tick_t* t) {
    if (!t) return LLAMPC_E_ARG;
    if (t->graph_state) {
        tick_graph_destroy(static_cast<TickGraph*>(t->graph_state));
        free(t->graph_state);
        t->graph_state = nullptr;
    }
    return 0;
}

// layout probes for FFI bindings that mirror llampc_tick_t by hand
extern "C" int llampc_tick_sizeof(void) { return (int)sizeof(llampc_tick_t); }
extern "C" int llampc_tick_offsetof(int which) {
    switch (which) {
        case 0: return (int)offsetof(llampc_tick_t, Ts);
        case 1: return (int)offsetof(llampc_tick_t, cta_lists);
        case 2: return (int)offsetof(llampc_tick_t, result_h);
        case 3: return (int)offsetof(llampc_tick_t, peer_seq);
        case 4: return (int)offsetof(llampc_tick_t, rolling);
        default: return -1;
    }
}
