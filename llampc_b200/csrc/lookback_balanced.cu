// K1b: work-balanced look-back window kernel with an in-kernel tree merge of the top-K.  sm_100a.
//
// Same arithmetic and same reference behaviour as K1 (lookback.cu): evaluate_models_vectorized
// (llampc/mpc/evaluate_models_vectorized.py:4-23) + errors / window mean / argmin / argsort[:K] of
// run_nmpc_orca_llampc_rt.py:349-360.  What changes is how the N x W candidate-steps are laid on the machine:
//
//   * K1 gives every CTA a fixed tile (128/SY candidates x W/SY rows), so the grid is a property of the problem and
//     its last wave is partially filled (C2: 1,024 CTAs on 888 resident slots = 1.15 waves, the tail runs one CTA per
//     SM).  K1b launches (SMs x resident CTAs per SM) persistent CTAs whose warps pull warp-tasks (32 candidates x R
//     window rows) from an atomic counter; see BalSched below.
//   * A warp-group of 32 candidates whose W rows are spread over several tasks is combined through a small L2-resident
//     scratch: every task stores its 32 partial sums, the last one to arrive (atomic counter per warp-group) adds the
//     partials in row order (fixed order => run-to-run deterministic) and finalises the warp-group.
//   * The top-K is finished by a tree of warp-level merges that overlaps the integration: the finaliser of a warp-group
//     writes its 16 smallest keys; the last of every 32 lists merges them into one (one warp, heads in registers,
//     REDUX minima), and so on until one list is left.  The serial tail after the last RK4 step is the two or three
//     32-way merges on the path to the root instead of one CTA walking every list of the launch.
#include "llampc_common.cuh"
#include "llampc_model.cuh"
#include "lookback_select.cuh"
#include "llampc_launch.cuh"
#include <stdio.h>

namespace llampc {

// Optional timeline (LLAMPC_BAL_TRACE=1, experiments only): per CTA the global timer at kernel entry, after the
// history staging, after its last RK4 row and at exit.
constexpr int BAL_TRACE_CTAS = 8192;
__device__ unsigned long long g_bal_trace[BAL_TRACE_CTAS * 4];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Warp tasks.  The unit of scheduling is one warp-task = 32 candidates x R consecutive window rows (task t ->
// candidate warp-group t / SY, row chunk t % SY, SY = ceil(W / R)); R is chosen on the host so that every resident
// warp sees several tasks.  Warps pull tasks from one atomic counter (the next index is prefetched while the current
// task runs) and never meet at a CTA barrier after the prologue: the warp schedulers of an SM favour its oldest warps,
// so equal STATIC shares finish anywhere between 45 % and 100 % of the kernel time and leave the SM half empty
// while the youngest warps drain; with short tasks handed out dynamically every scheduler keeps its warps busy until
// the pool is empty and the drain is at most one task long.  The schedule is a pure function of (N, W, grid), and the
// row chunks of a warp-group are always added in row order.
struct BalSched {
    int R;              // window rows per task
    int SY;             // row chunks per candidate warp-group
    int n_wg;           // candidate warp-groups = ceil(N / 32)
    int n_tasks;        // n_wg * SY
};

struct BalPlan {
    int grid;
    BalSched sch;
    TreeLayout lay;
};

template <bool GEOM_SHARED, bool MUFU_SIN>
__global__ void __launch_bounds__(LB_THREADS, LLAMPC_LB_MIN_BLOCKS)
lookback_balanced_kernel(const float4* __restrict__ bank, int N, int Npad, const float* __restrict__ hist, int W,
                         StepSize z, float* __restrict__ avg_err, int idx_offset, NewRow nr, int K, BalWs ws,
                         BalSched sch, u64* __restrict__ out, PeerXchg px, int trace_arg) {
#ifdef LLAMPC_BAL_TRACE
    const bool trace = trace_arg != 0;             // experiments build only (make trace): costs registers
#else
    constexpr bool trace = false;
#endif
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ u64 mrows[BAL_FAN][BAL_ROW_PAD];
    __shared__ u64 sfin[LB_THREADS / 32][LLAMPC_LIST_LEN];
    float4* srow = reinterpret_cast<float4*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bytes = (unsigned)W * (LLAMPC_HIST_ROW * 4);
    const int gw = (int)blockIdx.x * (LB_THREADS / 32) + (threadIdx.x >> 5);      // global warp (trace slot)
    unsigned long long tr_max = 0, tr_trips = 0, tr_t0 = 0, tr_rows = 0;
    if (tid == 0) {                                // one thread arms the barrier and starts the bulk copy right away
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, bytes);
        tma_bulk_g2s(srow, hist, bytes, &mbar);
    }
    __syncthreads();                               // the initialised barrier is visible to the waiting threads

    const int n_warps = (int)gridDim.x * (LB_THREADS / 32);
    int t = (int)blockIdx.x * (LB_THREADS / 32) + warp;        // first task: no atomic
    Cand p = load_cand(bank, Npad, min((t < sch.n_tasks ? t / sch.SY : 0) * 32 + lane, N - 1));   // overlaps the bulk copy
    mbar_wait(&mbar, 0);
    if (nr.slot >= 0) {                            // uniform over the grid
        if (tid < LLAMPC_HIST_ROW / 4) {
            const float4 q = make_float4(nr.v[4 * tid], nr.v[4 * tid + 1], nr.v[4 * tid + 2], nr.v[4 * tid + 3]);
            srow[nr.slot * 5 + tid] = q;
            if (blockIdx.x == 0) reinterpret_cast<float4*>(const_cast<float*>(hist))[nr.slot * 5 + tid] = q;
        }
        __syncthreads();
    }

    bool have_p = true;
    u64 run = ~0ull;                               // this warp's running top-16 (lanes 0..15 ascending, ~0 above)
    while (t < sch.n_tasks) {                      // uniform over the warp; no CTA barrier from here on
        unsigned nxt = 0;
        if (trace) tr_t0 = gtimer();
        const int wg = t / sch.SY, sy = t - wg * sch.SY;
        const int r0 = sy * sch.R, r1 = min(W, r0 + sch.R);
        const int cand = wg * 32 + lane;
        const bool valid = cand < N;
        const int ci = valid ? cand : N - 1;
        if (!have_p) p = load_cand(bank, Npad, ci);
        have_p = false;

        float acc = 0.0f;
#pragma unroll 1
        for (int w = r0; w < r1; ++w) {
            // the next task is claimed during the LAST row (latency hidden by that row): a warp starved by the
            // scheduler must not sit on a second task while it crawls through the current one
            if (w == r1 - 1 && lane == 0) nxt = atomicAdd(ws.next, 1u);
            HistRow r;
            r.q0 = srow[w * 5 + 0];
            r.q1 = srow[w * 5 + 1];
            r.q2 = srow[w * 5 + 2];
            r.q3 = srow[w * 5 + 3];
            r.q4 = srow[w * 5 + 4];
            bool ok;
            float e = lookback_step_fast<GEOM_SHARED, MUFU_SIN>(p, r, z, ok);
            if (!ok) e = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, ci, srow + w * 5, z);
            acc += e;
        }

        if (trace) { tr_rows = gtimer() - tr_t0; tr_trips += 1; }
        bool fin = true;
        if (sch.SY > 1) {                          // the warp-group's window is spread over SY tasks
            __stcg(ws.part + (size_t)t * 32 + lane, acc);
            __syncwarp();
            unsigned old = 0;
            if (lane == 0) {
                if (!(trace_arg & 2)) __threadfence();         // cumulative: the warp's partial sums are ordered before the count
                old = atomicAdd(ws.gcount + wg, 1u);
            }
            old = __shfl_sync(0xffffffffu, old, 0);
            fin = old == (unsigned)(sch.SY - 1);   // the last chunk to arrive finalises the warp-group
            if (fin) {
                if (lane == 0) ws.gcount[wg] = 0;  // ready for the next launch on the same stream
                if (!(trace_arg & 2)) __threadfence();
                const float* pp = ws.part + (size_t)wg * sch.SY * 32 + lane;
                acc = __ldcg(pp);
                for (int j = 1; j < sch.SY; ++j) acc += __ldcg(pp + j * 32);      // row order
            }
        }
        if (fin) {
            // errors = mean over the 4 scored states (rt.py:349); avg = mean over the window (rt.py:357)
            const float err = acc * (0.25f / (float)W);
            u64 key = ~0ull;
            if (valid) {
                if (avg_err) avg_err[cand] = err;
                key = pack_key(err, (unsigned)(idx_offset + cand));
            }
            key = warp_sort_u64(key, lane);
            run = warp_merge_low32(run, __shfl_sync(0xffffffffu, key, 31 - lane), lane);
            if (lane >= LLAMPC_LIST_LEN) run = ~0ull;
        }
        t = n_warps + (int)__shfl_sync(0xffffffffu, nxt, 0);
        if (trace) { const unsigned long long d = gtimer() - tr_t0; tr_max = d > tr_max ? d : tr_max; }
    }
    if (trace && lane == 0 && gw < BAL_TRACE_CTAS) {
        g_bal_trace[gw * 4 + 0] = tr_t0;           // start of the last task
        g_bal_trace[gw * 4 + 1] = tr_rows;         // its rows-only duration
        g_bal_trace[gw * 4 + 2] = gtimer();        // loop exit
        g_bal_trace[gw * 4 + 3] = tr_trips;        // tasks done
    }
    // CTA list = merge of its four warps' running lists; then the tree over the CTA lists
    if (lane < LLAMPC_LIST_LEN) sfin[warp][lane] = run;
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int w = 1; w < LB_THREADS / 32; ++w) {
        const u64 other_rev = lane >= LLAMPC_LIST_LEN ? sfin[w][31 - lane] : ~0ull;
        run = warp_merge_low32(run, other_rev, lane);
        if (lane >= LLAMPC_LIST_LEN) run = ~0ull;
    }
    tree_merge(run, lane, (int)blockIdx.x, (int)gridDim.x, K, ws, mrows, out, px);
}

// resident CTAs per SM for a window of W rows (all four kernel variants have the same footprint); cached per device
static int bal_occupancy(int W, int* occ_out) {
    struct Entry { int w, occ; };
    static Entry cache[64] = {};                   // one entry per device, keyed by the last W seen there (w = 0: empty)
    int dev = 0;
    LLAMPC_CUDA_TRY(cudaGetDevice(&dev));
    Entry local = {0, 0};
    Entry& e = (dev >= 0 && dev < 64) ? cache[dev] : local;
    if (W != e.w) {
        const size_t smem = (size_t)W * (LLAMPC_HIST_ROW * 4);
        if (smem > 32 * 1024) {
            LLAMPC_CUDA_TRY(cudaFuncSetAttribute(lookback_balanced_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            LLAMPC_CUDA_TRY(cudaFuncSetAttribute(lookback_balanced_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            LLAMPC_CUDA_TRY(cudaFuncSetAttribute(lookback_balanced_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            LLAMPC_CUDA_TRY(cudaFuncSetAttribute(lookback_balanced_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        }
        int o = 0;
        LLAMPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, lookback_balanced_kernel<true, true>, LB_THREADS, smem));
        if (o <= 0) return LLAMPC_E_RANGE;
        e.occ = o < TREE_MAX_CTAS_PER_SM ? o : TREE_MAX_CTAS_PER_SM;
        e.w = W;
    }
    *occ_out = e.occ;
    return 0;
}

// K1b schedule for (N, W) on this device + the workspace layout shared with the K1 tree mode
static int bal_plan(int N, int W, BalPlan& pl) {
    int occ = 0;
    const int rc = bal_occupancy(W, &occ);
    if (rc) return rc;
    long long G = (long long)device_sms() * occ;
    const LibEnv& env = lib_env();
    const long long e = env.bal_ctas;                            // experiments (LLAMPC_BAL_CTAS): CTAs per SM x 100
    if (e > 0) G = (long long)device_sms() * e / 100;
    if (G < 1) G = 1;
    if (G > (long long)device_sms() * TREE_MAX_CTAS_PER_SM) G = (long long)device_sms() * TREE_MAX_CTAS_PER_SM;
    const long long tpw = env.bal_tpw;                           // target number of tasks per resident warp
    const long long rmin = env.bal_rmin;
    BalSched& s = pl.sch;
    s.n_wg = (N + 31) / 32;
    const long long warp_rows = (long long)s.n_wg * W;
    long long R = (warp_rows + G * 4 * tpw - 1) / (G * 4 * tpw);
    const long long rmax = env.bal_rmax;                         // bounds what a starved warp can hold back at the end
    if (R > rmax) R = rmax;
    if (R < rmin) R = rmin;
    if (R > W) R = W;
    s.SY = (int)((W + R - 1) / R);
    s.R = (W + s.SY - 1) / s.SY;                                 // even chunks
    s.SY = (W + s.R - 1) / s.R;
    const long long n_tasks = (long long)s.n_wg * s.SY;
    if (n_tasks > 0x7fffffffll - G * 8) return LLAMPC_E_RANGE;
    s.n_tasks = (int)n_tasks;
    const long long ctas = (n_tasks + 3) / 4;
    pl.grid = (int)(G < ctas ? G : ctas);
    pl.lay = tree_layout(N, s.SY > 1 ? (size_t)s.n_tasks * 32 * sizeof(float) : 0);
    return 0;
}

template <bool GEOM, bool MUFU>
static int launch_balanced(const float* bank, int N, int Npad, const float* hist, int W, double Ts, float* avg_err,
                           int idx_offset, int K, const BalPlan& pl, unsigned char* wsb, u64* out, const NewRow& nr,
                           const PeerXchg& px, cudaStream_t st) {
    auto kern = lookback_balanced_kernel<GEOM, MUFU>;
    const size_t smem = (size_t)W * (LLAMPC_HIST_ROW * 4);
    const BalWs ws = tree_workspace(wsb, pl.lay);
    const LibEnv& env = lib_env();
    if (env.bal_verbose)
        fprintf(stderr, "K1b grid=%d R=%d SY=%d tasks=%d lists=%d\n", pl.grid, pl.sch.R, pl.sch.SY, pl.sch.n_tasks, pl.sch.n_wg);
    kern<<<(unsigned)pl.grid, LB_THREADS, smem, st>>>(reinterpret_cast<const float4*>(bank), N, Npad, hist, W, make_step(Ts),
                                                      avg_err, idx_offset, nr, K, ws, pl.sch, out, px,
                                                      (env.bal_trace ? 1 : 0) | (env.bal_nofence ? 2 : 0));
    return (int)cudaGetLastError();
}

int lookback_balanced_launch(const float* bank, int N, int Npad, const float* hist, int W, double Ts, float* avg_err,
                             int idx_offset, int geom_shared, int mufu_sin, int K, void* workspace,
                             unsigned long long workspace_bytes, u64* out, const NewRow& nr, const PeerXchg& px,
                             cudaStream_t st) {
    if (!bank || !hist || !workspace || !out || N <= 0 || Npad < N) return LLAMPC_E_ARG;
    if (W <= 0 || W > LLAMPC_MAX_W || K < 0 || K > LLAMPC_LIST_LEN) return LLAMPC_E_RANGE;
    if ((reinterpret_cast<uintptr_t>(bank) & 15u) || (reinterpret_cast<uintptr_t>(hist) & 15u) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15u))
        return LLAMPC_E_ALIGN;
    BalPlan pl;
    const int prc = bal_plan(N, W, pl);
    if (prc) return prc;
    if (workspace_bytes < pl.lay.bytes) return LLAMPC_E_ARG;
    unsigned char* wsb = static_cast<unsigned char*>(workspace);
    if (mufu_sin)
        return geom_shared ? launch_balanced<true, true>(bank, N, Npad, hist, W, Ts, avg_err, idx_offset, K, pl, wsb, out, nr, px, st)
                           : launch_balanced<false, true>(bank, N, Npad, hist, W, Ts, avg_err, idx_offset, K, pl, wsb, out, nr, px, st);
    return geom_shared ? launch_balanced<true, false>(bank, N, Npad, hist, W, Ts, avg_err, idx_offset, K, pl, wsb, out, nr, px, st)
                       : launch_balanced<false, false>(bank, N, Npad, hist, W, Ts, avg_err, idx_offset, K, pl, wsb, out, nr, px, st);
}

long long lookback_balanced_workspace_bytes(int N, int W) {
    BalPlan pl;
    const int rc = bal_plan(N, W, pl);
    if (rc) return rc > 0 ? -1000 - rc : rc;       // CUDA runtime error while querying the occupancy
    return (long long)pl.lay.bytes;
}

}  // namespace llampc

using namespace llampc;

// experiments only (not declared in the public header): copies the K1b timeline of the last traced launch
extern "C" int llampc_debug_balanced_trace(unsigned long long* dst_h, int n_ctas) {
    if (!dst_h || n_ctas <= 0 || n_ctas > BAL_TRACE_CTAS) return LLAMPC_E_ARG;
    return (int)cudaMemcpyFromSymbol(dst_h, g_bal_trace, (size_t)n_ctas * 4 * sizeof(unsigned long long));
}
