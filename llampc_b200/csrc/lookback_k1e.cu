// K1e: the packed look-back window kernel on EQUAL warp shares (opt-in: LLAMPC_KERNEL_K1E).  sm_100a.
//
// Same arithmetic and same reference behaviour as K1p (lookback_k1p.cu): evaluate_models_vectorized
// (llampc/mpc/evaluate_models_vectorized.py:4-23) + errors / window mean / argmin / argsort[:K] of
// run_nmpc_orca_llampc_rt.py:349-360.  What changes is how the N x W candidate-steps are laid on the machine.
//
// K1p's grid is a property of the problem (65,536 x 50: 1,024 CTAs on 592 resident slots = 1.73 waves).  K1e launches
// ONE 512-thread CTA per SM -- four warps per scheduler, all the registers of the SM -- and gives every WARP the same
// number of warp-steps: the (group of 64 candidates) x (window row) space is linearised group-major and cut into
// n_warps contiguous ranges of floor / ceil(L / n_warps) steps.  The parameters of a range's next group are fetched into
// shared memory with cp.async while the current one integrates.  A group whose rows are spread over several warps is
// combined through an L2-resident scratch: every warp stores its 64 partial sums, the last one to arrive (one acq_rel
// atomic per arrival) adds the partials in row order (fixed order: run-to-run deterministic) and finalises the group.
// Every warp keeps a running top-16; the CTA merges its lists at exit and enters the same merge tree as K1 / K1p / K1b.
//
// What the per-warp timelines of this kernel showed (profiles/r02_k1e_equal_share.md), and why K1p stays the default:
//   * warps of DIFFERENT CTAs that share a scheduler are served by CTA age: with four 128-thread CTAs per SM and equal
//     shares the oldest CTA's warps finish after 22 us, the youngest after 39 us (65,536 x 50), whatever the warp index;
//   * warps of the SAME CTA are served fairly: with one 512-thread CTA per SM all 16 warps finish within +-2 us;
//   * either way a scheduler completes one warp-step (64 candidate-steps) per ~790 clocks with four resident warps, ~865
//     with two -- against 548 FMA-pipe clocks of arithmetic: the step is a chain of dependent packed operations
//     (stall_wait is the top sampled stall), and the 110+ registers of the packed step allow no fifth warp.
// 86.5 warp-steps per scheduler x 790 clocks = 34.8 us is therefore the floor of the 65,536 x 50 tick on ANY tiling; K1p's
// hardware-dispatched grid already sits on it (rows end at 38 us, span 43 us) and K1e, 2 - 5 us slower at every size
// from 65,536 to 1,048,576 candidates, is kept as a measured alternative, not as the default.
#include "lookback_kernels.cuh"
#include "llampc_packed.cuh"
#include <stdio.h>

namespace llampc {

#ifndef LLAMPC_EQ_THREADS
#define LLAMPC_EQ_THREADS 512
#endif
constexpr int EQ_THREADS = LLAMPC_EQ_THREADS;
#ifndef LLAMPC_LB2_MIN_BLOCKS_E
#define LLAMPC_LB2_MIN_BLOCKS_E (512 / LLAMPC_EQ_THREADS)
#endif
constexpr int EQ_WARPS = EQ_THREADS / 32;
constexpr int EQ_GROUP = 64;                       // candidates per warp-step (two per lane, packed)
constexpr int EQ_STAGE_BYTES = 2 * 8 * 32 * 16;    // per warp: two buffers of 8 float4 per lane (the next group's parameters)

struct EqSched {
    long long L;        // warp-steps = groups x W
    int n_warps;        // warps that take a share (= resident warps unless the problem is tiny)
    int n_groups;       // ceil(N / 64)
    int maxch;          // scratch rows per group (upper bound of the warps a group can be spread over)
    int stagger;        // experiments: warps sharing a scheduler start this many clocks apart
};

#ifdef LLAMPC_K1E_TRACE
constexpr int K1E_TRACE_WARPS = 8192;
__device__ unsigned long long g_k1e_trace[K1E_TRACE_WARPS * 6];
__device__ __forceinline__ unsigned long long k1e_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define K1E_TRACE(slot)                                                                        \
    do {                                                                                       \
        if (lane == 0 && gw < K1E_TRACE_WARPS) g_k1e_trace[gw * 6 + (slot)] = k1e_gtimer();   \
    } while (0)
#else
#define K1E_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool GEOM_SHARED, bool MUFU_SIN>
__global__ void __launch_bounds__(EQ_THREADS, LLAMPC_LB2_MIN_BLOCKS_E)
lookback_equal_kernel(const float4* __restrict__ bank, int N, int Npad, const float* __restrict__ hist, int W, StepSize z,
                      float* __restrict__ avg_err, int idx_offset, NewRow nr, int K, BalWs ws, EqSched sch,
                      u64* __restrict__ out, PeerXchg px) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ u64 mrows[BAL_FAN][BAL_ROW_PAD];
    __shared__ u64 sfin[EQ_WARPS][LLAMPC_LIST_LEN];
    float4* srow = reinterpret_cast<float4*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bytes = (unsigned)W * (LLAMPC_HIST_ROW * 4);
    float4* stage = reinterpret_cast<float4*>(smem_raw + ((bytes + 127u) & ~127u) + (size_t)warp * EQ_STAGE_BYTES);
    const int gw = (int)blockIdx.x * EQ_WARPS + warp;
    K1E_TRACE(0);
    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, bytes);
        tma_bulk_g2s(srow, hist, bytes, &mbar);
    }
    __syncthreads();

    // this warp's share of the linearised (group, row) space
    // (small problems: only the first n_warps warps of the grid take a share, each of at least four steps)
    long long lo = gw < sch.n_warps ? (long long)gw * sch.L / sch.n_warps : 0;
    const long long hi = gw < sch.n_warps ? (long long)(gw + 1) * sch.L / sch.n_warps : 0;
    int g = (int)(lo / W);
    int r = (int)(lo - (long long)g * W);
    Cand2 p;
    {
        const int gg = g < sch.n_groups ? g : 0;
        p = load_cand2(bank, Npad, min(gg * EQ_GROUP + lane, N - 1), min(gg * EQ_GROUP + 32 + lane, N - 1));   // overlaps the bulk copy
    }
    mbar_wait(&mbar, 0);
    if (nr.slot >= 0) {                            // uniform over the grid
        if (tid < LLAMPC_HIST_ROW / 4) {
            const float4 q = make_float4(nr.v[4 * tid], nr.v[4 * tid + 1], nr.v[4 * tid + 2], nr.v[4 * tid + 3]);
            srow[nr.slot * 5 + tid] = q;
            if (blockIdx.x == 0) reinterpret_cast<float4*>(const_cast<float*>(hist))[nr.slot * 5 + tid] = q;
        }
        __syncthreads();
    }
    if (sch.stagger > 0 && warp >= 4) {            // de-synchronise the warps that share a scheduler
        const long long t0 = clock64(), d = (long long)(warp >> 2) * sch.stagger;
        while (clock64() - t0 < d) { }
    }
    K1E_TRACE(1);

#ifdef LLAMPC_K1E_TRACE
    unsigned long long n_fallback = 0;
#endif
    u64 run = ~0ull;                               // this warp's running top-16 (lanes 0..15 ascending, ~0 above)
    int buf = 0;
    bool staged = false;
    const float scale = 0.25f / (float)W;          // errors = mean over the 4 scored states (rt.py:349), avg over the window (rt.py:357)
    while (lo < hi) {                              // uniform over the warp; no CTA barrier until the exit
        const long long left = hi - lo;
        const int r1 = left < (long long)(W - r) ? r + (int)left : W;
        const int c0 = g * EQ_GROUP + lane, c1 = c0 + 32;
        const bool valid0 = c0 < N, valid1 = c1 < N;
        const int i0 = valid0 ? c0 : N - 1, i1 = valid1 ? c1 : N - 1;
        if (staged) {                              // parameters fetched while the previous group was integrated
            cp_async_wait_all();
            const float4* s = stage + (buf ^ 1) * (8 * 32) + lane;
            p = make_cand2(cand_from_groups(s[0], s[32], s[64], s[96]), cand_from_groups(s[128], s[160], s[192], s[224]));
        }
        staged = lo + (r1 - r) < hi;
        if (staged) {                              // the next group's parameters -> this lane's slots of the other buffer
            const int n0 = min(c0 + EQ_GROUP, N - 1), n1 = min(c1 + EQ_GROUP, N - 1);
            float4* s = stage + buf * (8 * 32) + lane;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cp_async16(s + 32 * j, bank + (size_t)j * Npad + n0);
                cp_async16(s + 32 * (4 + j), bank + (size_t)j * Npad + n1);
            }
            cp_async_commit();
            buf ^= 1;
        }

        float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
        for (int w = r; w < r1; ++w) {
            HistRow h;
            h.q0 = srow[w * 5 + 0];
            h.q1 = srow[w * 5 + 1];
            h.q2 = srow[w * 5 + 2];
            h.q3 = srow[w * 5 + 3];
            h.q4 = srow[w * 5 + 4];
            bool ok0, ok1;
            const F2 e = lookback_step_fast2<GEOM_SHARED, MUFU_SIN>(p, h, z, ok0, ok1);
            float e0, e1;
            up(e, e0, e1);
#ifdef LLAMPC_K1E_TRACE
            if (__any_sync(0xffffffffu, !ok0 || !ok1)) ++n_fallback;
#endif
            if (!ok0) e0 = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, i0, srow + w * 5, z);
            if (!ok1) e1 = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, i1, srow + w * 5, z);
            acc0 += e0;
            acc1 += e1;
        }

        bool fin = true;
        if (r != 0 || r1 != W) {                   // the group's window is spread over several warps
            const long long x0 = (long long)g * W;
            const int wf = (int)(((x0 + 1) * sch.n_warps - 1) / sch.L);           // warp that holds row 0 of the group
            const int wl = (int)(((x0 + W) * sch.n_warps - 1) / sch.L);           // ... and row W - 1
            const int nch = wl - wf + 1;
            float* part = ws.part + ((size_t)g * sch.maxch) * EQ_GROUP;
            __stcg(part + (size_t)(gw - wf) * EQ_GROUP + lane, acc0);
            __stcg(part + (size_t)(gw - wf) * EQ_GROUP + 32 + lane, acc1);
            __syncwarp();
            unsigned old = 0;
            // one acq_rel atomic: its release half (cumulative over the warp barrier) publishes the partial sums before
            // the count, its acquire half orders the loads below (issued after the shuffle) after the count
            if (lane == 0) old = atom_add_acq_rel_gpu(ws.gcount + g, 1u);
            old = __shfl_sync(0xffffffffu, old, 0);
            fin = old == (unsigned)(nch - 1);      // the last chunk to arrive finalises the group
            if (fin) {
                if (lane == 0) ws.gcount[g] = 0;   // ready for the next launch on the same stream
                acc0 = __ldcg(part + lane);
                acc1 = __ldcg(part + 32 + lane);
                for (int j = 1; j < nch; ++j) {    // row order
                    acc0 += __ldcg(part + (size_t)j * EQ_GROUP + lane);
                    acc1 += __ldcg(part + (size_t)j * EQ_GROUP + 32 + lane);
                }
            }
        }
        if (fin) {
            const float err0 = acc0 * scale, err1 = acc1 * scale;
            u64 k0 = ~0ull, k1 = ~0ull;
            if (valid0) {
                if (avg_err) avg_err[c0] = err0;
                k0 = pack_key(err0, (unsigned)(idx_offset + c0));
            }
            if (valid1) {
                if (avg_err) avg_err[c1] = err1;
                k1 = pack_key(err1, (unsigned)(idx_offset + c1));
            }
            k0 = warp_sort_u64(k0, lane);
            k1 = warp_sort_u64(k1, lane);
            k0 = warp_merge_low32(k0, __shfl_sync(0xffffffffu, k1, 31 - lane), lane);
            run = warp_merge_low32(run, __shfl_sync(0xffffffffu, k0, 31 - lane), lane);
            if (lane >= LLAMPC_LIST_LEN) run = ~0ull;
        }
        lo += r1 - r;
        ++g;
        r = 0;
    }
    K1E_TRACE(2);
#ifdef LLAMPC_K1E_TRACE
    if (lane == 0 && gw < K1E_TRACE_WARPS) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_k1e_trace[gw * 6 + 4] = smid;
        g_k1e_trace[gw * 6 + 5] = n_fallback;
    }
#endif
    // CTA list = merge of its four warps' running lists; then the tree over the CTA lists
    if (lane < LLAMPC_LIST_LEN) sfin[warp][lane] = run;
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int w = 1; w < EQ_WARPS; ++w) {
        const u64 other_rev = lane >= LLAMPC_LIST_LEN ? sfin[w][31 - lane] : ~0ull;
        run = warp_merge_low32(run, other_rev, lane);
        if (lane >= LLAMPC_LIST_LEN) run = ~0ull;
    }
    tree_merge(run, lane, (int)blockIdx.x, (int)gridDim.x, K, ws, mrows, out, px);
    K1E_TRACE(3);
}

static inline size_t eq_smem_bytes(int W) {
    return (((size_t)W * (LLAMPC_HIST_ROW * 4) + 127) & ~(size_t)127) + (size_t)EQ_WARPS * EQ_STAGE_BYTES;
}

// resident CTAs per SM for a window of W rows (the four variants have the same footprint); cached per device
static int eq_occupancy(int W, int* occ_out) {
    struct Entry { int w, occ; };
    static Entry cache[64] = {};
    int dev = 0;
    LLAMPC_CUDA_TRY(cudaGetDevice(&dev));
    Entry local = {0, 0};
    Entry& e = (dev >= 0 && dev < 64) ? cache[dev] : local;
    if (W != e.w) {
        const size_t smem = eq_smem_bytes(W);
        LLAMPC_CUDA_TRY((cudaError_t)raise_dynamic_smem(lookback_equal_kernel<true, true>, smem));
        LLAMPC_CUDA_TRY((cudaError_t)raise_dynamic_smem(lookback_equal_kernel<true, false>, smem));
        LLAMPC_CUDA_TRY((cudaError_t)raise_dynamic_smem(lookback_equal_kernel<false, true>, smem));
        LLAMPC_CUDA_TRY((cudaError_t)raise_dynamic_smem(lookback_equal_kernel<false, false>, smem));
        int o = 0;
        LLAMPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, lookback_equal_kernel<true, true>, EQ_THREADS, smem));
        if (o <= 0) return LLAMPC_E_RANGE;
        e.occ = o < TREE_MAX_CTAS_PER_SM ? o : TREE_MAX_CTAS_PER_SM;
        e.w = W;
    }
    *occ_out = e.occ;
    return 0;
}

struct EqPlan { int grid; EqSched sch; TreeLayout lay; };

static int eq_plan(int N, int W, EqPlan& pl) {
    int occ = 0;
    const int rc = eq_occupancy(W, &occ);
    if (rc) return rc;
    EqSched& s = pl.sch;
    s.n_groups = (N + EQ_GROUP - 1) / EQ_GROUP;
    s.L = (long long)s.n_groups * W;
    long long G = (long long)device_sms() * occ;
    const long long e = lib_env().eq_ctas;                       // experiments (LLAMPC_EQ_CTAS): CTAs per SM x 100
    if (e > 0) G = (long long)device_sms() * e / 100;
    const long long cap = (s.L + 4 * EQ_WARPS - 1) / (4 * EQ_WARPS);   // at least four warp-steps per warp
    if (G > cap) G = cap;
    if (G < 1) G = 1;
    pl.grid = (int)G;
    s.n_warps = pl.grid * EQ_WARPS;
    if (s.n_warps > s.L / 4) s.n_warps = s.L / 4 > 0 ? (int)(s.L / 4) : 1;      // every sharing warp gets >= 4 steps (or all of them)
    const long long per = s.L / s.n_warps > 0 ? s.L / s.n_warps : 1;      // every warp's range holds >= per steps
    long long maxch = (W + per - 1) / per + 1;
    if (maxch > s.n_warps) maxch = s.n_warps;
    s.maxch = (int)maxch;
    s.stagger = (int)lib_env().eq_stagger;
    pl.lay = tree_layout(N, (size_t)s.n_groups * s.maxch * EQ_GROUP * sizeof(float));
    return 0;
}

template <bool GEOM, bool MUFU>
static int launch_equal(const float4* bank, int N, int Npad, const float* hist, int W, const StepSize& z, float* avg_err,
                        int idx_offset, int K, const EqPlan& pl, unsigned char* wsb, u64* out, const NewRow& nr,
                        const PeerXchg& px, cudaStream_t st) {
    auto kern = lookback_equal_kernel<GEOM, MUFU>;
    const size_t smem = eq_smem_bytes(W);
    const int rc = raise_dynamic_smem(kern, smem);
    if (rc) return rc;
    const BalWs ws = tree_workspace(wsb, pl.lay);
    return issue(kern, dim3((unsigned)pl.grid), dim3(EQ_THREADS), smem, st, bank, N, Npad, hist, W, z, avg_err, idx_offset,
                 nr, K, ws, pl.sch, out, px);
}

int lookback_equal_launch(const float4* bank, int N, int Npad, const float* hist, int W, const StepSize& z, float* avg_err,
                          int idx_offset, bool geom, bool mufu, int K, void* workspace, unsigned long long workspace_bytes,
                          u64* out, const NewRow& nr, const PeerXchg& px, cudaStream_t st) {
    if (!bank || !hist || !workspace || !out || N <= 0 || Npad < N) return LLAMPC_E_ARG;
    if (W <= 0 || W > LLAMPC_MAX_W || K < 0 || K > LLAMPC_LIST_LEN) return LLAMPC_E_RANGE;
    EqPlan pl;
    const int prc = eq_plan(N, W, pl);
    if (prc) return prc;
    if (workspace_bytes < pl.lay.bytes) return LLAMPC_E_ARG;
    unsigned char* wsb = static_cast<unsigned char*>(workspace);
    if (mufu)
        return geom ? launch_equal<true, true>(bank, N, Npad, hist, W, z, avg_err, idx_offset, K, pl, wsb, out, nr, px, st)
                    : launch_equal<false, true>(bank, N, Npad, hist, W, z, avg_err, idx_offset, K, pl, wsb, out, nr, px, st);
    return geom ? launch_equal<true, false>(bank, N, Npad, hist, W, z, avg_err, idx_offset, K, pl, wsb, out, nr, px, st)
                : launch_equal<false, false>(bank, N, Npad, hist, W, z, avg_err, idx_offset, K, pl, wsb, out, nr, px, st);
}

int lookback_equal_plan(int N, int W, int* grid, int* block, size_t* bytes, TreeLayout* lay) {
    EqPlan pl;
    const int rc = eq_plan(N, W, pl);
    if (rc) return rc;
    if (grid) *grid = pl.grid;
    if (block) *block = EQ_THREADS;
    if (bytes) *bytes = pl.lay.bytes;
    if (lay) *lay = pl.lay;
    return 0;
}

}  // namespace llampc

#ifdef LLAMPC_K1E_TRACE
// experiments only (not declared in the public header): copies the K1e per-warp timeline of the last traced launch
extern "C" int llampc_debug_k1e_trace(unsigned long long* dst_h, int n_warps) {
    if (!dst_h || n_warps <= 0 || n_warps > llampc::K1E_TRACE_WARPS) return LLAMPC_E_ARG;
    return (int)cudaMemcpyFromSymbol(dst_h, llampc::g_k1e_trace, (size_t)n_warps * 6 * sizeof(unsigned long long));
}
#endif
