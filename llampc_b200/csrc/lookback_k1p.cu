// K1p: the look-back window kernel with TWO candidates per thread in packed f32x2 arithmetic.  sm_100a.
//
// Reference behaviour being replaced: evaluate_models_vectorized (llampc/mpc/evaluate_models_vectorized.py:4-23)
// + the scoring / selection block of run_nmpc_orca_llampc_rt.py:347-360.
#include "lookback_kernels.cuh"
#include "llampc_packed.cuh"

namespace llampc {

// ---------------------------------------------------------------------------------------------------
// K1p.  K1 with TWO candidates per thread in packed f32x2 arithmetic (llampc_packed.cuh): FFMA2 / FMUL2 / FADD2 carry
// both candidates through one issue slot, the history-row values are broadcast scalar operands.  Same tiling as K1
// with twice the candidates per CTA: block = 128 threads = (128/SY candidate pairs) x (SY window splits); thread
// (c, sy) owns candidates base + c and base + 128/SY + c (both bank loads stay coalesced).
// ---------------------------------------------------------------------------------------------------
#ifndef LLAMPC_LB2_MIN_BLOCKS
#define LLAMPC_LB2_MIN_BLOCKS 4
#endif
// Optional per-CTA timeline (experiments only: `make -C llampc_b200/csrc trace`, tools/gpu_k1p_trace.py): global timer at
// CTA entry, after the history staging, after the RK4 rows, after the CTA-level selection and at the exit of warp 0, + SM id.
#ifdef LLAMPC_K1P_TRACE
constexpr int K1P_TRACE_CTAS = 8192;
__device__ unsigned long long g_k1p_trace[K1P_TRACE_CTAS * 6];
__device__ __forceinline__ unsigned long long k1p_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define K1P_TRACE(slot)                                                                                   \
    do {                                                                                                  \
        if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < K1P_TRACE_CTAS)                           \
            g_k1p_trace[blockIdx.x * 6 + (slot)] = k1p_gtimer();                                          \
    } while (0)
#else
#define K1P_TRACE(slot) do { } while (0)
#endif
template <int SY, bool GEOM_SHARED, bool MUFU_SIN, bool WIDE>
__global__ void __launch_bounds__(LB_THREADS, LLAMPC_LB2_MIN_BLOCKS)
lookback_window2_kernel(const float4* __restrict__ bank, int N, int Npad, const float* __restrict__ hist, int W,
                        long hist_stride_floats, StepSize z, float* __restrict__ avg_err,
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
    K1P_TRACE(0);
#ifdef LLAMPC_K1P_TRACE
    if (tid == 0 && blockIdx.y == 0 && blockIdx.x < K1P_TRACE_CTAS) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_k1p_trace[blockIdx.x * 6 + 5] = smid;
    }
#endif

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
    K1P_TRACE(1);
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
    int w = sy;
    for (; w < W; w += SY) {
        HistRow r;
        r.q0 = srow[w * 5 + 0];
        r.q1 = srow[w * 5 + 1];
        r.q2 = srow[w * 5 + 2];
        r.q3 = srow[w * 5 + 3];
        r.q4 = srow[w * 5 + 4];
        bool ok0, ok1;
        const F2 e = lookback_step_fast2<GEOM_SHARED, MUFU_SIN, WIDE>(p, r, z, ok0, ok1);
        float e0, e1;
        up(e, e0, e1);
        if (!ok0) e0 = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, i0, srow + w * 5, z);
        if (!ok1) e1 = lookback_step_general<GEOM_SHARED, MUFU_SIN>(bank, Npad, i1, srow + w * 5, z);
        acc0 += e0;
        acc1 += e1;
    }

    K1P_TRACE(2);
    // Programmatic dependent launch (LLAMPC_LB_FLAG_PDL; both instructions are no-ops in an ordinary launch): the next launch
    // on the stream may start its bank loads and RK4 rows now, beside this launch's selection / merge-tree tail; this launch
    // in turn must not write avg_err, the workspace or `out` before the previous one has completed.
    pdl_launch_dependents();
    pdl_wait();
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
        K1P_TRACE(3);
        if (tid < 32) tree_merge(key, tid, (int)blockIdx.x, (int)gridDim.x, tm.K, tm.ws, mrows, tm.out, px);
        K1P_TRACE(4);
        return;
    }
    cta_select_emit<KW, true>(key, skeys, v, cta_lists);
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
        merge_lists_device<LB_THREADS>(cta_lists + (size_t)v * gridDim.x * LLAMPC_LIST_LEN, gridDim.x, fm.K, outv, msm);
        if (tid == 0) fm.ticket[v] = 0;
        if (px.world > 1 && tid < 32) {
            __syncwarp();
            const u64 mine = __shfl_sync(0xffffffffu, tid == 0 ? *reinterpret_cast<volatile u64*>(outv) : 0ull, 0);
            const u64 g = peer_minloc(px, mine, tid);
            if (tid == 0) outv[0] = g;
        }
    }
}


template <int SY, bool GEOM, bool MUFU, bool WIDE>
static int launch_one(const LbArgs& a, cudaStream_t st) {
    auto kern = lookback_window2_kernel<SY, GEOM, MUFU, WIDE>;
    const size_t smem = (size_t)a.W * (LLAMPC_HIST_ROW * 4) + (SY > 1 ? LB_THREADS * 4 * 2 : 0);
    const int rc = raise_dynamic_smem(kern, smem);
    if (rc) return rc;
    const int CPB = (2 * LB_THREADS) / SY;
    dim3 grid((a.N + CPB - 1) / CPB, a.n_vehicles);
    if (a.pdl)
        return issue_pdl(kern, grid, dim3(LB_THREADS), smem, st, a.bank, a.N, a.Npad, a.hist, a.W, a.hist_stride_floats, a.z,
                         a.avg_err, a.cta_lists, a.idx_offset, a.nr, a.fm, a.px, a.tm);
    return issue(kern, grid, dim3(LB_THREADS), smem, st, a.bank, a.N, a.Npad, a.hist, a.W, a.hist_stride_floats, a.z, a.avg_err,
                 a.cta_lists, a.idx_offset, a.nr, a.fm, a.px, a.tm);
}

template <int SY, bool WIDE>
static int launch_gm(const LbArgs& a, bool geom, bool mufu, cudaStream_t st) {
    if (mufu) return geom ? launch_one<SY, true, true, WIDE>(a, st) : launch_one<SY, false, true, WIDE>(a, st);
    return geom ? launch_one<SY, true, false, WIDE>(a, st) : launch_one<SY, false, false, WIDE>(a, st);
}

// a.wide: the instantiation whose step takes its slip angles from the full-range atan (windows measured at low speed)
template <int SY>
static int launch_sy(const LbArgs& a, bool geom, bool mufu, cudaStream_t st) {
    return a.wide ? launch_gm<SY, true>(a, geom, mufu, st) : launch_gm<SY, false>(a, geom, mufu, st);
}

int launch_k1_packed(const LbArgs& a, int sy, bool geom, bool mufu, cudaStream_t st) {
    switch (sy) {
        case 1: return launch_sy<1>(a, geom, mufu, st);
        case 2: return launch_sy<2>(a, geom, mufu, st);
        case 4: return launch_sy<4>(a, geom, mufu, st);
        case 8: return launch_sy<8>(a, geom, mufu, st);
        case 16: return launch_sy<16>(a, geom, mufu, st);
        default: return LLAMPC_E_ARG;
    }
}

}  // namespace llampc

#ifdef LLAMPC_K1P_TRACE
// experiments only (not declared in the public header): copies the K1p timeline of the last traced launch
extern "C" int llampc_debug_k1p_trace(unsigned long long* dst_h, int n_ctas) {
    if (!dst_h || n_ctas <= 0 || n_ctas > llampc::K1P_TRACE_CTAS) return LLAMPC_E_ARG;
    return (int)cudaMemcpyFromSymbol(dst_h, llampc::g_k1p_trace, (size_t)n_ctas * 6 * sizeof(unsigned long long));
}
#endif
