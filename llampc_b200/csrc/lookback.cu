// Look-back C ABI (llampc_lookback_plan / _launch / _tick / _push), the list-merge, top-K and fp64 re-score kernels, and
// the one-step batch kernels.  sm_100a.  The scoring kernels live in lookback_k1.cu (K1), lookback_k1p.cu (K1p),
// lookback_rolling.cu (K1r, K1v) and lookback_balanced.cu (K1b).
//
// Reference behaviour being replaced: evaluate_models_vectorized (llampc/mpc/evaluate_models_vectorized.py:4-23)
// + the scoring / selection block of run_nmpc_orca_llampc_rt.py:347-360.
#include "lookback_kernels.cuh"
#include <chrono>
#include <stdio.h>
#include "llampc_model_f64.cuh"
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

namespace llampc {

LaunchCollector& launch_collector() {
    static thread_local LaunchCollector lc = {nullptr, 0};
    return lc;
}

static long long env_ll(const char* name, long long dflt) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    const long long v = atoll(e);
    return v > 0 ? v : dflt;
}

const LibEnv& lib_env() {
    static const LibEnv env = [] {
        LibEnv e;
        const char* g = getenv("LLAMPC_TICK_GRAPH");
        e.tick_graph = (g && g[0] == '0') ? 0 : 1;
        e.bal_ctas = env_ll("LLAMPC_BAL_CTAS", 0);
        e.bal_tpw = env_ll("LLAMPC_BAL_TPW", 6);
        e.bal_rmin = env_ll("LLAMPC_BAL_RMIN", 1);
        e.bal_rmax = env_ll("LLAMPC_BAL_RMAX", 10);
        e.bal_verbose = getenv("LLAMPC_BAL_VERBOSE") != nullptr;
        e.bal_trace = getenv("LLAMPC_BAL_TRACE") != nullptr;
        e.bal_nofence = getenv("LLAMPC_BAL_NOFENCE") != nullptr;
        e.eq_ctas = env_ll("LLAMPC_EQ_CTAS", 0);
        e.eq_stagger = env_ll("LLAMPC_EQ_STAGGER", 0);
        return e;
    }();
    return env;
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
// K4' stand-alone merge of the per-CTA sorted lists written by K1 / K1r into the global top-K (K <= LLAMPC_LIST_LEN),
// one CTA per vehicle; the algorithm is merge_lists_device (llampc_common.cuh).  Used when a launch produces more
// than 1,024 lists (otherwise the last CTA of K1 runs the same routine itself); optionally carries the NVLink
// min-loc exchange.  out[0] = arg-min key, out[1..K] = ascending top-K.
// ---------------------------------------------------------------------------------------------------
template <int MERGE_THREADS>
__global__ void __launch_bounds__(MERGE_THREADS)
topk_merge_lists_kernel(const u64* __restrict__ lists, int n_lists, int K, u64* __restrict__ out, PeerXchg px) {
    __shared__ MergeSmem<MERGE_THREADS> sm;
    const int v = blockIdx.x;                      // vehicle
    u64* outv = out + (size_t)v * (LLAMPC_LIST_LEN + 1);
    merge_lists_device<MERGE_THREADS>(lists + (size_t)v * n_lists * LLAMPC_LIST_LEN, n_lists, K, outv, sm);
    if (px.world > 1 && threadIdx.x < 32) {        // warp 0: min-loc across the GPUs of the box (single vehicle)
        __syncwarp();
        const u64 mine = __shfl_sync(0xffffffffu, threadIdx.x == 0 ? *reinterpret_cast<volatile u64*>(outv) : 0ull, 0);
        const u64 g = peer_minloc(px, mine, threadIdx.x);
        if (threadIdx.x == 0) outv[0] = g;
    }
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
    // In a pipelined replay this kernel AND the next tick's scoring kernel are launched with programmatic stream serialization:
    // this grid may be scheduled as soon as every CTA of its own tick's scoring kernel has finished its RK4 rows; it passes the
    // permission on at once (launch_dependents: the next tick's scoring kernel starts its bank loads and rows beside the
    // merge-tree tail of this tick's and beside this re-score) and then waits for its scoring kernel to complete (wait) before
    // it reads the finalists.  The next scoring kernel writes nothing this kernel reads before its own griddepcontrol.wait,
    // which holds until this grid has completed.  Both instructions are no-ops in ordinary launches.
    pdl_launch_dependents();
    pdl_wait();
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
            // fc.src = [arg-min key | Kt keys | Kt scores]; exchange the 2 Kt finalist words with every peer.  Every
            // 64-bit payload travels as two self-validating 8-byte words (low half | seq << 32, high half | seq << 32):
            // no system-scope fence between data and flag, the receiver polls the words themselves (one NVLink traversal)
            const int npay = 2 * pg.kt, wpr = 2 * npay, parity = pg.seq & 1;
            const size_t my_slot = ((size_t)parity * pg.world + pg.rank) * wpr;
            const u64 tag = (u64)pg.seq << 32;
            if (tid == 0) peer_fail = 0;
            for (int i = tid; i < npay * pg.world; i += RF_THREADS) {
                const int q = i / npay, j = i % npay;
                const u64 v = __ldcg(fc.src + 1 + j);
                volatile u64* dst = reinterpret_cast<volatile u64*>(pg.peers[q]) + my_slot + 2 * j;
                dst[0] = tag | (v & 0xffffffffull);
                dst[1] = tag | (v >> 32);
            }
            __syncthreads();
            fc.dst_host[0] = __ldcg(fc.src);
            volatile u64* own = pg.peers[pg.rank] + (size_t)parity * pg.world * wpr;
            for (int i = tid; i < npay * pg.world; i += RF_THREADS) {
                const int q = i / npay, j = i % npay;
                volatile u64* src = own + (size_t)q * wpr + 2 * j;
                const long long t0 = clock64();
                u64 w0, w1;
                bool ok = true;
                for (;;) {
                    w0 = src[0];
                    w1 = src[1];
                    if ((w0 >> 32) == (u64)pg.seq && (w1 >> 32) == (u64)pg.seq) break;
                    if (clock64() - t0 > 2000000000ll) { ok = false; break; }      // ~1 s: the peer never arrived
                }
                if (!ok) peer_fail = 1;
                fc.dst_host[1 + i] = ok ? ((w1 << 32) | (w0 & 0xffffffffull)) : ~0ull;
            }
            __syncthreads();
            if (peer_fail) {                                           // poison everything: keys ~0, scores NaN
                for (int i = tid; i < npay * pg.world; i += RF_THREADS) fc.dst_host[1 + i] = ~0ull;
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

// K1p (two candidates per thread, packed f32x2) is the default for launches that can fill the GPU with pair-threads;
// small problems keep one candidate per thread (twice the threads: measured 12.4 us against 14.2 us per tick at
// 1,024 x 20).  The count that matters is candidates x vehicles of the launch.
constexpr long K1P_MIN_CANDIDATES = 8192;
constexpr int MAX_MERGE_LISTS = LB_THREADS * MERGE_LPT;          // lists one CTA can merge in-kernel (1,024)

static int choose_split(int N, int W, int n_vehicles, bool packed) {
    // SM time ~ (CTAs on the busiest SM) x (rows per thread) while the FMA pipe is the limiter.
    const int sms = device_sms();
    int best = 1;
    long best_cost = -1;
    for (int sy = 1; sy <= 16; sy *= 2) {          // 8 and 16 only pay off for launches too small to fill the GPU
        if (sy > W) break;
        const int cpc = k1_cands_per_cta(packed, sy);
        const long ctas = (long)n_vehicles * ((N + cpc - 1) / cpc);
        const long cost = ((ctas + sms - 1) / sms) * ((W + sy - 1) / sy);
        // a finer split must win by > 3 %: it doubles the number of per-CTA lists the top-K merge has to read
        if (best_cost < 0 || cost * 100 < best_cost * 97) { best_cost = cost; best = sy; }
    }
    return best;
}

// Everything llampc_lookback_launch decides before launching; llampc_lookback_plan reports the public part.
struct LbPlan {
    int kernel, split, sine, launches;
    int grid_x, grid_y, block;
    int n_lists;                 // per vehicle, grid path
    bool in_kernel_merge;        // grid path: the last CTA of a vehicle merges its lists
    bool tree;                   // single history: in-kernel merge tree (K1 / K1P / K1B)
    size_t off_lists;            // grid path workspace: [n_vehicles] tickets, then the lists
    size_t bytes;
    TreeLayout lay;
};

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

static int resolve_plan(const llampc_lookback_desc_t& d, LbPlan& p) {
    if (!d.bank || d.N <= 0 || d.Npad < d.N || d.n_vehicles <= 0) return LLAMPC_E_ARG;
    if (d.W <= 0 || d.W > LLAMPC_MAX_W || d.n_vehicles > 65535 || d.K < 0 || d.K > LLAMPC_LIST_LEN) return LLAMPC_E_RANGE;
    if (!aligned16(d.bank) || (d.hist && !aligned16(d.hist))) return LLAMPC_E_ALIGN;
    if ((d.K > 0) != (d.out != nullptr)) return LLAMPC_E_ARG;
    if (d.K == 0 && !d.avg_err && !(d.mode == LLAMPC_LB_ROLLING && !d.emit)) return LLAMPC_E_ARG;
    if (d.row32_h && d.n_vehicles != 1) return LLAMPC_E_ARG;
    if (d.peer_bufs && (d.world < 2 || d.world > 32 || d.rank < 0 || d.rank >= d.world)) return LLAMPC_E_RANGE;
    p = LbPlan{};
    p.block = LB_THREADS;
    p.launches = 1;
    p.split = 1;
    switch (d.sine) {
        case LLAMPC_SIN_SFU: case LLAMPC_SIN_STRICT: p.sine = d.sine; break;
        // MUFU.SIN holds its rated absolute error (2^-21.4) on [-pi, pi]; beyond it the error grows with the argument
        // (DESIGN.md section 4), so banks whose tyre-sine argument |C atan(.)| <= |C| pi/2 can leave that range run strict
        case LLAMPC_SIN_AUTO: p.sine = (d.sin_arg_max > 0.0f && d.sin_arg_max <= 3.14159274f) ? LLAMPC_SIN_SFU : LLAMPC_SIN_STRICT; break;
        default: return LLAMPC_E_ARG;
    }
    if (d.mode == LLAMPC_LB_ROLLING) {
        if (!d.err_ring || d.slot < 0 || d.slot >= d.W || (!d.hist && !d.row32_h) || d.peer_bufs) return LLAMPC_E_ARG;
        const bool selecting = d.emit && d.K > 0;
        const bool k1v_ok = !d.row32_h && d.N <= RV_MAX_N && d.Npad % 4 == 0 && aligned16(d.err_ring) && (!d.emit || selecting);
        p.kernel = d.kernel ? d.kernel : (k1v_ok ? LLAMPC_KERNEL_K1V : LLAMPC_KERNEL_K1R);
        if (p.kernel == LLAMPC_KERNEL_K1V) {
            if (!k1v_ok) return LLAMPC_E_ARG;
            p.grid_x = d.n_vehicles; p.grid_y = 1; p.block = RV_THREADS;
            p.bytes = 0;
            return 0;
        }
        if (p.kernel != LLAMPC_KERNEL_K1R) return LLAMPC_E_ARG;
        p.n_lists = (d.N + LB_THREADS - 1) / LB_THREADS;
        if (p.n_lists > MAX_MERGE_LISTS * MERGE_LPT) return LLAMPC_E_RANGE;
        p.in_kernel_merge = selecting && p.n_lists <= MAX_MERGE_LISTS;
        if (selecting && !p.in_kernel_merge) p.launches = 2;
        p.grid_x = p.n_lists; p.grid_y = d.n_vehicles;
        p.off_lists = up256((size_t)d.n_vehicles * sizeof(unsigned));
        p.bytes = p.off_lists + up256((size_t)d.n_vehicles * p.n_lists * LLAMPC_LIST_LEN * sizeof(u64));
        return 0;
    }
    if (d.mode != LLAMPC_LB_RECOMPUTE) return LLAMPC_E_ARG;
    if (!d.hist || d.hist_stride_rows < d.W) return LLAMPC_E_ARG;
    if (((long)d.hist_stride_rows * LLAMPC_HIST_ROW * 4) % 16) return LLAMPC_E_ALIGN;
    const long total = (long)d.N * d.n_vehicles;
    if (d.n_vehicles > 1 && d.peer_bufs) return LLAMPC_E_ARG;
    int kernel = d.kernel;
    bool packed = kernel ? kernel == LLAMPC_KERNEL_K1P : total >= K1P_MIN_CANDIDATES;
    int sy = d.split ? d.split : choose_split(d.N, d.W, d.n_vehicles, packed);
    if (sy > d.W) sy = 1;
    if (sy != 1 && sy != 2 && sy != 4 && sy != 8 && sy != 16) return LLAMPC_E_ARG;
    const int cpc = k1_cands_per_cta(packed, sy);
    p.n_lists = (d.N + cpc - 1) / cpc;
    p.split = sy;
    p.grid_x = p.n_lists; p.grid_y = d.n_vehicles;
    if (d.n_vehicles == 1 && d.K > 0) {
        // ---- one history: in-kernel merge tree.  K1b pays off only where K1 leaves SMs idle AND every K1 thread has a
        // long serial walk (measured on B200: 4,096 candidates x 1,024 rows 115 us against 133 us)
        p.tree = true;
        if (!kernel) kernel = (p.n_lists < device_sms() && (d.W + sy - 1) / sy >= 64) ? LLAMPC_KERNEL_K1B
                                                                                     : (packed ? LLAMPC_KERNEL_K1P : LLAMPC_KERNEL_K1);
        if (kernel == LLAMPC_KERNEL_K1B) {
            const long long b = lookback_balanced_workspace_bytes(d.N, d.W);
            if (b < 0) return b <= -1000 ? (int)(-1000 - b) : (int)b;
            p.kernel = kernel;
            p.bytes = (size_t)b;
            p.grid_x = 0;                          // persistent grid: SMs x resident CTAs
            return 0;
        }
        if (kernel == LLAMPC_KERNEL_K1E) {
            size_t b = 0;
            const int rc = lookback_equal_plan(d.N, d.W, &p.grid_x, &p.block, &b, &p.lay);
            if (rc) return rc;
            p.kernel = kernel;
            p.split = 1;
            p.bytes = b;
            return 0;
        }
        if (kernel != LLAMPC_KERNEL_K1 && kernel != LLAMPC_KERNEL_K1P) return LLAMPC_E_ARG;
        p.kernel = kernel;
        p.lay = tree_layout(d.N, 0);
        p.bytes = p.lay.bytes;
        return 0;
    }
    if (d.peer_bufs) return LLAMPC_E_ARG;          // the exchange rides on the merge tree
    if (!kernel) kernel = packed ? LLAMPC_KERNEL_K1P : LLAMPC_KERNEL_K1;
    if (kernel != LLAMPC_KERNEL_K1 && kernel != LLAMPC_KERNEL_K1P) return LLAMPC_E_ARG;
    p.kernel = kernel;
    if (d.K > 0) {
        if (p.n_lists > MAX_MERGE_LISTS * MERGE_LPT) return LLAMPC_E_RANGE;
        p.in_kernel_merge = p.n_lists <= MAX_MERGE_LISTS;
        if (!p.in_kernel_merge) p.launches = 2;
        p.off_lists = up256((size_t)d.n_vehicles * sizeof(unsigned));
        p.bytes = p.off_lists + up256((size_t)d.n_vehicles * p.n_lists * LLAMPC_LIST_LEN * sizeof(u64));
    }
    return 0;
}

extern "C" int llampc_lookback_plan(const llampc_lookback_desc_t* desc, llampc_lookback_plan_t* plan) {
    if (!desc || !plan) return LLAMPC_E_ARG;
    LbPlan p;
    const int rc = resolve_plan(*desc, p);
    if (rc) return rc;
    plan->kernel = p.kernel; plan->split = p.split; plan->sine = p.sine;
    plan->grid_x = p.grid_x; plan->grid_y = p.grid_y; plan->block = p.block;
    plan->launches = p.launches;
    plan->workspace_bytes = p.bytes;
    return 0;
}

static int merge_lists_launch(const u64* cta_lists, int n_lists, int n_vehicles, int K, u64* out, cudaStream_t st) {
    const PeerXchg none = {nullptr, 0, 0, 0};
    if (n_lists <= 128 * MERGE_LPT)
        topk_merge_lists_kernel<128><<<n_vehicles, 128, 0, st>>>(cta_lists, n_lists, K, out, none);
    else if (n_lists <= 256 * MERGE_LPT)
        topk_merge_lists_kernel<256><<<n_vehicles, 256, 0, st>>>(cta_lists, n_lists, K, out, none);
    else
        topk_merge_lists_kernel<1024><<<n_vehicles, 1024, 0, st>>>(cta_lists, n_lists, K, out, none);
    return (int)cudaGetLastError();
}

static int lookback_launch_planned(const llampc_lookback_desc_t& d, const LbPlan& p, cudaStream_t st) {
    if (p.bytes > 0) {
        if (!d.workspace || d.workspace_bytes < p.bytes) return LLAMPC_E_ARG;
        if (!aligned16(d.workspace)) return LLAMPC_E_ALIGN;
    }
    unsigned char* wsb = static_cast<unsigned char*>(d.workspace);
    const bool geom = d.geom_shared != 0, mufu = p.sine == LLAMPC_SIN_SFU;
    const float4* bank = reinterpret_cast<const float4*>(d.bank);
    NewRow nr;
    nr.slot = -1;
    if (d.row32_h) {
        for (int i = 0; i < LLAMPC_HIST_ROW; ++i) nr.v[i] = d.row32_h[i];
        nr.slot = d.slot;
    }
    PeerXchg px = {nullptr, 0, 0, 0};
    if (d.peer_bufs) { px.peers = d.peer_bufs; px.world = d.world; px.rank = d.rank; px.seq = d.seq; }
    const StepSize z = make_step(d.Ts);
    if (p.kernel == LLAMPC_KERNEL_K1V)
        return launch_k1v(bank, d.N, d.Npad, d.W, z, d.slot, d.hist, d.n_vehicles, d.err_ring, d.avg_err, d.idx_offset,
                          d.emit, d.K, d.out, geom, mufu, st);
    if (p.kernel == LLAMPC_KERNEL_K1R) {
        nr.slot = d.slot;                          // the kernel reads the slot from here in both row modes
        unsigned* ticket = reinterpret_cast<unsigned*>(wsb);
        u64* lists = reinterpret_cast<u64*>(wsb + p.off_lists);
        const FusedMerge fm = {p.in_kernel_merge ? ticket : nullptr, p.in_kernel_merge ? d.out : nullptr,
                               p.in_kernel_merge ? d.K : 0};
        const int rc = launch_k1r(bank, d.N, d.Npad, d.W, z, nr, d.row32_h ? nullptr : d.hist, d.n_vehicles, d.err_ring,
                                  d.avg_err, lists, d.idx_offset, d.emit, fm, geom, mufu, st);
        if (rc || p.launches == 1) return rc;
        return merge_lists_launch(lists, p.n_lists, d.n_vehicles, d.K, d.out, st);
    }
    if (p.kernel == LLAMPC_KERNEL_K1E)
        return lookback_equal_launch(bank, d.N, d.Npad, d.hist, d.W, z, d.avg_err, d.idx_offset, geom, mufu, d.K, d.workspace,
                                     d.workspace_bytes, d.out, nr, px, st);
    if (p.kernel == LLAMPC_KERNEL_K1B)
        return lookback_balanced_launch(d.bank, d.N, d.Npad, d.hist, d.W, d.Ts, d.avg_err, d.idx_offset, geom, mufu, d.K,
                                        d.workspace, d.workspace_bytes, d.out, nr, px, st);
    LbArgs a;
    a.bank = bank; a.N = d.N; a.Npad = d.Npad;
    a.hist = d.hist; a.W = d.W; a.hist_stride_floats = (long)d.hist_stride_rows * LLAMPC_HIST_ROW; a.n_vehicles = d.n_vehicles;
    a.z = z; a.avg_err = d.avg_err; a.cta_lists = nullptr; a.idx_offset = d.idx_offset; a.nr = nr;
    a.fm = FusedMerge{nullptr, nullptr, 0};
    a.px = px;
    a.tm = TreeMerge{{nullptr, nullptr, nullptr, nullptr, nullptr}, nullptr, 0};
    const bool packed = p.kernel == LLAMPC_KERNEL_K1P;
    // PDL only where the kernel orders itself behind its predecessor (K1p with the merge tree), and not while a tick records
    // its launches for the graph.  A row patch riding in the launch is fine: every CTA of a preceding K1p launch has read the
    // ring (prologue) before it signals, and the re-score kernel does not read the fp32 ring.
    a.pdl = (d.flags & LLAMPC_LB_FLAG_PDL) && packed && p.tree && !launch_collector().slots;
    a.wide = (d.flags & LLAMPC_LB_FLAG_WIDE) && packed;
    if (p.tree) {
        a.tm = TreeMerge{tree_workspace(wsb, p.lay), d.out, d.K};
    } else if (d.K > 0) {
        a.cta_lists = reinterpret_cast<u64*>(wsb + p.off_lists);
        if (p.in_kernel_merge) a.fm = FusedMerge{reinterpret_cast<unsigned*>(wsb), d.out, d.K};
    }
    const int rc = packed ? launch_k1_packed(a, p.split, geom, mufu, st) : launch_k1_scalar(a, p.split, geom, mufu, st);
    if (rc || p.launches == 1) return rc;
    return merge_lists_launch(a.cta_lists, p.n_lists, d.n_vehicles, d.K, d.out, st);
}

extern "C" int llampc_lookback_launch(const llampc_lookback_desc_t* desc, llampc_stream_t stream) {
    if (!desc) return LLAMPC_E_ARG;
    LbPlan p;
    const int rc = resolve_plan(*desc, p);
    if (rc) return rc;
    return lookback_launch_planned(*desc, p, static_cast<cudaStream_t>(stream));
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

// ---------------------------------------------------------------------------------------------------
// The tick: workspace = [256 B of tick counters | look-back workspace | 17-key scratch | stand-alone top-K scratch]
// ---------------------------------------------------------------------------------------------------
constexpr size_t TICK_WS_HEAD = 256;               // word 0: re-score ticket, word 1: stand-alone top-K counter

struct TickLayout { llampc_lookback_desc_t d; LbPlan p; int Kt; bool fused; size_t off_lb, off_tmp, off_scratch, bytes; };

static int tick_layout(const llampc_tick_t* t, TickLayout& L) {
    if (!t || !t->bank || t->K < 0 || t->n_refine < 0) return LLAMPC_E_ARG;
    L.Kt = t->K > t->n_refine ? t->K : t->n_refine;
    if (L.Kt > LLAMPC_MAX_K) return LLAMPC_E_RANGE;
    L.fused = L.Kt <= LLAMPC_LIST_LEN;
    llampc_lookback_desc_t& d = L.d;
    memset(&d, 0, sizeof(d));
    d.bank = t->bank; d.N = t->N; d.Npad = t->Npad; d.idx_offset = t->idx_offset;
    d.geom_shared = t->geom_shared; d.sin_arg_max = t->sin_arg_max; d.sine = t->sine;
    d.hist = t->hist; d.W = t->W; d.n_vehicles = 1; d.hist_stride_rows = t->W; d.Ts = t->Ts;
    d.mode = t->rolling ? LLAMPC_LB_ROLLING : LLAMPC_LB_RECOMPUTE;
    d.slot = t->slot; d.emit = t->rolling == 2 ? 0 : 1;
    d.row32_h = t->row32_h; d.err_ring = t->err_ring;
    d.K = L.fused ? (L.Kt > 0 ? L.Kt : 1) : 1;
    d.avg_err = t->avg_err; d.out = t->result;
    d.kernel = t->kernel; d.split = t->split;
    if (t->rolling == 2) { d.K = 0; d.out = nullptr; }
    // a rolling tick always carries its row in the kernel parameters (checked by the tick); while only the workspace is
    // being sized the row pointer may not be set yet, and the plan must not depend on that
    static const float planning_row[LLAMPC_HIST_ROW] = {0.0f};
    if (t->rolling && !t->row32_h) d.row32_h = planning_row;
    const int rc = resolve_plan(d, L.p);
    if (rc) return rc;
    L.off_lb = TICK_WS_HEAD;
    L.off_tmp = L.off_lb + up256(L.p.bytes);
    L.off_scratch = L.off_tmp + up256((LLAMPC_LIST_LEN + 1) * sizeof(u64));
    size_t scratch = 0;
    if (!L.fused) scratch = (size_t)llampc_topk_scratch_ctas(t->N) * L.Kt * sizeof(u64);
    L.bytes = L.off_scratch + up256(scratch);
    return 0;
}

extern "C" long long llampc_lookback_tick_workspace_bytes(const llampc_tick_t* t) {
    TickLayout L;
    llampc_tick_t probe;
    if (!t) return LLAMPC_E_ARG;
    probe = *t;
    // the rolling layout is the larger of the two rolling phases and does not depend on the phase
    if (probe.rolling == 2) probe.rolling = 1;
    if (!probe.result) probe.result = reinterpret_cast<llampc_key_t*>(&probe);   // planning only: any non-NULL value
    if (probe.rolling && !probe.err_ring) return LLAMPC_E_ARG;
    if (probe.slot < 0 || probe.slot >= probe.W) probe.slot = 0;
    const int rc = tick_layout(&probe, L);
    if (rc) return rc > 0 ? -1000 - rc : rc;
    return (long long)L.bytes;
}

extern "C" int llampc_lookback_finish(llampc_tick_t* t, llampc_stream_t stream);

// Pipelined replays issue plain stream launches: a graph exec has one launch in flight at a time, and re-parameterising it
// while the previous tick is still running stalls the host (measured: 97 us per tick against 67 us for synchronous pushes).
static thread_local bool g_plain_launches = false;
// ... and put the fp64 re-score of a tick on a SIDE stream behind an event, so that the scoring kernels follow each other on the
// caller's stream with programmatic dependent launch (an event record between two launches does not break it: measured
// 36.98 against 37.03 us per tick) and the re-score of tick t runs beside the rows of tick t + 1.
struct SideLaunch { cudaStream_t stream; cudaEvent_t ev; };
static thread_local SideLaunch g_side = {nullptr, nullptr};

struct ReplayState {
    cudaStream_t side; cudaEvent_t ev[64]; cudaEvent_t done; llampc_key_t* dev_ring; size_t words; int depth; int device;
};

static void replay_state_destroy(ReplayState* rs) {
    if (!rs) return;
    if (rs->side) cudaStreamDestroy(rs->side);
    for (int j = 0; j < 64; ++j) if (rs->ev[j]) cudaEventDestroy(rs->ev[j]);
    if (rs->done) cudaEventDestroy(rs->done);
    if (rs->dev_ring) cudaFree(rs->dev_ring);
    memset(rs, 0, sizeof(*rs));
}

extern "C" int llampc_lookback_tick(llampc_tick_t* t, llampc_stream_t stream) {
    if (!t || !t->bank || !t->hist || !t->result || !t->result_h || !t->workspace) return LLAMPC_E_ARG;
    if (t->slot < 0 || t->slot >= t->W) return LLAMPC_E_ARG;
    if (!aligned16(t->workspace)) return LLAMPC_E_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TickLayout L;
    int rc = tick_layout(t, L);
    if (rc) return rc;
    if (t->workspace_bytes < L.bytes) return LLAMPC_E_ARG;
    if (t->rolling && (!t->err_ring || !t->row32_h || !L.fused)) return LLAMPC_E_ARG;
    const int Kt = L.Kt;
    unsigned char* wsb = static_cast<unsigned char*>(t->workspace);
    unsigned* counters = reinterpret_cast<unsigned*>(wsb);
    llampc_key_t* keys = t->result;                                  // [0] best, [1..Kt] finalists
    double* errs = reinterpret_cast<double*>(t->result + 1 + Kt);    // [Kt] fp64 scores
    L.d.workspace = wsb + L.off_lb;
    L.d.workspace_bytes = L.p.bytes;
    const bool refine = Kt > 0 && t->n_refine > 0 && t->rolling != 2;
    // ---- the two kernels of a tick (scoring + fp64 re-score) are replayed as one re-parameterised CUDA graph when the
    // tick is a one-launch scoring kernel followed by the re-score (LLAMPC_TICK_GRAPH=0: plain stream launches)
    PendingLaunch pending[TICK_GRAPH_MAX_NODES];
    LaunchCollector& lc = launch_collector();
    struct CollectGuard {                                            // never leave the collector armed on an error return
        LaunchCollector& lc; bool on;
        ~CollectGuard() { if (on) { lc.slots = nullptr; lc.n = 0; } }
    } guard = {lc, false};
    if (lib_env().tick_graph && !g_plain_launches && refine && L.fused && L.p.launches == 1) {
        if (!t->graph_state) t->graph_state = calloc(1, sizeof(TickGraph));
        if (t->graph_state) {
            lc.slots = pending;
            lc.n = 0;
            guard.on = true;
        }
    }
    if (t->hard_h && t->row32_h && !t->rolling) {                    // low-speed / drift rows in the window -> the wide form
        const float vx = fabsf(t->row32_h[6]), vy = fabsf(t->row32_h[7]), w = fabsf(t->row32_h[8]);
        const unsigned char hard = (vx < 0.6f || vy + 0.06f * w > 0.4f * vx) ? 1 : 0;
        t->n_hard += (int)hard - (int)t->hard_h[t->slot];
        t->hard_h[t->slot] = hard;
    }
    if (t->hard_h && !t->rolling && t->n_hard * 10 >= t->W) L.d.flags |= LLAMPC_LB_FLAG_WIDE;
    if (L.fused) {
        // K1 / K1p / K1b with the merge tree (or K1r with the last-CTA merge): writes keys[0..LIST_LEN] itself
        if (g_plain_launches) L.d.flags |= LLAMPC_LB_FLAG_PDL;       // pipelined replay: overlap with the previous tick's re-score
        rc = lookback_launch_planned(L.d, L.p, st);
        if (rc) return rc;
        if (t->rolling == 2) {                                       // window still filling: column stored, no decision
            if (t->n_refine > 0 && t->row64_h && t->hist64)
                LLAMPC_CUDA_TRY(cudaMemcpyAsync(t->hist64 + (size_t)t->slot * LLAMPC_HIST64_ROW, t->row64_h,
                                                LLAMPC_HIST64_ROW * sizeof(double), cudaMemcpyHostToDevice, st));
            return 0;
        }
    } else {
        // more finalists than a CTA list holds: scores + arg-min from the launch, then the stand-alone top-K over avg_err
        if (!t->avg_err) return LLAMPC_E_ARG;
        llampc_key_t* tmp = reinterpret_cast<llampc_key_t*>(wsb + L.off_tmp);
        L.d.out = tmp;
        rc = lookback_launch_planned(L.d, L.p, st);
        if (rc) return rc;
        LLAMPC_CUDA_TRY(cudaMemcpyAsync(keys, tmp, sizeof(llampc_key_t), cudaMemcpyDeviceToDevice, st));
        rc = llampc_topk_f32(t->avg_err, t->N, t->idx_offset, Kt, reinterpret_cast<llampc_key_t*>(wsb + L.off_scratch),
                             counters + 1, keys + 1, stream);
        if (rc) return rc;
    }
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
        // zero-copy hand-off: the last re-score block writes the result into mapped pinned memory
        const bool gather = t->peer_world > 1 && t->peer_bufs != nullptr;
        const int words = gather ? 1 + 2 * Kt * t->peer_world : 1 + 2 * Kt;
        FinalCopy fc = {nullptr, nullptr, nullptr, 0, 0};
        PeerGather pg = {nullptr, 0, 0, 0, 0};
        if (gather) {
            if (!t->zero_copy) return LLAMPC_E_ARG;                  // the gather rides on the zero-copy hand-off
            pg.peers = t->peer_bufs; pg.world = t->peer_world; pg.rank = t->peer_rank; pg.seq = t->peer_seq; pg.kt = Kt;
        }
        if (t->zero_copy) {
            void* dptr = t->mapped_for == t->result_h ? t->mapped_dev : nullptr;
            if (!dptr && cudaHostGetDevicePointer(&dptr, t->result_h, 0) == cudaSuccess && dptr) {
                t->mapped_dev = dptr;
                t->mapped_for = t->result_h;
            }
            if (dptr) {
                static unsigned long long seq_counter = 1;
                fc.src = t->result;
                fc.dst_host = static_cast<volatile u64*>(dptr);
                fc.ticket = counters;
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
        if (g_side.stream) {                                         // pipelined replay: the re-score goes to the side stream
            LLAMPC_CUDA_TRY(cudaEventRecord(g_side.ev, st));
            LLAMPC_CUDA_TRY(cudaStreamWaitEvent(g_side.stream, g_side.ev, 0));
            rc = issue(refine_f64_kernel, dim3(Kt), dim3(RF_THREADS), 0, g_side.stream, t->bank64, t->N, t->hist64, t->W, t->Ts,
                       static_cast<const u64*>(keys + 1), t->idx_offset, errs, nr64, fc, pg);
        } else if (g_plain_launches && (L.d.flags & LLAMPC_LB_FLAG_PDL) && L.p.kernel == LLAMPC_KERNEL_K1P && L.p.tree)
            rc = issue_pdl(refine_f64_kernel, dim3(Kt), dim3(RF_THREADS), 0, st, t->bank64, t->N, t->hist64, t->W, t->Ts,
                           static_cast<const u64*>(keys + 1), t->idx_offset, errs, nr64, fc, pg);
        else
            rc = issue(refine_f64_kernel, dim3(Kt), dim3(RF_THREADS), 0, st, t->bank64, t->N, t->hist64, t->W, t->Ts,
                       static_cast<const u64*>(keys + 1), t->idx_offset, errs, nr64, fc, pg);
        if (rc) return rc;
    }
    if (guard.on) {                                                  // replay what was collected as one graph
        const int n = lc.n;
        lc.slots = nullptr;
        lc.n = 0;
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

// T ticks of a recorded run, pipelined (see the header): one C loop instead of T FFI crossings, up to `depth` ticks in flight.
extern "C" int llampc_lookback_replay(llampc_tick_t* t, const double* x_k, const double* u_k, const double* x_k1, int x1_stride,
                                      int T, int first_slot, double lf_shared, double lr_shared, float* rows32_h,
                                      double* rows64_h, llampc_key_t* const* slots_h, int depth, unsigned* peer_seq,
                                      long long* idx_out, double* score_out, int* n_valid, llampc_stream_t stream) {
    if (!t || !x_k || !u_k || !x_k1 || !rows32_h || !rows64_h || !slots_h || !idx_out || !score_out || !n_valid) return LLAMPC_E_ARG;
    if (T < 0 || x1_stride < 4 || depth < 1 || depth > 64 || first_slot < 0 || first_slot >= t->W) return LLAMPC_E_RANGE;
    if (!t->zero_copy || t->n_refine <= 0 || t->rolling > 1) return LLAMPC_E_ARG;
    const int kt = t->K > t->n_refine ? t->K : t->n_refine;
    struct InFlight { llampc_key_t* rh; unsigned long long seq; int words; int tick; } ring[64];
    void* dev_of[64];
    for (int j = 0; j < depth; ++j) {
        if (!slots_h[j]) return LLAMPC_E_ARG;
        LLAMPC_CUDA_TRY(cudaHostGetDevicePointer(&dev_of[j], slots_h[j], 0));
    }
    // side stream + one event and one DEVICE result buffer per slot (recompute mode on the packed kernel only: that is where the
    // scoring kernels can follow each other with programmatic dependent launch)
    ReplayState* rs = nullptr;
    if (!t->rolling && depth > 1) {
        int dev = 0;
        LLAMPC_CUDA_TRY(cudaGetDevice(&dev));
        const size_t words = (size_t)((1 + 2 * kt) > (LLAMPC_LIST_LEN + 1) ? (1 + 2 * kt) : (LLAMPC_LIST_LEN + 1)) + 1;
        rs = static_cast<ReplayState*>(t->replay_state);
        if (rs && (rs->device != dev || rs->depth < depth || rs->words < words)) { replay_state_destroy(rs); free(rs); rs = nullptr; t->replay_state = nullptr; }
        if (!rs) {
            rs = static_cast<ReplayState*>(calloc(1, sizeof(ReplayState)));
            if (!rs) return LLAMPC_E_ARG;
            t->replay_state = rs;
            rs->device = dev; rs->depth = depth; rs->words = words;
            bool ok = cudaStreamCreateWithFlags(&rs->side, cudaStreamNonBlocking) == cudaSuccess;
            for (int j = 0; j < depth && ok; ++j) ok = cudaEventCreateWithFlags(&rs->ev[j], cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&rs->done, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaMalloc(&rs->dev_ring, (size_t)depth * words * sizeof(llampc_key_t)) == cudaSuccess;
            ok = ok && cudaMemset(rs->dev_ring, 0, (size_t)depth * words * sizeof(llampc_key_t)) == cudaSuccess;
            if (!ok) { (void)cudaGetLastError(); replay_state_destroy(rs); free(rs); rs = nullptr; t->replay_state = nullptr; }
        }
    }
    const llampc_tick_t saved = *t;
    int head = 0, count = 0, rc = 0;
    struct PlainGuard {
        PlainGuard() { g_plain_launches = true; }
        ~PlainGuard() { g_plain_launches = false; g_side.stream = nullptr; g_side.ev = nullptr; }
    } plain_guard;
    auto finish_oldest = [&]() -> int {
        const InFlight& f = ring[head];
        t->result_h = f.rh; t->pending_seq = f.seq; t->pending_words = f.words;
        int r = llampc_lookback_finish(t, stream);
        if (!r) r = llampc_lookback_decode(t, idx_out + (size_t)f.tick * kt, score_out + (size_t)f.tick * kt, n_valid + f.tick);
        head = (head + 1) % depth;
        --count;
        return r;
    };
    const bool trace = getenv("LLAMPC_REPLAY_TRACE") != nullptr;
    double t_fin = 0, t_pack = 0, t_tick = 0;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (int i = 0; i < T && !rc; ++i) {
        const double a0 = trace ? now() : 0;
        if (count == depth) { rc = finish_oldest(); if (rc) break; }
        const double a1 = trace ? now() : 0;
        const int slot = (first_slot + i) % t->W, j = i % depth;
        float* r32 = rows32_h + (size_t)slot * LLAMPC_HIST_ROW;
        double* r64 = rows64_h + (size_t)slot * LLAMPC_HIST64_ROW;
        rc = llampc_hist_row_pack_h(x_k + (size_t)i * 6, u_k + (size_t)i * 2, x_k1 + (size_t)i * x1_stride, t->Ts, lf_shared,
                                    lr_shared, r32, r64);
        if (rc) break;
        const double a2 = trace ? now() : 0;
        t->row32_h = r32; t->row64_h = r64; t->slot = slot;
        t->result_h = slots_h[j]; t->mapped_for = slots_h[j]; t->mapped_dev = dev_of[j];
        t->sync = 0;
        if (peer_seq) { *peer_seq = *peer_seq % 0xFFFFFFFFu + 1u; t->peer_seq = *peer_seq; }
        if (rs) {                                                    // this tick's own device result buffer and event
            t->result = rs->dev_ring + (size_t)j * rs->words;
            g_side.stream = rs->side;
            g_side.ev = rs->ev[j];
        }
        rc = llampc_lookback_tick(t, stream);
        if (rc) break;
        ring[(head + count) % depth] = InFlight{slots_h[j], t->pending_seq, t->pending_words, i};
        ++count;
        if (trace) { const double a3 = now(); t_fin += a1 - a0; t_pack += a2 - a1; t_tick += a3 - a2; }
    }
    if (trace && T > 0)
        fprintf(stderr, "replay T=%d depth=%d: per tick  wait+decode %.1f us  pack %.1f us  enqueue %.1f us  (pending_seq %llu)\n", T, depth,
                t_fin / T, t_pack / T, t_tick / T, (unsigned long long)t->pending_seq);
    while (count > 0) {                                              // drain (also after an error: the slots must go quiet)
        const int r = finish_oldest();
        if (!rc) rc = r;
    }
    if (rs) {                                                        // later work on the caller's stream sees the re-scores done
        if (cudaEventRecord(rs->done, rs->side) == cudaSuccess) (void)cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), rs->done, 0);
        else (void)cudaGetLastError();
    }
    void* gs = t->graph_state;                                       // the graph may have been (re)built meanwhile
    const int n_hard = t->n_hard;                                    // ... and the ring's low-speed flags were updated
    void* rstate = t->replay_state;
    *t = saved;
    t->graph_state = gs;
    t->n_hard = n_hard;
    t->replay_state = rstate;
    t->pending_seq = 0; t->pending_words = 0;
    return rc;
}

// Frees what llampc_lookback_tick attached to the struct (the CUDA graph of the tick).  Safe to call more than once.
extern "C" int llampc_lookback_tick_release(llampc_tick_t* t) {
    if (!t) return LLAMPC_E_ARG;
    if (t->graph_state) {
        tick_graph_destroy(static_cast<TickGraph*>(t->graph_state));
        free(t->graph_state);
        t->graph_state = nullptr;
    }
    if (t->replay_state) {
        replay_state_destroy(static_cast<ReplayState*>(t->replay_state));
        free(t->replay_state);
        t->replay_state = nullptr;
    }
    return 0;
}

// layout probes for FFI bindings that mirror llampc_tick_t by hand
extern "C" int llampc_tick_sizeof(void) { return (int)sizeof(llampc_tick_t); }
extern "C" int llampc_tick_offsetof(int which) {
    switch (which) {
        case 0: return (int)offsetof(llampc_tick_t, Ts);
        case 1: return (int)offsetof(llampc_tick_t, avg_err);
        case 2: return (int)offsetof(llampc_tick_t, result_h);
        case 3: return (int)offsetof(llampc_tick_t, peer_seq);
        case 4: return (int)offsetof(llampc_tick_t, rolling);
        case 5: return (int)offsetof(llampc_tick_t, workspace);
        default: return -1;
    }
}
extern "C" int llampc_lookback_desc_sizeof(void) { return (int)sizeof(llampc_lookback_desc_t); }
