// Selection / exchange pieces shared by the look-back kernels (K1, K1r, K1b): CTA-level key selection, the history
// row that rides in the kernel parameters, the in-kernel list merge descriptor and the NVLink peer min-loc.
#pragma once
#include "llampc_common.cuh"

namespace llampc {

constexpr int LB_THREADS = 128;

#ifndef LLAMPC_LB_MIN_BLOCKS
#define LLAMPC_LB_MIN_BLOCKS 6
#endif

// CTA-level selection: every key-holding warp (the first KW warps) sorts its 32 keys (registers + shuffles), sorted
// runs are merged pairwise through shared memory; warp 0 ends up with the CTA's 32 smallest keys in ascending lane
// order (the return value is meaningful in warp 0 only).  skeys: LB_THREADS keys of shared memory.
template <int KW, bool PRESORTED = false>
__device__ __forceinline__ u64 cta_select32(u64 key, u64* skeys) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!PRESORTED && warp < KW) key = warp_sort_u64(key, lane);
    if (KW == 4) {
        if (warp == 1 || warp == 3) skeys[warp * 32 + lane] = key;
        __syncthreads();
        if (warp == 0 || warp == 2) key = warp_merge_low32(key, skeys[(warp + 1) * 32 + 31 - lane], lane);
        __syncthreads();
        if (warp == 2) skeys[lane] = key;
        __syncthreads();
        if (warp == 0) key = warp_merge_low32(key, skeys[31 - lane], lane);
    } else if (KW == 2) {
        if (warp == 1) skeys[lane] = key;
        __syncthreads();
        if (warp == 0) key = warp_merge_low32(key, skeys[31 - lane], lane);
    }
    return key;
}

// K1 / K1r grid path: lanes 0..15 of warp 0 write this CTA's ascending list for the per-vehicle top-K merge.
template <int KW, bool PRESORTED = false>
__device__ __forceinline__ void cta_select_emit(u64 key, u64* skeys, int v, u64* __restrict__ cta_lists) {
    key = cta_select32<KW, PRESORTED>(key, skeys);
    if ((threadIdx.x >> 5) == 0) {
        const int lane = threadIdx.x & 31;
        if (cta_lists && lane < LLAMPC_LIST_LEN)
            cta_lists[((size_t)v * gridDim.x + blockIdx.x) * LLAMPC_LIST_LEN + lane] = key;
    }
}

// The newest history row can travel with the launch as a kernel parameter (80 bytes) instead of a separate
// H2D copy: every CTA patches its shared-memory copy of ring slot `slot`, CTA 0 also stores it to the ring.
struct NewRow { float v[LLAMPC_HIST_ROW]; int slot; };

// Optional in-kernel finish of the top-K: the last CTA of a vehicle to retire (atomic ticket) merges the per-CTA
// lists itself, so a tick is ONE launch.  Needs gridDim.x <= 128 * MERGE_LPT lists; K = 0 disables it.
struct FusedMerge { unsigned* ticket; u64* out; int K; };

// Optional multi-GPU min-loc fused into the same launch, over NVLink peer memory (no NCCL call on the path): every
// rank owns a small symmetric buffer [2 parities][world][2] of u64; the root warp of rank r stores its packed arg-min key
// into slot r of EVERY peer's buffer, then spins on its own buffer until all `world` slots carry the current sequence
// number and reduces them.  The two words of a slot are self-validating ("LL" style): word 0 = (low half of the key,
// sequence), word 1 = (high half of the key, sequence), each an atomic 8-byte remote store -- no system-scope fence between
// data and flag, so the exchange costs ONE NVLink traversal instead of a fence round trip plus a traversal.
// Double-buffered by the parity of `seq`, which the host increments identically on every rank each tick (never 0: the
// buffers start zeroed).  world = 0 disables it.
struct PeerXchg { u64* const* peers; int world; int rank; unsigned seq; };

// A peer that never arrives (~1 s) POISONS the result: every lane then returns ~0ull -- "no candidate", which no valid
// key equals -- instead of a plausible wrong winner; the host side turns it into LLAMPC_E_PEER.
__device__ __forceinline__ u64 peer_minloc(const PeerXchg& px, u64 my_key, int lane) {
    const int parity = px.seq & 1;
    const u64 tag = (u64)px.seq << 32;
    u64 got = ~0ull;
    bool ok = true;
    if (lane < px.world) {
        volatile u64* dst = px.peers[lane] + ((size_t)parity * px.world + px.rank) * 2;
        dst[0] = tag | (my_key & 0xffffffffull);
        dst[1] = tag | (my_key >> 32);
        volatile u64* src = px.peers[px.rank] + ((size_t)parity * px.world + lane) * 2;
        const long long t0 = clock64();
        u64 w0, w1;
        for (;;) {
            w0 = src[0];
            w1 = src[1];
            if ((w0 >> 32) == (u64)px.seq && (w1 >> 32) == (u64)px.seq) break;
            if (clock64() - t0 > 2000000000ll) { ok = false; break; }     // ~1 s: a peer never arrived
        }
        if (ok) got = (w1 << 32) | (w0 & 0xffffffffull);
    }
    if (!__all_sync(0xffffffffu, ok)) return ~0ull;
    return warp_min_key(got);
}

// ---------------------------------------------------------------------------------------------------
// Tree merge of per-CTA top-16 lists inside the producing launch (K1 tree mode and K1b), and the workspace it lives in.
// ---------------------------------------------------------------------------------------------------
// release / acquire fence at GPU scope (MEMBAR.ALL.GPU): lighter than __threadfence()'s sequentially consistent fence
// and sufficient for the store -> count / count -> load message passing of the merge tree
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned atom_add_acq_rel_gpu(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

constexpr int BAL_FAN = 32;                        // lists merged per tree node (one per lane)
constexpr int BAL_ROW_PAD = LLAMPC_LIST_LEN + 1;   // shared-memory row pitch of the merge staging (bank spread)

struct BalWs {
    float* part;        // K1b only: [n_tasks][32] partial window sums, one row of 32 per task (when SY > 1)
    unsigned* next;     // K1b only: task counter (zero between launches)
    unsigned* gcount;   // K1b only: [ceil(N/32)] arrivals per candidate warp-group (zero between launches)
    unsigned* mcount;   // [sum over levels of ceil(n_level / 32)] arrivals per tree node  (zero between launches)
    u64* lists;         // [sum over levels of n_level][16]  level 0 = one list per CTA
};

// One warp: this list (lanes 0..15 ascending) is list `idx` of `n` at the current tree level.  Climbs the tree while
// this warp is the last arrival of its node; the warp that produces the root writes `out` (and runs the NVLink
// min-loc when several GPUs share the bank).
__device__ __forceinline__ void tree_merge(u64 key, int lane, int idx, int n, int K, const BalWs& ws,
                                               u64 (*mrows)[BAL_ROW_PAD], u64* __restrict__ out, const PeerXchg& px) {
    u64* lists = ws.lists;
    unsigned* cnt = ws.mcount;
    const int rounds = K > 0 ? K : 1;              // every level keeps the K smallest keys (the rest padded with ~0)
    while (n > 1) {
        if (lane < LLAMPC_LIST_LEN) __stcg(lists + (size_t)idx * LLAMPC_LIST_LEN + lane, key);
        __syncwarp();
        const int node = idx / BAL_FAN;
        const int c = min(BAL_FAN, n - node * BAL_FAN);
        unsigned old = 0;
        // ONE acq_rel atomic instead of fence + atomic + fence: its release half (cumulative over the warp barrier above)
        // publishes this warp's list before the count, its acquire half orders the loads of the other lists -- which the
        // other lanes issue after the shuffle below, a warp-level synchronisation -- after the count
        if (lane == 0) old = atom_add_acq_rel_gpu(cnt + node, 1u);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != (unsigned)(c - 1)) return;      // a later arrival merges this node
        if (lane == 0) cnt[node] = 0;              // ready for the next launch on the same stream
        const u64* src = lists + (size_t)node * BAL_FAN * LLAMPC_LIST_LEN;
#pragma unroll
        for (int j = 0; j < LLAMPC_LIST_LEN; ++j) {            // c lists are contiguous: coalesced, one L2 round trip
            const int i = lane + 32 * j;
            mrows[i / LLAMPC_LIST_LEN][i % LLAMPC_LIST_LEN] = i < c * LLAMPC_LIST_LEN ? __ldcg(src + i) : ~0ull;
        }
        __syncwarp();
        int pos = 0;
        u64 h = mrows[lane][0], mine = ~0ull;
#pragma unroll 1
        for (int r = 0; r < rounds; ++r) {                     // 32-way merge: lane = list, head in a register
            const u64 sel = warp_min_key(h);
            if (lane == r) mine = sel;
            if (h == sel && sel != ~0ull) {
                ++pos;
                h = pos < LLAMPC_LIST_LEN ? mrows[lane][pos] : ~0ull;
            }
        }
        __syncwarp();
        key = mine;
        lists += (size_t)n * LLAMPC_LIST_LEN;
        cnt += (n + BAL_FAN - 1) / BAL_FAN;
        idx = node;
        n = (n + BAL_FAN - 1) / BAL_FAN;
    }
    // root: out[0] = arg-min key, out[1..K] = ascending top-K, the rest of the LIST_LEN slots padded
    u64 best = __shfl_sync(0xffffffffu, key, 0);
    if (px.world > 1) best = peer_minloc(px, best, lane);
    if (lane == 0) {
        out[0] = best;
        *ws.next = 0;                              // every CTA has left its task loop: re-arm the counter for the next launch
    }
    if (lane < LLAMPC_LIST_LEN) out[1 + lane] = lane < K ? key : ~0ull;
}

constexpr int TREE_MAX_CTAS_PER_SM = 8;

static inline int device_sms() {
    static int sms[64] = {0};                      // per device (one process may drive several GPUs)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { (void)cudaGetLastError(); return 148; }
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
            (void)cudaGetLastError();
            v = 148;                               // B200
        }
        sms[dev] = v;
    }
    return sms[dev];
}

// Workspace layout.  Counters come first at offsets that depend on N only, so a workspace zeroed once stays valid when
// W, the window split or the K1b schedule change; the level-0 list count is sized for the larger of the finest K1
// tiling (16 window splits: ceil(N / 8) CTAs) and the largest persistent K1b grid.
struct TreeLayout { size_t off_cnt, off_gcount, off_lists, off_part, bytes; };

static inline TreeLayout tree_layout(int N, size_t part_bytes) {
    TreeLayout l;
    const long long max_k1 = ((long long)N + 7) / 8;
    const long long max_k1b = (long long)device_sms() * TREE_MAX_CTAS_PER_SM;
    const long long n0 = max_k1 > max_k1b ? max_k1 : max_k1b;
    size_t n_lists = 0, n_nodes = 0;
    for (long long n = n0;; n = (n + BAL_FAN - 1) / BAL_FAN) {
        n_lists += (size_t)n;
        if (n == 1) break;
        n_nodes += (size_t)((n + BAL_FAN - 1) / BAL_FAN);
    }
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    l.off_cnt = 0;                                               // task counter, then the tree-node counters
    l.off_gcount = up((n_nodes + 4) * sizeof(unsigned));
    l.off_lists = l.off_gcount + up((size_t)((N + 31) / 32) * sizeof(unsigned));
    l.off_part = l.off_lists + up(n_lists * LLAMPC_LIST_LEN * sizeof(u64));
    l.bytes = l.off_part + up(part_bytes);
    return l;
}

static inline BalWs tree_workspace(unsigned char* wsb, const TreeLayout& l) {
    BalWs ws;
    ws.next = reinterpret_cast<unsigned*>(wsb + l.off_cnt);
    ws.mcount = ws.next + 4;
    ws.gcount = reinterpret_cast<unsigned*>(wsb + l.off_gcount);
    ws.lists = reinterpret_cast<u64*>(wsb + l.off_lists);
    ws.part = reinterpret_cast<float*>(wsb + l.off_part);
    return ws;
}

// Optional tree finish of K1 (single history): warp 0 of every CTA enters tree_merge with the CTA's list.  K = 0 disables it.
struct TreeMerge { BalWs ws; u64* out; int K; };

// K1b (lookback_balanced.cu): internal launcher shared with the C ABI in lookback.cu.
long long lookback_balanced_workspace_bytes(int N, int W);
int lookback_balanced_launch(const float* bank, int N, int Npad, const float* hist, int W, double Ts, float* avg_err,
                             int idx_offset, int geom_shared, int mufu_sin, int K, void* workspace,
                             unsigned long long workspace_bytes, u64* out, const NewRow& nr, const PeerXchg& px,
                             cudaStream_t st);

}  // namespace llampc
