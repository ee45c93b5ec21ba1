// Launch plumbing shared by the look-back translation units: the launch recorder behind the one-graph tick, the
// library-wide tunables (environment read ONCE, at first use) and per-device caches.
#pragma once
#include "llampc_common.cuh"
#include <string.h>

namespace llampc {

// ---------------------------------------------------------------------------------------------------
// Launch helper.  Normally issue() is kern<<<...>>>(...).  While llampc_lookback_tick collects (collector armed), the
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

struct LaunchCollector { PendingLaunch* slots; int n; };
LaunchCollector& launch_collector();               // thread-local; defined in lookback.cu

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
    LaunchCollector& lc = launch_collector();
    if (lc.slots && lc.n < TICK_GRAPH_MAX_NODES) {
        PendingLaunch& pl = lc.slots[lc.n++];
        pl.func = reinterpret_cast<void*>(kern);
        pl.grid = grid; pl.block = block; pl.smem = smem; pl.n_args = 0; pl.used = 0;
        (pending_push<KArgs>(pl, static_cast<KArgs>(args)), ...);
        return 0;
    }
    kern<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
    return (int)cudaGetLastError();
}

// Launch with the programmatic-stream-serialization attribute: the grid may start while the previous grid on the stream is
// still running, once every CTA of that grid has executed griddepcontrol.launch_dependents (or exited); the kernel itself
// orders its first conflicting access behind the previous grid with griddepcontrol.wait.
template <class... KArgs, class... Args>
static int issue_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static_assert(sizeof...(KArgs) == sizeof...(Args), "argument count");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Experiment switches.  The environment is read once per process (first use); everything a caller may legitimately
// want to steer per call is a field of llampc_lookback_desc_t instead (kernel, split, sine).
struct LibEnv {
    int tick_graph;          // LLAMPC_TICK_GRAPH=0: plain stream launches instead of the re-parameterised graph
    long long bal_ctas;      // LLAMPC_BAL_CTAS: K1b CTAs per SM x 100 (0 = occupancy)
    long long bal_tpw;       // LLAMPC_BAL_TPW: K1b target tasks per resident warp
    long long bal_rmin, bal_rmax;   // LLAMPC_BAL_RMIN / RMAX: K1b rows per task bounds
    int bal_verbose, bal_trace, bal_nofence;
    long long eq_ctas;       // LLAMPC_EQ_CTAS: K1e CTAs per SM x 100 (0 = occupancy)
    long long eq_stagger;    // LLAMPC_EQ_STAGGER: K1e start offset in clocks between the warps of a scheduler
};
const LibEnv& lib_env();                           // defined in lookback.cu

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute and setting it is cheap and idempotent: the
// launchers call this before every launch that needs more than 48 KB (no process-wide "raised" flag).
template <class F>
static inline int raise_dynamic_smem(F kern, size_t bytes) {
    if (bytes <= 48 * 1024) return 0;
    return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace llampc
