/* llampc_b200.h -- C ABI of the B200-native LLA-MPC look-back / look-ahead hot path.
 *
 * The reference (tianhao-stan-wu/LLA-MPC) is 100 % Python and has no FFI of its own; the boundary it
 * exposes for this path is a set of Python call signatures.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference root).  The Python host in
 * llampc_b200/ binds these symbols with ctypes (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - plain pointers and sizes only; pointers are DEVICE pointers unless the name ends in _h;
 *   - the caller owns every buffer, nothing is allocated or freed inside the library;
 *   - every device entry point is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = OK, > 0 = cudaError_t of the failing runtime call, < 0 = argument error
 *     (LLAMPC_E_*); the library never throws and keeps no global state.
 */
#ifndef LLAMPC_B200_H
#define LLAMPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LLAMPC_ABI_VERSION 6

#define LLAMPC_E_ARG   (-1)  /* null pointer / non-positive size / size not supported       */
#define LLAMPC_E_ALIGN (-2)  /* pointer or stride not 16-byte aligned                        */
#define LLAMPC_E_RANGE (-3)  /* W, K, H ... outside the compiled limits                      */
#define LLAMPC_E_PEER  (-4)  /* multi-GPU exchange: a peer did not deliver within ~1 s; the tick has NO decision */

#define LLAMPC_NPARAM        14   /* lf lr mass Iz Bf Br Cf Cr Df Dr Cm1 Cm2 Cr0 Cr2 (Dynamic.__init__, llampc/models/dynamic.py:24-57) */
#define LLAMPC_BANK_GROUPS    4   /* packed bank = 4 float4 groups per candidate                */
#define LLAMPC_HIST_ROW      20   /* floats per history row (80 B)                              */
#define LLAMPC_HIST64_ROW    12   /* doubles per f64 history row: x_k[6] u_k[2] x_k1[0:4]       */
#define LLAMPC_MAX_W       1024   /* history rows staged in shared memory per CTA               */
#define LLAMPC_MAX_K         64   /* top-K (llampc_topk_f32)                                    */
#define LLAMPC_LIST_LEN      16   /* keys per CTA list written by K1; top-K of the fused path   */
#define LLAMPC_MAX_H        256   /* look-ahead horizon                                         */

typedef void* llampc_stream_t;           /* cudaStream_t */
typedef unsigned long long llampc_key_t; /* (float_bits(avg_err) << 32) | global candidate index */

int llampc_abi_version(void);
const char* llampc_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * Host-side packing (pure C, no CUDA): fp64 reference-layout data -> the fp32 device layouts.
 * ------------------------------------------------------------------------------------------- */

/* Model bank (replaces the gather of six (N,) arrays, run_nmpc_orca_llampc_rt.py:172-179, and the
 * array-parameter Dynamic built in llampc/mpc/evaluate_models_vectorized.py:13-22).
 * params_h[j] points to n_j doubles, n_j = N if is_array[j] else 1, in LLAMPC_NPARAM order.
 * packed_h receives 4 groups of Npad float4 (group-major: [g][i] at float offset (g*Npad+i)*4):
 *   g0 = Bf Cf Df Br | g1 = Cr Dr 1/mass lf | g2 = lr lf/Iz lr/Iz Cm1 | g3 = Cm2 Cr0 Cr2 0
 * derived quantities are formed in fp64 and rounded once.  Npad >= N; rows N..Npad-1 repeat row N-1. */
int llampc_bank_pack_h(const double* const* params_h, const int* is_array, int N, int Npad, float* packed_h,
                       float* sin_arg_max_h /* or NULL: receives max(|Cf|, |Cr|) pi/2 over the bank, the bound of the tyre-sine
                                               argument C atan(B alpha) that LLAMPC_SIN_AUTO compares with pi */);

/* One history row from one measured transition (x_k, u_k) -> x_k1 (the tick body of
 * run_nmpc_orca_llampc_rt.py:347-349 needs exactly these three vectors).  All trigonometry and the
 * measured increments are formed in fp64 here so the fp32 kernel never subtracts O(1) numbers.
 * lf_shared/lr_shared: the bank-wide lf, lr when geometry is not varied (else NaN): enables the
 * candidate-invariant stage-1 slip angles.  row32_h: LLAMPC_HIST_ROW floats, row64_h (may be NULL):
 * LLAMPC_HIST64_ROW doubles for llampc_refine_f64. */
int llampc_hist_row_pack_h(const double* x_k, const double* u_k, const double* x_k1, double Ts,
                           double lf_shared, double lr_shared, float* row32_h, double* row64_h);

/* ---------------------------------------------------------------------------------------------
 * LOOK-BACK.  Three entry points:
 *   llampc_lookback_launch   the scoring + selection launch on device-resident inputs (any layout / mode / GPU count)
 *   llampc_lookback_tick     one MPC tick around it: newest row in, launch, fp64 re-score, result to the host
 *   llampc_lookback_push     the tick from three fp64 host vectors (one FFI crossing per MPC tick)
 *   llampc_lookback_replay   T ticks of a recorded run from host arrays, pipelined (one FFI crossing per run)
 * (+ the helpers llampc_lookback_plan, ..._finish, ..._decode, ..._tick_release, ..._tick_workspace_bytes).
 *
 * For every candidate: W one-step RK4 predictions re-anchored at the measured states, mean squared error over
 * (x, y, psi, vx) and over the window, arg-min and top-K.  Replaces evaluate_models_vectorized
 * (llampc/mpc/evaluate_models_vectorized.py:4-23) called once per tick + errors / error_windows / mean / argmin /
 * argsort()[:K] of run_nmpc_orca_llampc_rt.py:349-360.
 * ------------------------------------------------------------------------------------------- */
#define LLAMPC_LB_RECOMPUTE 0   /* re-integrate the whole W-row window every call (N*W steps, stateless w.r.t. the bank)  */
#define LLAMPC_LB_ROLLING   1   /* the reference's own bookkeeping (rt.py:352-354): integrate only the newest row, replace
                                   column `slot` of err_ring, re-sum the ring (N steps per call)                           */
/* rows of an error ring for a window of W: W error columns + ceil(W / 4) rows of partial sums (groups of 4 columns; the
 * window sum is formed group by group in a fixed order, identical in every rolling kernel).  Zero it before the first call. */
#define LLAMPC_RING_ROWS(W) ((W) + (((W) + 3) >> 2))

#define LLAMPC_SIN_AUTO     0   /* MUFU.SIN while sin_arg_max <= pi (its rated range), else the polynomial; 0 / unknown
                                   sin_arg_max selects the polynomial                                                      */
#define LLAMPC_SIN_SFU      1   /* tyre sine on the SFU (MUFU.SIN): abs. error 2^-21.4 on [-pi, pi], growing beyond        */
#define LLAMPC_SIN_STRICT   2   /* Cody-Waite reduction + polynomial on the FMA pipe: ~17 % slower, range-independent      */

#define LLAMPC_KERNEL_AUTO  0   /* chosen from the shape (llampc_lookback_plan reports the choice)                         */
#define LLAMPC_KERNEL_K1    1   /* window kernel, one candidate per thread                                                  */
#define LLAMPC_KERNEL_K1P   2   /* window kernel, two candidates per thread in packed f32x2 (FFMA2 / FMUL2 / FADD2)        */
#define LLAMPC_KERNEL_K1B   3   /* persistent warp-task window kernel (single history, few CTAs x long windows)            */
#define LLAMPC_KERNEL_K1R   4   /* rolling kernel, grid over (candidates, vehicles)                                         */
#define LLAMPC_KERNEL_K1V   5   /* rolling kernel, one CTA per vehicle (N <= 2,048, Npad % 4 == 0), in-CTA top-K           */
#define LLAMPC_KERNEL_K1E   6   /* packed window kernel on equal warp shares: one 512-thread CTA per SM, every warp the same
                                   number of (64 candidates x 1 row) steps (single history, merge tree); on request only   */

typedef struct llampc_lookback_desc {
    /* bank: packed (llampc_bank_pack_h layout), device, 16-byte aligned; keys carry idx_offset + i */
    const float* bank; int N; int Npad; int idx_offset;
    int geom_shared;             /* non-zero: lf, lr identical for all candidates (rows carry valid stage-1 slip angles)   */
    float sin_arg_max;           /* max |C| pi/2 over the bank (llampc_bank_pack_h / llampc_bank_generate_f32 report it)   */
    int sine;                    /* LLAMPC_SIN_*                                                                            */
    /* history: [n_vehicles][hist_stride_rows][LLAMPC_HIST_ROW] floats; RECOMPUTE reads rows 0..W-1 of every vehicle,
       ROLLING reads row `slot` (unless row32_h is given)                                                                   */
    const float* hist; int W; int n_vehicles; int hist_stride_rows; double Ts;
    int mode;                    /* LLAMPC_LB_*                                                                              */
    int slot;                    /* ROLLING: ring slot of the newest row / error column; RECOMPUTE with row32_h: the slot
                                    the row is patched into (and stored to `hist`, which must then be writable)            */
    int emit;                    /* ROLLING: 0 = only store the error column (window still filling), no selection          */
    const float* row32_h;        /* HOST row (LLAMPC_HIST_ROW floats) riding in the kernel parameters, or NULL; single
                                    vehicle only                                                                            */
    float* err_ring;             /* ROLLING: [n_vehicles][LLAMPC_RING_ROWS(W)][Npad] floats, zeroed by the caller           */
    /* results */
    int K;                       /* 1..LLAMPC_LIST_LEN; 0 with out = NULL: scores only (needs avg_err)                      */
    float* avg_err;              /* [n_vehicles][N] or NULL                                                                 */
    llampc_key_t* out;           /* [n_vehicles][LLAMPC_LIST_LEN + 1]: out[0] = arg-min key, out[1..K] = ascending top-K,
                                    the remaining slots ~0ull                                                               */
    void* workspace;             /* llampc_lookback_plan(...).workspace_bytes bytes, device, 16-byte aligned, ZEROED once by
                                    the caller (the kernels leave their counters at zero); one workspace per stream         */
    unsigned long long workspace_bytes;
    /* multi-GPU (single history, RECOMPUTE): min-loc of out[0] across the GPUs of the box over NVLink peer memory, done by
       the warp that finishes the rank's merge tree, inside the same launch (no NCCL call on the path).
       peer_bufs: device array [world] of pointers, peer_bufs[q] = rank q's symmetric exchange buffer of 4 * world u64 words
       ([2 parities][world][2]: word 0 = low half of the key | seq << 32, word 1 = high half of the key | seq << 32 -- two
       self-validating 8-byte remote stores per peer, no system-scope fence), zero-initialised and mapped into this process (CUDA IPC / torch symmetric
       memory).  seq: tick counter >= 1, identical on every rank, incremented by the caller every call.  After the launch
       out[0] is the GLOBAL arg-min key on every rank (~0ull = no decision: a peer did not arrive within ~1 s);
       out[1..K] stay the rank-local top-K.  NULL = single GPU.                                                             */
    llampc_key_t* const* peer_bufs; int world; int rank; unsigned seq;
    /* overrides (0 = automatic) */
    int kernel;                  /* LLAMPC_KERNEL_*: force a kernel the shape supports (else LLAMPC_E_ARG)                  */
    int split;                   /* window splits per candidate inside a CTA: 1, 2, 4, 8, 16 (K1 / K1P)                     */
    int flags;                   /* LLAMPC_LB_FLAG_*                                                                        */
} llampc_lookback_desc_t;

/* Programmatic dependent launch (K1P, one history): the launch may begin -- bank loads, history staging, RK4 rows -- while
   the PREVIOUS launch on the same stream is still in its selection / merge-tree tail (every CTA of a K1P launch signals
   griddepcontrol.launch_dependents after its last RK4 row, and waits with griddepcontrol.wait for the previous grid to
   complete before it writes avg_err, the workspace or `out`).  For back-to-back launches that do not depend on each
   other's results (a sweep over banks); the results are identical. */
#define LLAMPC_LB_FLAG_PDL 1
/* K1P only: run the instantiation whose RK4 step takes the slip angles from the full-range atan (12 % more FMA-pipe work, no
   slip-tangent guard) instead of the |tan| <= 0.5 form with its per-candidate fallback.  For windows measured at low speed or
   in a drift, where most candidates would leave the fast form and be redone one by one (measured on a window at 0.1 - 0.3
   m/s: 135 us per tick instead of 46).  Scores agree with the default form to rounding; llampc_lookback_tick sets the flag
   itself from the rows of the ring when t->hard_h is given. */
#define LLAMPC_LB_FLAG_WIDE 2

typedef struct llampc_lookback_plan {
    int kernel;                  /* LLAMPC_KERNEL_* that a launch of this descriptor runs                                   */
    int split;                   /* window splits per candidate (K1 / K1P), rows per task (K1B), else 1 (K1E: 1)            */
    int sine;                    /* LLAMPC_SIN_SFU or LLAMPC_SIN_STRICT after resolving LLAMPC_SIN_AUTO                     */
    int grid_x, grid_y, block;
    int launches;                /* kernel launches per call: 1, or 2 when a vehicle has more than 1,024 per-CTA lists      */
    unsigned long long workspace_bytes;
} llampc_lookback_plan_t;

/* Resolves the automatic choices for this descriptor on the current device (needs a CUDA context for K1B's occupancy
 * query) without launching anything; desc->workspace may be NULL here. */
int llampc_lookback_plan(const llampc_lookback_desc_t* desc, llampc_lookback_plan_t* plan);

/* The launch.  Kernel choice (LLAMPC_KERNEL_AUTO):
 *   RECOMPUTE, one history       K1P when N >= 8,192 else K1, top-K finished INSIDE the launch by a tree of 32-way warp
 *                                merges that overlaps the integration (the root writes `out` and runs the NVLink min-loc);
 *                                K1B when that tiling would leave SMs idle while every thread walks >= 64 rows
 *   RECOMPUTE, many vehicles     K1 / K1P over (candidate tiles, vehicles): packed and window split chosen from
 *                                N * n_vehicles (the launch, not the vehicle, has to fill the GPU); the last CTA of a vehicle
 *                                to retire merges its lists inside the launch
 *   ROLLING                      K1V when N <= 2,048, Npad % 4 == 0 and rows come from `hist`; else K1R */
int llampc_lookback_launch(const llampc_lookback_desc_t* desc, llampc_stream_t stream);

/* K4  top-K (replaces avg_errors.argsort()[:K], run_nmpc_orca_llampc_rt.py:360).
 *   err [N] -> out_keys [K] ascending packed keys.  scratch: [n_ctas*K] keys with
 *   n_ctas = llampc_topk_scratch_ctas(N); counter: one unsigned, zero before the FIRST call
 *   (the kernel resets it). */
int llampc_topk_scratch_ctas(int N);
int llampc_topk_f32(const float* err, int N, int idx_offset, int K,
                    llampc_key_t* scratch, unsigned* counter, llampc_key_t* out_keys, llampc_stream_t stream);

/* fp64 re-score of a few finalists (same arithmetic as the reference, IEEE double on the device):
 *   bank64 [LLAMPC_NPARAM][N] doubles, hist64 [W][LLAMPC_HIST64_ROW], keys [n_fin] packed keys whose low
 *   32 bits - idx_offset select the candidates; out_err64 [n_fin] receives the window-mean error. */
int llampc_refine_f64(const double* bank64, int N, const double* hist64, int W, double Ts,
                      const llampc_key_t* keys, int n_fin, int idx_offset, double* out_err64,
                      llampc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * One MPC tick of the look-back step in ONE call (the body of run_nmpc_orca_llampc_rt.py:347-360): newest history row
 * into ring slot `slot`, llampc_lookback_launch over the window, Kt = max(K, n_refine) finalists, optional fp64 re-score
 * of the finalists, results to pinned host memory.  Every buffer is caller-owned; the struct only carries pointers
 * (device unless suffixed _h).
 * ------------------------------------------------------------------------------------------- */
typedef struct llampc_tick {
    const float* bank; int N; int Npad;
    float* hist;                    /* device ring [W][LLAMPC_HIST_ROW]                                    */
    const float* row32_h;           /* HOST row (LLAMPC_HIST_ROW floats) for ring slot `slot`, or NULL: it is
                                       passed to the kernel as a parameter (no separate H2D copy)          */
    int slot; int W; double Ts;
    int geom_shared;                /* as llampc_lookback_desc_t                                            */
    float sin_arg_max; int sine;    /* as llampc_lookback_desc_t                                            */
    int kernel; int split;          /* overrides, as llampc_lookback_desc_t (0 = automatic)                 */
    int idx_offset;
    float* avg_err;                 /* [N] or NULL (required when Kt > LLAMPC_LIST_LEN)                     */
    int K;                          /* top-K wanted by the caller (rt.py:360 uses 10)                      */
    int n_refine;                   /* 0 = no fp64 re-score; Kt = max(K, n_refine) finalists are produced  */
    const double* bank64;           /* [LLAMPC_NPARAM][N] (n_refine > 0)                                   */
    double* hist64;                 /* device ring [W][LLAMPC_HIST64_ROW] (n_refine > 0)                   */
    const double* row64_h;          /* HOST row for hist64 (kernel parameter of the re-score), or NULL     */
    llampc_key_t* result;           /* device, max(1 + 2*Kt, LLAMPC_LIST_LEN + 1) words:
                                       best key | Kt finalist keys | Kt fp64 scores                         */
    llampc_key_t* result_h;         /* pinned host, same layout + 1 word; with sync != 0 the finalists come back
                                       ordered by score (fp64 if re-scored), ties by lower index           */
    int sync;                       /* non-zero: wait for the result + host ordering before returning      */
    int zero_copy;                  /* non-zero (with n_refine > 0): the last re-score block writes the result straight
                                       into result_h (mapped pinned memory, 2 + 2*Kt words) and the host polls a sequence
                                       word instead of a D2H copy + stream synchronisation                        */
    llampc_key_t* const* peer_bufs; /* multi-GPU finalist all-gather over NVLink peer memory (needs n_refine > 0 and
                                       zero_copy): device array [peer_world] of every rank's symmetric buffer of
                                       2 * peer_world * 4*Kt zeroed words (every 64-bit payload travels as two
                                       self-validating 8-byte words: half | seq << 32); result_h then needs 2 + 2*Kt*peer_world
                                       words.  The ordered finalists returned are the GLOBAL ones; a peer that does not
                                       deliver within ~1 s makes llampc_lookback_finish return LLAMPC_E_PEER.  NULL = single GPU */
    int peer_world; int peer_rank;
    unsigned peer_seq;              /* tick counter >= 1, identical on all ranks, incremented by the caller  */
    unsigned long long pending_seq; /* internal: state between llampc_lookback_tick (sync = 0) and ..._finish        */
    int pending_words;
    float* err_ring;                /* [LLAMPC_RING_ROWS(W)][Npad] per-tick error columns + partial sums, zeroed by the caller
                                       (rolling mode only)                                                  */
    int rolling;                    /* 0: LLAMPC_LB_RECOMPUTE; 1: LLAMPC_LB_ROLLING (needs row32_h, err_ring and
                                       Kt <= LLAMPC_LIST_LEN; the fp64 re-score still walks the whole hist64 ring);
                                       2: rolling mode while the window is filling: store the column, no decision */
    void* workspace;                /* llampc_lookback_tick_workspace_bytes(t) bytes of device memory, 16-byte aligned,
                                       ZEROED once by the caller (counters, per-CTA lists, merge tree, top-K scratch) */
    unsigned long long workspace_bytes;
    void* mapped_dev; const void* mapped_for;   /* internal: device alias of result_h (cudaHostGetDevicePointer), cached */
    void* graph_state;              /* internal, NULL-initialised: the tick's scoring kernel and fp64 re-score are replayed
                                       as one CUDA graph whose kernel nodes are re-parameterised every tick (2 us of host
                                       enqueue time instead of 8 us); freed by llampc_lookback_tick_release            */
    unsigned char* hard_h;          /* HOST [W] or NULL: one flag per ring slot, "this row was measured at low speed or in a
                                       drift" (|vx| < 0.6 m/s or |vy| + 0.06 |w| > 0.4 |vx|).  The tick updates the flag of
                                       `slot` from row32_h and runs the LLAMPC_LB_FLAG_WIDE form of K1P while at least a tenth
                                       of the window is flagged (a caller that fills the ring itself sets the flags too)   */
    int n_hard;                     /* number of flagged slots (maintained with hard_h)                                     */
    void* replay_state;             /* internal, NULL-initialised: side stream, events and the device result ring of
                                       llampc_lookback_replay; freed by llampc_lookback_tick_release                       */
} llampc_tick_t;

/* Bytes of t->workspace for this tick configuration (bank, N, W, K, n_refine, rolling, overrides must be filled in);
 * needs a CUDA context.  Negative = LLAMPC_E_* / -1000 - cudaError_t. */
long long llampc_lookback_tick_workspace_bytes(const llampc_tick_t* t);

int llampc_lookback_tick(llampc_tick_t* t, llampc_stream_t stream);

/* Releases the resources llampc_lookback_tick attached to the struct (graph_state).  Call before discarding it. */
int llampc_lookback_tick_release(llampc_tick_t* t);

/* Asynchronous use: llampc_lookback_tick with t->sync = 0 only enqueues the work; llampc_lookback_finish waits for
 * the result (polling the zero-copy sequence word, else synchronising the stream) and orders the finalists;
 * llampc_lookback_decode turns them into plain arrays (idx_out / score_out [max(K, n_refine) or 1], *n_valid = 0 while
 * a rolling window is filling).  The host is free between tick and finish, e.g. for the next NMPC solve. */
int llampc_lookback_finish(llampc_tick_t* t, llampc_stream_t stream);
int llampc_lookback_decode(const llampc_tick_t* t, long long* idx_out, double* score_out, int* n_valid);

/* Layout probes for bindings that mirror llampc_tick_t by hand: sizeof, and offsetof of
 * Ts (0), avg_err (1), result_h (2), peer_seq (3), rolling (4), workspace (5). */
int llampc_tick_sizeof(void);
int llampc_tick_offsetof(int which);
int llampc_lookback_desc_sizeof(void);

/* The whole body of run_nmpc_orca_llampc_rt.py:347-360 in one call from three fp64 host vectors: packs the
 * transition (x_k, u_k) -> x_k1 into t->row32_h / t->row64_h (which must point to writable host scratch), runs
 * llampc_lookback_tick with sync, and decodes the ordered finalists: idx_out / score_out [max(K, n_refine) or 1],
 * *n_valid = number of valid entries (0 while a rolling window is still filling). */
int llampc_lookback_push(llampc_tick_t* t, const double* x_k, const double* u_k, const double* x_k1,
                         double lf_shared, double lr_shared, long long* idx_out, double* score_out, int* n_valid,
                         llampc_stream_t stream);

/* T consecutive ticks from host arrays in ONE call, PIPELINED: a replay of a recorded run (the loop of
 * run_nmpc_orca_llampc_rt.py:347-360 over a dataset).  The look-back's inputs are measurements -- no tick depends on the
 * decision of an earlier one -- so up to `depth` ticks are in flight: while the GPU works on tick i the host packs and
 * enqueues ticks i + 1 .. i + depth - 1.  Every tick still carries its own row in the launch parameters and hands its
 * decision back through its own result slot; decisions are identical to T calls of llampc_lookback_push.
 *   x_k [T][6], u_k [T][2], x_k1 [T][x1_stride] (x1_stride >= 4) fp64 host arrays, oldest first; tick i uses ring slot
 *   (first_slot + i) % W.  The window must be full already (t->rolling = 0, or 1 with a primed error ring).
 *   rows32_h [W][LLAMPC_HIST_ROW], rows64_h [W][LLAMPC_HIST64_ROW]: writable host staging rings (row i is packed into its slot)
 *   slots_h [depth]: pinned, mapped host result buffers, each as large as t->result_h (t->zero_copy and n_refine > 0 needed)
 *   peer_seq: in/out, NULL on one GPU: the last tick counter used (incremented once per tick, identically on every rank)
 *   idx_out / score_out [T][kt] with kt = max(K, n_refine), n_valid [T].
 * In recompute mode the scoring kernels follow each other on `stream` with programmatic dependent launch and the fp64
 * re-score of every tick runs on an internal non-blocking side stream behind an event, on the tick's own device result
 * buffer (t->replay_state, freed by llampc_lookback_tick_release); before returning, `stream` is made to wait for the side
 * stream.  LLAMPC_REPLAY_TRACE=1 prints the host-side phase times per tick to stderr. */
int llampc_lookback_replay(llampc_tick_t* t, const double* x_k, const double* u_k, const double* x_k1, int x1_stride, int T,
                           int first_slot, double lf_shared, double lr_shared, float* rows32_h, double* rows64_h,
                           llampc_key_t* const* slots_h, int depth, unsigned* peer_seq, long long* idx_out,
                           double* score_out, int* n_valid, llampc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * One RK4 step for N (model, state, input) triples: Model._integrate_batch (llampc/models/model.py:32-40)
 * -> odeintRK4_batch (llampc/utils/rk6.py:50-68) -> Dynamic._diffequation_batch (dynamic.py:98-154).
 *   x64 [N][6] (or one row if x_shared), u64 [N][2] (or one row if u_shared), out64 [N][out_cols] with
 *   out_cols = 6 (_integrate_batch) or 4 (evaluate_models_vectorized returns [:,0:4]).
 *   The increment is integrated in fp32, the final x0 + increment sum is formed in fp64.
 * ------------------------------------------------------------------------------------------- */
int llampc_rk4_batch_f32(const float* bank, int N, int Npad, const double* x64, int x_shared,
                         const double* u64, int u_shared, double Ts, double* out64, int out_cols,
                         llampc_stream_t stream);

/* Right-hand side only: Dynamic._diffequation_batch (dynamic.py:98-115), fp32 arithmetic, f64 in/out. */
int llampc_rhs_batch_f32(const float* bank, int N, int Npad, const double* x64, int x_shared,
                         const double* u64, int u_shared, double* out64, llampc_stream_t stream);

/* Forces and slip angles: Dynamic.calc_forces_batch(x, u, return_slip=True) (dynamic.py:117-154).
 * out64 [N][5] = Ffy Frx Fry alphaf alphar. */
int llampc_forces_batch_f32(const float* bank, int N, int Npad, const double* x64, int x_shared,
                            const double* u64, int u_shared, double* out64, llampc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K2  look-ahead rollout: M models x K control sequences x H RK4 steps (Model._integrate_batch chained,
 * llampc/models/model.py:32-40) scored with the NMPC objective (llampc/mpc/nmpc.py:48,66-71,111):
 *   J = sum_{h=1..H} (p_h-xref_h)'Q(p_h-xref_h) + (p_H-xref_H)'P(p_H-xref_H) + sum_{h=0..H-1} du_h' R du_h
 *   bank       packed bank with Mpad rows per group; model_idx [M] bank row of each model, or NULL = identity
 *   x0         [n_x0][6] doubles (n_x0 = 1: shared start state, or M)
 *   U          [K][H][2] floats, or [M][K][H][2] if per_model_flags & 1
 *   xref       [H+1][2] floats (row h = reference position at step h), or [M][H+1][2] if per_model_flags & 2
 *   uprev      [2] floats, or [M][2] if per_model_flags & 4   (diagnostics: per_model_flags & 8 forces the scalar
 *              kernel with the general branchy step, & 16 the scalar kernel with its straight-line step, instead of
 *              the packed two-models-per-thread kernels; & 32: programmatic dependent launch of the shared-layout kernel for
 *              back-to-back rollouts that do not consume each other's results -- the next launch's rollouts start beside
 *              this launch's last wave, bit-identical results)
 *              shared U / xref tables are fetched with 16-byte-granular bulk copies: the buffers must be
 *              16-byte aligned and readable up to the next multiple of 16 bytes
 *   qrp_h      HOST pointer, 6 floats = Q00 Q11 R00 R11 P00 P11
 *   J          [M][K] floats;  best_k [M] ints (first index on ties);  x_final [M][K][6] doubles or NULL
 *   x_traj     [M][K][H+1][6] doubles or NULL: every state of every rollout (meant for small K, e.g. the best
 *              sequence per model as a warm start of the NLP: xvars = [x(:,0..H); u(:,0..H-1)], llampc/mpc/nmpc.py:113-117)
 * ------------------------------------------------------------------------------------------- */
int llampc_lookahead_rollout_f32(const float* bank, int Mpad, const int* model_idx, int M,
                                 const double* x0, int n_x0, const float* U, int K, int H,
                                 const float* xref, const float* uprev, int per_model_flags,
                                 const float* qrp_h, double Ts, float* J, int* best_k, double* x_final,
                                 double* x_traj, llampc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Planner: ConstantSpeed (llampc/mpc/planner.py:12-67) for V vehicles, fp64.  Tables (device, doubles, 16-byte aligned)
 * come from a raceline (Track._load_raceline, llampc/tracks/track.py:52-83):
 *   s        [n]  cumulative arc length of the raceline points (Spline2D.s)
 *   xy       [n][2]   raceline points
 *   coef_xy  [n-1][8]            per segment: a b c d of x(s), then a b c d of y(s)
 *   coef_vp  [n_mu][n-1][8]      per speed-profile PAIR j and segment: a b c d of v_{(j-1) mod n_mu}(s), then of v_j(s)
 *                                (the two profiles the friction interpolation of planner.py:49-62 blends for
 *                                mus[j-1] <= mu <= mus[j]; j-1 wraps like Python's spline_v[i-1] for j = 0)
 *   mus      [n_mu]   friction level of each speed profile (ascending)
 * PADDING: each vehicle's window of the tables is copied into shared memory by fixed-size bulk copies, so the
 * allocations must be readable past their last used element: s  n + LLAMPC_PLAN_SPAD doubles,  coef_xy
 * (n - 1 + LLAMPC_PLAN_WSEG) rows,  coef_vp  n_mu planes of (n - 1 + LLAMPC_PLAN_WSEG) rows each (the plane stride
 * INCLUDES the padding).  The padding values are never used.
 * states [V][6] (x, y and vx are used), projidx_in [V], curr_mu [V] (or one value if mu_shared).
 * Outputs (any may be NULL): xref32 [V][N+1][2] floats (feeds llampc_lookahead_rollout_f32 with
 * per_model_flags & 2), xref64 [V][N+1][2] doubles, projidx_out [V], vr_out [V].
 * ------------------------------------------------------------------------------------------- */
#define LLAMPC_PLAN_WSEG 48   /* table segments staged per vehicle (a march leaving them re-stages the window) */
#define LLAMPC_PLAN_SPAD 66   /* arc-length marks staged per vehicle = padding of the s allocation (WSEG + 18)     */
int llampc_planner_constant_speed_f64(const double* s, const double* xy, const double* coef_xy, const double* coef_vp,
                                      const double* mus, int n, int n_mu, const double* states, int V,
                                      const int* projidx_in, const double* curr_mu, int mu_shared, int N, double Ts,
                                      double scale, float* xref32, double* xref64, int* projidx_out, double* vr_out,
                                      llampc_stream_t stream);

/* K3  plant step for V independent vehicles: Model._integrate -> odeintRK6 (llampc/models/model.py:18-30,
 * llampc/utils/rk6.py:13-28), fp64 throughout.  params64 [V][LLAMPC_NPARAM], x64 [V][6], u64 [V][2] -> out64 [V][6]. */
int llampc_plant_rk6_f64(const double* params64, int V, const double* x64, const double* u64, double Ts,
                         double* out64, llampc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Monte-Carlo closed loop (thousands of independent vehicles, everything device-resident, no host round trip
 * per tick).  Layouts: per-vehicle history rings hist [V][W][LLAMPC_HIST_ROW] / hist64 [V][W][LLAMPC_HIST64_ROW]
 * (llampc_lookback_launch with n_vehicles = V, hist_stride_rows = W).
 * ------------------------------------------------------------------------------------------- */

/* Device twin of llampc_hist_row_pack_h for V vehicles: x_k [V][6], u_k [V][2], x_k1 [V][6] -> ring slot `slot`. */
int llampc_pack_rows_f64(const double* x_k, const double* u_k, const double* x_k1, int V, double Ts,
                         double lf_shared, double lr_shared, int slot, int W, float* hist, double* hist64,
                         llampc_stream_t stream);

/* Friction estimate from the K best candidates (run_nmpc_orca_llampc_rt.py:326-344): mean Dr, Df of the top-K appended
 * to the per-vehicle lists Drs_preds / Dfs_preds, MU_pred = (mean of the last `smoothing` entries of each) / (g m) (:341),
 * display value = exponential smoother (alpha, :103-113) x gain (:344).
 *   topk [V][topk_stride] keys ([0] arg-min, [1..K] top-K) of the PREVIOUS tick's look-back (ind_best_KM, :360)
 *   state [V][2*smoothing+3] doubles, zeroed by the caller, then seeded once with llampc_mu_seed_f64
 *   mu_raw [V]      MU_pred, the raw moving average: what ConstantSpeed receives as curr_mu from tick W + 2 on (:278-280)
 *   mu_display [V]  or NULL: smoothed x gain, the value the reference only logs / plots (MU_preds, :344, :465)
 * llampc_mu_seed_f64 appends the n_seed = W + 1 warm-up entries of :326-330 (seed_dr = mu_init m 9.8 lr / (lf + lr),
 * seed_df = mu_init m 9.8 lf / (lf + lr): note g = 9.8 there and 9.81 in the estimate). */
int llampc_mu_estimate_f64(const llampc_key_t* topk, int topk_stride, int K, int idx_offset,
                           const double* bank64, int N, int V, int smoothing, double alpha, double gain,
                           double g, double* state, double* mu_raw, double* mu_display, llampc_stream_t stream);
int llampc_mu_seed_f64(double* state, int V, int smoothing, int n_seed, double seed_dr, double seed_df,
                       llampc_stream_t stream);

/* Control samples U [V][K][H][2] = clip(nominal [V][H][2] + eps [K][H][2]); box_h (HOST) = pwm_min pwm_max steer_min steer_max
 * (limits of llampc/params/orca.py:29-35). */
int llampc_sample_controls_f32(const float* nominal, const float* eps, int V, int K, int H, const float* box_h,
                               float* U, llampc_stream_t stream);

/* Best-of-K controller step: u_applied [V][2] (doubles) = U[v][best_k[v]][0]; uprev [V][2] likewise (floats);
 * nominal [V][H][2] = that sequence shifted by one step (last input repeated). */
int llampc_apply_best_f32(const float* U, const int* best_k, int V, int K, int H, float* nominal, float* uprev,
                          double* u_applied, llampc_stream_t stream);

/* Monte-Carlo scenario glue (one launch each instead of a dozen element-wise framework launches per tick).
 * friction schedule, 'sudden' style of run_nmpc_orca_llampc_nrt_avg_runs.py:163-166: while drop_start[v] < *t_dev <
 * drop_start[v] + drop_len, columns col0 .. col0 + ncols - 1 (Df, Dr = 8, 9) of plant [V][LLAMPC_NPARAM] doubles are
 * multiplied by 1 - drop_rate.
 * advance, every part optional: model_idx[v] = low word of topk[v * topk_stride] (skipped when topk or model_idx is
 * NULL), x <- x_next ([V][6] doubles; both NULL skips it), *t_dev += Ts (NULL skips it). */
int llampc_mc_friction_schedule_f64(double* plant, int V, int col0, int ncols, const double* drop_start,
                                    double drop_len, double drop_rate, const double* t_dev, llampc_stream_t stream);
int llampc_mc_advance_tick_f64(const llampc_key_t* topk, int topk_stride, int* model_idx, double* x,
                               const double* x_next, int V, double* t_dev, double Ts, llampc_stream_t stream);

/* Bank generation / resampling on the device (run_nmpc_orca_llampc_rt.py:145-179: parameter = centre x (1 + sigma
 * randn) for every varied parameter of every model).  center_h / sigma_h: HOST arrays of LLAMPC_NPARAM doubles
 * (sigma = 0: parameter not varied).  Counter-based Philox4x32-10 + Box-Muller: candidate i / parameter j get the same
 * draw for a given seed whatever the launch.  Writes the packed bank [4][Npad] float4 and, if not NULL, bank64
 * [LLAMPC_NPARAM][N].  Re-centring the bank on a selected candidate = calling it with that candidate's parameters. */
int llampc_bank_generate_f32(const double* center_h, const double* sigma_h, int N, int Npad,
                             unsigned long long seed, float* packed, double* bank64,
                             float* sin_arg_max /* DEVICE float, zeroed by the caller, or NULL: max(|Cf|, |Cr|) pi/2 */,
                             llampc_stream_t stream);

/* Measurement helper: runs an FMA-bound loop of `iters` iterations on every SM and stores, for one thread,
 * out2[0] = elapsed SM cycles (clock64) and out2[1] = elapsed nanoseconds (globaltimer): the SM clock actually
 * sustained under load, used for the roofline denominator "at the measured clock".  sink: one device float. */
int llampc_clock_probe(int iters, unsigned long long* out2, float* sink, llampc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LLAMPC_B200_H */
