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

#define LLAMPC_ABI_VERSION 4

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
int llampc_bank_pack_h(const double* const* params_h, const int* is_array, int N, int Npad, float* packed_h);

/* One history row from one measured transition (x_k, u_k) -> x_k1 (the tick body of
 * run_nmpc_orca_llampc_rt.py:347-349 needs exactly these three vectors).  All trigonometry and the
 * measured increments are formed in fp64 here so the fp32 kernel never subtracts O(1) numbers.
 * lf_shared/lr_shared: the bank-wide lf, lr when geometry is not varied (else NaN): enables the
 * candidate-invariant stage-1 slip angles.  row32_h: LLAMPC_HIST_ROW floats, row64_h (may be NULL):
 * LLAMPC_HIST64_ROW doubles for llampc_refine_f64. */
int llampc_hist_row_pack_h(const double* x_k, const double* u_k, const double* x_k1, double Ts,
                           double lf_shared, double lr_shared, float* row32_h, double* row64_h);

/* ---------------------------------------------------------------------------------------------
 * K1  look-back window: for every candidate, W one-step RK4 predictions re-anchored at the measured
 * states, mean squared error over (x, y, psi, vx) and over the window, fused block arg-min.
 * Replaces evaluate_models_vectorized (llampc/mpc/evaluate_models_vectorized.py:4-23) called once per
 * tick + errors / error_windows / mean / argmin of run_nmpc_orca_llampc_rt.py:349-358.
 *   bank      packed bank (llampc_bank_pack_h layout) on the device, 16-byte aligned
 *   hist      [n_vehicles][hist_stride_rows][LLAMPC_HIST_ROW] floats; the first W rows of each vehicle are used
 *   avg_err   [n_vehicles][N] or NULL
 *   best_key  [n_vehicles] or NULL; MUST be preset to ~0ull by the caller (llampc_fill_keys); receives the
 *             min over candidates of (float_bits(avg_err)<<32 | idx_offset+i)  (np.argmin tie-break)
 *   cta_lists [n_vehicles][n_lists][LLAMPC_LIST_LEN] or NULL, n_lists = llampc_lookback_num_lists(N, W, split):
 *             the ascending LLAMPC_LIST_LEN smallest keys of every CTA (input of llampc_topk_merge_lists)
 *   geom_shared  non-zero: rows carry valid stage-1 slip angles (lf, lr identical for all candidates)
 *   split     window splits per candidate inside a CTA (1, 2, 4, 8, 16) or 0 = choose from N, W; + 32 selects the
 *             MUFU.SIN (SFU) tyre sine instead of the FMA-pipe polynomial (faster, ~2x the fp32 score error)
 * Banks of >= 8,192 candidates run the packed kernel K1p: two candidates per thread in f32x2 arithmetic (FFMA2 /
 * FMUL2 / FADD2), i.e. 256 / split candidates per CTA instead of 128 / split; llampc_lookback_num_lists accounts for
 * it.  Same operations per candidate, so the scores differ from the scalar kernel's only by re-association of a few
 * signs (LLAMPC_K1_PACKED=0 / 1 in the environment forces the scalar / packed kernel for every size).
 * ------------------------------------------------------------------------------------------- */
int llampc_lookback_window_f32(const float* bank, int N, int Npad,
                               const float* hist, int W, int n_vehicles, int hist_stride_rows, double Ts,
                               float* avg_err, llampc_key_t* best_key, llampc_key_t* cta_lists, int idx_offset,
                               int geom_shared, int split, llampc_stream_t stream);
int llampc_lookback_num_lists(int N, int W, int split);

/* K1 with the top-K finished inside the same launch: the last CTA of each vehicle to retire (atomic ticket) merges
 * the per-CTA lists.  ticket [n_vehicles] unsigned, zero before the first call (self-resetting);
 * out [n_vehicles][LLAMPC_LIST_LEN + 1] as llampc_topk_merge_lists.  With more than 1,024 lists per vehicle the call
 * falls back to two launches (K1 + llampc_topk_merge_lists). */
int llampc_lookback_window_topk_f32(const float* bank, int N, int Npad, const float* hist, int W,
                                    int n_vehicles, int hist_stride_rows, double Ts, float* avg_err,
                                    llampc_key_t* best_key, llampc_key_t* cta_lists, int idx_offset,
                                    int geom_shared, int split, int K, unsigned* ticket, llampc_key_t* out,
                                    llampc_stream_t stream);

/* K1r for many vehicles (Monte-Carlo layout): the newest row of vehicle v is ring slot `slot` of hist [V][W][20]
 * (written by llampc_pack_rows_f64), err_ring is [V][W][Npad], avg_err [V][N] or NULL, best_key [V] (armed),
 * cta_lists [V][ceil(N/128)][LLAMPC_LIST_LEN], ticket [V] zeroed, out [V][LLAMPC_LIST_LEN + 1].  emit = 0 only stores
 * the error columns (windows still filling); K = 0 skips the top-K.
 * Banks of N <= 2,048 with Npad % 4 == 0, K > 0 and out != NULL run K1v (one CTA per vehicle, top-K by threshold filter in
 * shared memory, SFU tyre sine): same scores and keys, but best_key / cta_lists / ticket are then neither read nor written
 * (out[0] is the arg-min key either way).  The environment switch LLAMPC_K1R_CTA=0 keeps K1r.
 * geom_shared: bit 0 = lf, lr bank-wide (stage-1 slip angles ride in the rows); bit 1 = strict mode, the FMA-pipe polynomial
 * tyre sine instead of MUFU.SIN (for banks as wide as sigma = 2, plot_comp_time.py:178-192, where single candidates with
 * C > 9 reach a relative score error of 1.2e-4 with the SFU sine against 3.6e-5 strict: tools/gpu_wide_bank_check.py). */
int llampc_lookback_rolling_multi_f32(const float* bank, int N, int Npad, const float* hist, int n_vehicles,
                                      int slot, int W, double Ts, float* err_ring, float* avg_err,
                                      llampc_key_t* best_key, llampc_key_t* cta_lists, int idx_offset,
                                      int geom_shared, int emit, int K, unsigned* ticket, llampc_key_t* out,
                                      llampc_stream_t stream);

/* K1 + top-K + multi-GPU min-loc in ONE launch per rank, over NVLink peer memory (no NCCL on the path).
 *   peer_bufs  device array [world] of pointers: peer_bufs[q] = rank q's symmetric exchange buffer of
 *              4 * world u64 words ([2 parities][world][key, sequence]), zero-initialised, mapped into this process
 *              (CUDA IPC / torch symmetric memory); peer_bufs[rank] is this rank's own buffer
 *   seq        tick counter, identical on every rank, incremented by the caller every call (>= 1)
 * After the launch out[0] holds the GLOBAL arg-min key on every rank (~0ull = no decision: a peer did not arrive within ~1 s);
 * out[1..K] stay the rank-local top-K.  With more than 1,024 per-CTA lists (shards above 131,072 candidates per
 * split) the exchange is carried by the stand-alone merge kernel instead (two launches, still no NCCL call); the
 * limit is 8,192 lists. */
int llampc_lookback_window_topk_peer_f32(const float* bank, int N, int Npad, const float* hist, int W,
                                         int hist_stride_rows, double Ts, float* avg_err,
                                         llampc_key_t* best_key, llampc_key_t* cta_lists, int idx_offset,
                                         int geom_shared, int split, int K, unsigned* ticket,
                                         llampc_key_t* out, llampc_key_t* const* peer_bufs, int world,
                                         int rank, unsigned seq, llampc_stream_t stream);

/* One-launch look-back tick (single history, any N and W): same scores and same selection as
 * llampc_lookback_window_topk_f32, with the top-K finished INSIDE the launch by a tree of 32-way warp merges: warp 0 of
 * every CTA publishes the CTA's 16 smallest keys, the last arrival of every 32 lists merges them (heads in registers,
 * REDUX minima) and climbs one level, the warp that produces the root writes `out`.  The merges overlap the
 * integration; the serial tail after the last RK4 step is the two or three merges on the path to the root instead of
 * one CTA walking every list, and there is no list-count limit (1,048,576 candidates: one launch).
 * The kernel underneath is chosen from the shape: K1 (llampc_lookback_window_f32's kernel, bit-identical scores), or
 * K1b when K1's tiling would leave SMs idle while every thread walks a long window (fewer CTAs than SMs and >= 64 rows
 * per thread): K1b runs persistent CTAs whose warps pull warp-tasks (32 candidates x R window rows) from an atomic
 * counter; the row chunks of a warp-group are combined in row order by the last task to arrive (deterministic).
 *   workspace  llampc_lookback_balanced_workspace_bytes(N, W) bytes of device memory, 16-byte aligned, ZEROED once by
 *              the caller before the first call (the kernels leave their counters at zero; the counter offsets depend
 *              on N only, so W may change between calls); one workspace per stream
 *   fast_sin   non-zero: MUFU.SIN tyre sine (as split + 32 of llampc_lookback_window_f32)
 *   K          1..LLAMPC_LIST_LEN;  out [LLAMPC_LIST_LEN + 1]: out[0] = arg-min key, out[1..K] = ascending top-K,
 *              the remaining slots ~0ull
 *   peer_bufs / world / rank / seq   as llampc_lookback_window_topk_peer_f32 (NVLink min-loc of out[0], done by the
 *              root warp inside the launch); peer_bufs = NULL: single GPU
 * Replaces evaluate_models_vectorized (llampc/mpc/evaluate_models_vectorized.py:4-23) + errors / mean / argmin /
 * argsort[:K] of run_nmpc_orca_llampc_rt.py:349-360. */
long long llampc_lookback_balanced_workspace_bytes(int N, int W);
int llampc_lookback_window_balanced_f32(const float* bank, int N, int Npad, const float* hist, int W, double Ts,
                                        float* avg_err, int idx_offset, int geom_shared, int fast_sin, int K,
                                        void* workspace, unsigned long long workspace_bytes, llampc_key_t* out,
                                        llampc_key_t* const* peer_bufs, int world, int rank, unsigned seq,
                                        llampc_stream_t stream);

/* K1r  rolling window, the reference's own bookkeeping (error_windows = np.roll(...); [:, -1] = errors; mean,
 * run_nmpc_orca_llampc_rt.py:349-358): one RK4 step per candidate for the newest transition (row32_h, HOST pointer,
 * passed as kernel parameter), error column `slot` of err_ring [W][Npad] replaced, window mean re-summed from the
 * ring.  emit = 0: only store the column (window not full yet).  CTA lists: ceil(N/128) lists. */
int llampc_lookback_rolling_f32(const float* bank, int N, int Npad, const float* row32_h, int slot, int W,
                                double Ts, float* err_ring, float* avg_err, llampc_key_t* best_key,
                                llampc_key_t* cta_lists, int idx_offset, int geom_shared, int emit,
                                llampc_stream_t stream);

/* Fused top-K (K <= LLAMPC_LIST_LEN): K-way merge of the per-CTA lists of K1.  Per vehicle v:
 *   out[v][0] = best_key[v] (which is then re-armed to ~0ull for the next tick; skipped if best_key is NULL),
 *   out[v][1..K] = ascending top-K keys; out has LLAMPC_LIST_LEN + 1 keys per vehicle.
 * Replaces avg_errors.argsort()[:K] (run_nmpc_orca_llampc_rt.py:360) without re-reading avg_err. */
int llampc_topk_merge_lists(const llampc_key_t* cta_lists, int n_lists, int n_vehicles, int K,
                            llampc_key_t* best_key, llampc_key_t* out, llampc_stream_t stream);

int llampc_fill_keys(llampc_key_t* keys, int n, llampc_stream_t stream);   /* keys[i] = ~0ull */

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
 * One MPC tick of the look-back step in ONE call (the body of run_nmpc_orca_llampc_rt.py:347-360):
 * upload the newest history row into ring slot `slot`, K1 over the whole window, K4 top-Kt with
 * Kt = max(K, n_refine), optional fp64 re-score of the Kt finalists, results to pinned host memory.
 * Every buffer is caller-owned; the struct only carries pointers (device unless suffixed _h).
 * ------------------------------------------------------------------------------------------- */
typedef struct llampc_tick {
    const float* bank; int N; int Npad;
    float* hist;                    /* device ring [W][LLAMPC_HIST_ROW]                                    */
    const float* row32_h;           /* HOST row (LLAMPC_HIST_ROW floats) for ring slot `slot`, or NULL: it is
                                       passed to K1 as a kernel parameter (no separate H2D copy)           */
    int slot; int W; double Ts;
    int geom_shared; int split; int idx_offset;
    float* avg_err;                 /* [N] or NULL (fused path only)                                       */
    llampc_key_t* best_key;         /* [1]; armed (~0ull) once by the caller with llampc_fill_keys         */
    int K;                          /* top-K wanted by the caller (rt.py:360 uses 10)                      */
    int n_refine;                   /* 0 = no fp64 re-score; Kt = max(K, n_refine) finalists are produced  */
    llampc_key_t* cta_lists;        /* [n_lists][LLAMPC_LIST_LEN] or NULL; with Kt <= LLAMPC_LIST_LEN selects the
                                       fused path (K1 + list merge: two launches per tick)                 */
    llampc_key_t* topk_scratch; unsigned* topk_counter;   /* only for the unfused path (Kt > LLAMPC_LIST_LEN)    */
    const double* bank64;           /* [LLAMPC_NPARAM][N] (n_refine > 0)                                   */
    double* hist64;                 /* device ring [W][LLAMPC_HIST64_ROW] (n_refine > 0)                   */
    const double* row64_h;          /* HOST row for hist64 (kernel parameter of the re-score), or NULL     */
    llampc_key_t* result;           /* device, 1 + 2*Kt words: best key | Kt finalist keys | Kt fp64 scores */
    llampc_key_t* result_h;         /* pinned host, same layout; with sync != 0 the finalists come back
                                       ordered by score (fp64 if re-scored), ties by lower index           */
    int sync;                       /* non-zero: cudaStreamSynchronize + host ordering before returning    */
    unsigned* ticket;               /* [2], zero-initialised: finish the top-K inside K1 (one launch per tick) when
                                       the bank yields <= 1,024 per-CTA lists; NULL = always use the merge kernel */
    int zero_copy;                  /* non-zero (with sync, n_refine > 0, ticket): the last re-score block writes the
                                       result straight into result_h (mapped pinned memory, 2 + 2*Kt words) and the
                                       host polls a sequence word instead of a D2H copy + stream synchronisation  */
    llampc_key_t* const* peer_bufs; /* multi-GPU finalist all-gather over NVLink peer memory (needs n_refine > 0, sync,
                                       zero_copy): device array [peer_world] of every rank's symmetric buffer of
                                       2 * peer_world * (2*Kt + 1) zeroed words; result_h then needs 2 + 2*Kt*peer_world
                                       words.  The ordered finalists returned are the GLOBAL ones.  NULL = single GPU */
    int peer_world; int peer_rank;
    unsigned peer_seq;              /* tick counter >= 1, identical on all ranks, incremented by the caller  */
    unsigned long long pending_seq; /* internal: state between llampc_lookback_tick (sync = 0) and ..._finish        */
    int pending_words;
    float* err_ring;                /* [W][Npad] per-tick error columns (rolling mode only)                */
    int rolling;                    /* 0: recompute the whole window from the history ring (K1);
                                       1: rolling mode (K1r): integrate only the newest row, replace ring column
                                          `slot`, re-sum the ring; needs row32_h, cta_lists and Kt <= LLAMPC_LIST_LEN
                                          (the fp64 re-score still walks the whole hist64 ring);
                                       2: rolling mode while the window is filling: store the column, no decision */
    void* workspace;                /* device scratch of llampc_lookback_balanced_workspace_bytes(N, W) bytes, zeroed once by
                                       the caller, or NULL.  Non-NULL (with cta_lists, rolling = 0, 0 < Kt <=
                                       LLAMPC_LIST_LEN) runs the one-launch tick with the in-kernel tree merge
                                       (llampc_lookback_window_balanced_f32) instead of K1 + list merge            */
    unsigned long long workspace_bytes;
    void* mapped_dev; const void* mapped_for;   /* internal: device alias of result_h (cudaHostGetDevicePointer), cached */
    void* graph_state;              /* internal, NULL-initialised: the tick's scoring kernel and fp64 re-score are replayed
                                       as one CUDA graph whose kernel nodes are re-parameterised every tick (2 us of host
                                       enqueue time instead of 8 us); freed by llampc_lookback_tick_release            */
} llampc_tick_t;

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
 * Ts (0), cta_lists (1), result_h (2), peer_seq (3), rolling (4). */
int llampc_tick_sizeof(void);
int llampc_tick_offsetof(int which);

/* The whole body of run_nmpc_orca_llampc_rt.py:347-360 in one call from three fp64 host vectors: packs the
 * transition (x_k, u_k) -> x_k1 into t->row32_h / t->row64_h (which must point to writable host scratch), runs
 * llampc_lookback_tick with sync, and decodes the ordered finalists: idx_out / score_out [max(K, n_refine) or 1],
 * *n_valid = number of valid entries (0 while a rolling window is still filling). */
int llampc_lookback_push(llampc_tick_t* t, const double* x_k, const double* u_k, const double* x_k1,
                         double lf_shared, double lr_shared, long long* idx_out, double* score_out, int* n_valid,
                         llampc_stream_t stream);

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
 *   uprev      [2] floats, or [M][2] if per_model_flags & 4   (per_model_flags & 8: diagnostics, force the
 *              general branchy step instead of the straight-line one)
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
 * Planner: ConstantSpeed (llampc/mpc/planner.py:12-67) for V vehicles, fp64.  Tables (device, doubles) come from
 * a raceline (Track._load_raceline, llampc/tracks/track.py:52-83):
 *   s    [n]      cumulative arc length of the raceline points (Spline2D.s)
 *   xy   [n][2]   raceline points
 *   coef [n-1][4*(2+n_mu)]  per segment: a b c d of x(s), of y(s), then of each speed profile v_j(s)
 *   mus  [n_mu]   friction level of each speed profile (ascending)
 * states [V][6] (x, y and vx are used), projidx_in [V], curr_mu [V] (or one value if mu_shared).
 * Outputs (any may be NULL): xref32 [V][N+1][2] floats (feeds llampc_lookahead_rollout_f32 with
 * per_model_flags & 2), xref64 [V][N+1][2] doubles, projidx_out [V], vr_out [V].
 * ------------------------------------------------------------------------------------------- */
int llampc_planner_constant_speed_f64(const double* s, const double* xy, const double* coef, const double* mus,
                                      int n, int n_mu, const double* states, int V, const int* projidx_in,
                                      const double* curr_mu, int mu_shared, int N, double Ts, double scale,
                                      float* xref32, double* xref64, int* projidx_out, double* vr_out,
                                      llampc_stream_t stream);

/* K3  plant step for V independent vehicles: Model._integrate -> odeintRK6 (llampc/models/model.py:18-30,
 * llampc/utils/rk6.py:13-28), fp64 throughout.  params64 [V][LLAMPC_NPARAM], x64 [V][6], u64 [V][2] -> out64 [V][6]. */
int llampc_plant_rk6_f64(const double* params64, int V, const double* x64, const double* u64, double Ts,
                         double* out64, llampc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Monte-Carlo closed loop (thousands of independent vehicles, everything device-resident, no host round trip
 * per tick).  Layouts: per-vehicle history rings hist [V][W][LLAMPC_HIST_ROW] / hist64 [V][W][LLAMPC_HIST64_ROW]
 * (use llampc_lookback_window_f32 with n_vehicles = V, hist_stride_rows = W).
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
                             unsigned long long seed, float* packed, double* bank64, llampc_stream_t stream);

/* Measurement helper: runs an FMA-bound loop of `iters` iterations on every SM and stores, for one thread,
 * out2[0] = elapsed SM cycles (clock64) and out2[1] = elapsed nanoseconds (globaltimer): the SM clock actually
 * sustained under load, used for the roofline denominator "at the measured clock".  sink: one device float. */
int llampc_clock_probe(int iters, unsigned long long* out2, float* sink, llampc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LLAMPC_B200_H */
