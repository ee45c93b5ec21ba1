"""Look-back adaptation step on the GPU.

The reference has no function for this: the logic is inline in the experiment scripts
(run_nmpc_orca_llampc_rt.py:347-366; same block in ..._nrt.py:421-441, ..._nrt_avg_runs.py:446-466).
``LookBack.push`` is that block: it takes the transition measured at one MPC tick, scores every candidate
model of the bank over the last W transitions and returns the arg-min and the K best candidates.

Window semantics (= the reference): W independent one-step RK4 predictions, each re-anchored at the measured
state; no decision until W transitions have been seen (rt.py:354-357).  The window is recomputed from the
device-resident history ring every tick (stateless w.r.t. the bank, so the bank may be replaced at any time).

``LookbackLaunch`` wraps the device-resident launch underneath (``llampc_lookback_launch``: many vehicles, rolling rings,
NVLink min-loc across GPUs); ``LookBack`` is the per-tick object a controller loop holds.
"""
import ctypes as C
import os

import numpy as np

from .. import _lib
from ..bank import ModelBank
from .. import dist as _dist


def decode_keys(keys):
    """packed keys (uint64 ndarray) -> (fp32 scores, candidate indices)."""
    keys = np.asarray(keys, dtype=np.uint64)
    err = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
    idx = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    return err, idx


_SINE = {None: _lib.SIN_AUTO, "auto": _lib.SIN_AUTO, "sfu": _lib.SIN_SFU, "strict": _lib.SIN_STRICT,
         True: _lib.SIN_SFU, False: _lib.SIN_STRICT}
_KERNEL = {None: 0, "auto": 0, "k1": _lib.KERNEL_K1, "k1p": _lib.KERNEL_K1P, "k1b": _lib.KERNEL_K1B,
           "k1r": _lib.KERNEL_K1R, "k1v": _lib.KERNEL_K1V, "k1e": _lib.KERNEL_K1E}


def sine_mode(fast_sin):
    """``fast_sin`` argument of the host classes -> LLAMPC_SIN_*: None / "auto" lets the library pick MUFU.SIN while the
    bank's tyre-sine argument stays within [-pi, pi] (else the polynomial); True / "sfu" and False / "strict" force a
    mode; the environment variable LLAMPC_FAST_SIN=1 / 0 overrides the automatic choice (experiments)."""
    if fast_sin is None or fast_sin == "auto":
        env = os.environ.get("LLAMPC_FAST_SIN")
        if env is not None:
            return _lib.SIN_SFU if env == "1" else _lib.SIN_STRICT
    return _SINE[fast_sin]


class LookbackLaunch:
    """One ``llampc_lookback_desc_t`` plus the workspace it needs: the device-resident scoring + selection launch
    (``llampc_lookback_launch``) for any layout -- one history or many vehicles, recompute or rolling, one GPU or the
    NVLink min-loc across several.  ``plan`` holds the library's choices (kernel, window split, sine mode, launches).

    bank      ModelBank;  hist: float32 CUDA tensor [n_vehicles][stride_rows][20] (or [W][20])
    out       int64 CUDA tensor [n_vehicles][17] (allocated when K > 0 and not given)
    """

    def __init__(self, bank, hist, W, Ts, K=10, n_vehicles=1, hist_stride_rows=None, idx_offset=0, mode="recompute",
                 err_ring=None, avg_err=None, out=None, fast_sin=None, kernel=None, split=0, peer=None, pdl=False, wide=False):
        torch = _lib.require_cuda()
        self.torch, self.bank, self.hist, self.peer = torch, bank, hist, peer
        self._L = _lib.lib()
        d = _lib.LookbackDesc()
        d.bank, d.N, d.Npad, d.idx_offset = bank.packed.data_ptr(), bank.N, bank.Npad, int(idx_offset)
        d.geom_shared, d.sin_arg_max, d.sine = int(bank.geom_shared), float(bank.sin_arg_max), sine_mode(fast_sin)
        d.hist, d.W, d.n_vehicles = (hist.data_ptr() if hist is not None else None), int(W), int(n_vehicles)
        d.hist_stride_rows, d.Ts = int(W if hist_stride_rows is None else hist_stride_rows), float(Ts)
        d.mode = _lib.LB_ROLLING if mode == "rolling" else _lib.LB_RECOMPUTE
        d.emit = 1
        self.err_ring, self.avg_err = err_ring, avg_err
        d.err_ring = err_ring.data_ptr() if err_ring is not None else None
        d.avg_err = avg_err.data_ptr() if avg_err is not None else None
        d.K = int(K)
        if K > 0 and out is None:
            out = torch.zeros((n_vehicles, _lib.LIST_LEN + 1), dtype=torch.int64, device=bank.device)
        self.out = out
        d.out = out.data_ptr() if out is not None else None
        d.kernel, d.split = _KERNEL[kernel] if not isinstance(kernel, int) else kernel, int(split)
        d.flags = (_lib.LB_FLAG_PDL if pdl else 0) | (_lib.LB_FLAG_WIDE if wide else 0)
        if peer is not None:
            d.peer_bufs, d.world, d.rank, d.seq = peer.peer_ptrs.data_ptr(), peer.world, peer.rank, 1
        self.desc = d
        self.plan = _lib.LookbackPlan()
        with torch.cuda.device(bank.device):
            if mode == "rolling":
                d.slot = 0
            _lib.check(self._L.llampc_lookback_plan(C.byref(d), C.byref(self.plan)), "llampc_lookback_plan")
        nbytes = int(self.plan.workspace_bytes)
        self.workspace = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=bank.device)
        d.workspace, d.workspace_bytes = self.workspace.data_ptr(), nbytes
        d.peer_bufs = peer.peer_ptrs.data_ptr() if peer is not None else None
        self._ref = C.byref(d)

    @property
    def kernel_name(self):
        return _lib.KERNEL_NAMES[self.plan.kernel]

    @property
    def sine_name(self):
        return _lib.SIN_NAMES[self.plan.sine]

    def launch(self, stream=None, slot=None, emit=None, row32_h=None, use_peer=True):
        """Enqueue the launch on `stream` (raw cudaStream_t; default: torch's current stream)."""
        d = self.desc
        if slot is not None:
            d.slot = int(slot)
        if emit is not None:
            d.emit = int(emit)
        d.row32_h = row32_h
        if self.peer is not None:
            if use_peer:
                d.peer_bufs, d.seq = self.peer.peer_ptrs.data_ptr(), self.peer.next_seq()
            else:
                d.peer_bufs = None
        st = self.torch.cuda.current_stream().cuda_stream if stream is None else stream
        rc = self._L.llampc_lookback_launch(self._ref, st)
        if rc:
            _lib.check(rc, "llampc_lookback_launch")

    def keys(self):
        """out as uint64 ndarray [n_vehicles][17] (synchronises)."""
        return self.out.cpu().numpy().view(np.uint64)


class LookBack:
    """bank_params: dict of the 14 ``Dynamic`` parameters (scalar or (N,) arrays) or a ``ModelBank``.

    W        look-back window length (LookBack_W, rt.py:67)
    K        number of best candidates returned (smoothing_mu_over_mod = 10, rt.py:69,360)
    refine   re-score the max(K, refine) best fp32 candidates in fp64 on the device and order them by the fp64 score:
             ordering and scores among the fp32 finalists are then exact in the reference's arithmetic (membership in
             the finalist set is decided by the fp32 scores; 0 = fp32 scores only).  max(K, refine) <= 16 finishes the
             top-K inside the scoring launch (merge tree); larger values add the stand-alone top-K kernel.
    fast_sin tyre sine: None / "auto" (default) = MUFU.SIN (SFU) while the bank's tyre-sine argument |C| pi/2 stays
             within [-pi, pi], where MUFU.SIN keeps its rated 2^-21.4 absolute error, else the FMA-pipe polynomial
             (~17 % slower, range-independent).  True / False force the SFU / polynomial
             (LLAMPC_FAST_SIN=1 / 0 does the same for every object).  ``sine_name`` reports what runs.
    mode     "recompute" (default): every tick re-integrates the whole W-row window from the history ring (N*W RK4
             steps, stateless w.r.t. the bank); "rolling": the reference's own bookkeeping (rt.py:352-354) -- only the
             newest transition is integrated and its error column replaces the oldest one in a device-resident
             (W, N) ring, the window mean is re-summed (N steps + N*W*4 bytes per tick)
    kernel / split   overrides of the library's kernel choice (None / 0 = automatic; "k1", "k1p", "k1b")
    idx_offset / group   multi-GPU: this rank's bank is the slice starting at global index idx_offset;
             `group` is a torch.distributed process group (None = single GPU)
    """

    def __init__(self, bank_params, W, Ts=0.02, K=10, refine=16, device=None, idx_offset=0, group=None, split=0,
                 mode="recompute", fast_sin=None, kernel=None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.bank = bank_params if isinstance(bank_params, ModelBank) else ModelBank(bank_params, device)
        dev = self.bank.device
        if not (1 <= W <= _lib.MAX_W):
            raise ValueError("W must be in [1, %d]" % _lib.MAX_W)
        self.W, self.Ts, self.K, self.n_refine = int(W), float(Ts), int(K), int(refine)
        self.Kt = max(self.K, self.n_refine)
        if self.Kt > _lib.MAX_K:
            raise ValueError("max(K, refine) must be <= %d" % _lib.MAX_K)
        self.sine = sine_mode(fast_sin)
        self.idx_offset, self.group, self.split = int(idx_offset), group, int(split)
        self.kernel = _KERNEL[kernel] if not isinstance(kernel, int) else kernel
        if mode not in ("recompute", "rolling"):
            raise ValueError("mode must be 'recompute' or 'rolling'")
        self.rolling = mode == "rolling"
        N = self.bank.N
        L = _lib.lib()
        self.hist = torch.zeros((self.W, _lib.HIST_ROW), dtype=torch.float32, device=dev)
        self.hist64 = torch.zeros((self.W, _lib.HIST64_ROW), dtype=torch.float64, device=dev)
        self.avg_err = torch.empty(N, dtype=torch.float32, device=dev)
        self.fused = self.Kt <= _lib.LIST_LEN
        if self.rolling and not self.fused:
            raise ValueError("rolling mode needs max(K, refine) <= %d" % _lib.LIST_LEN)
        self.err_ring = (torch.zeros((_lib.ring_rows(self.W), self.bank.Npad), dtype=torch.float32, device=dev)
                         if self.rolling else None)
        # result: best key | Kt finalist keys | Kt fp64 scores  (the scoring launch writes LIST_LEN + 1 words)
        words = max(2 + 2 * self.Kt, _lib.LIST_LEN + 2)           # + 1 spare word: zero-copy sequence flag
        self._peer = None
        if group is not None and self.n_refine > 0 and os.environ.get("LLAMPC_PEER_GATHER", "1") == "1":
            import torch.distributed as td
            if td.get_backend(group) == "nccl":
                try:                                             # NVLink peer-memory finalist gather (no NCCL per tick)
                    w = td.get_world_size(group)
                    self._peer = _dist.PeerExchange(group, dev, words=2 * w * 4 * self.Kt)
                    words = max(words, 2 + 2 * self.Kt * w)
                except Exception:                                # symmetric memory unavailable: NCCL all-gather path
                    self._peer = None
                agree = torch.tensor([1 if self._peer is not None else 0], device=dev)
                td.all_reduce(agree, op=td.ReduceOp.MIN, group=group)      # every rank must take the same path
                if int(agree.item()) == 0:
                    self._peer = None
                    words = max(2 + 2 * self.Kt, _lib.LIST_LEN + 2)
        self.result = torch.zeros(words, dtype=torch.int64, device=dev)
        self.result_h = torch.zeros(words, dtype=torch.int64, pin_memory=True)
        self._res_keys = self.result_h.numpy().view(np.uint64)
        self._res_errs = self.result_h.numpy().view(np.float64)
        self.rows32_h = np.zeros((self.W, _lib.HIST_ROW), dtype=np.float32)
        self.rows64_h = np.zeros((self.W, _lib.HIST64_ROW), dtype=np.float64)
        self._r32_base, self._r64_base = self.rows32_h.ctypes.data, self.rows64_h.ctypes.data
        self._xk, self._uk, self._xk1 = np.zeros(6), np.zeros(2), np.zeros(6)
        self._xk_p, self._uk_p, self._xk1_p = self._xk.ctypes.data, self._uk.ctypes.data, self._xk1.ctypes.data
        self.window_count = 0
        self._next_slot = 0
        self._async_pending, self._async_out = False, None
        self._geom = (self.bank.lf_shared, self.bank.lr_shared)
        t = _lib.Tick()
        t.bank, t.N, t.Npad = self.bank.packed.data_ptr(), N, self.bank.Npad
        t.hist, t.W, t.Ts = self.hist.data_ptr(), self.W, self.Ts
        t.geom_shared, t.sin_arg_max, t.sine = int(self.bank.geom_shared), float(self.bank.sin_arg_max), self.sine
        t.kernel, t.split, t.idx_offset = self.kernel, self.split, self.idx_offset
        t.avg_err = self.avg_err.data_ptr()
        t.K, t.n_refine = self.K, self.n_refine
        if self.n_refine > 0:
            t.bank64, t.hist64 = self.bank.bank64.data_ptr(), self.hist64.data_ptr()
        t.result, t.result_h = self.result.data_ptr(), self.result_h.data_ptr()
        t.sync = 1
        t.zero_copy = int(os.environ.get("LLAMPC_ZERO_COPY", "1") == "1")
        if self._peer is not None:
            t.zero_copy = 1
            t.peer_bufs, t.peer_world, t.peer_rank = self._peer.peer_ptrs.data_ptr(), self._peer.world, self._peer.rank
        if self.rolling:
            t.err_ring, t.rolling = self.err_ring.data_ptr(), 1
        # one flag per ring slot: "measured at low speed or in a drift" -- the tick runs the wide form of K1p while a tenth of
        # the window is flagged (include/llampc_b200.h, LLAMPC_LB_FLAG_WIDE)
        self._hard = np.zeros(self.W, dtype=np.uint8)
        t.hard_h, t.n_hard = self._hard.ctypes.data, 0
        with torch.cuda.device(dev):
            nbytes = int(L.llampc_lookback_tick_workspace_bytes(C.byref(t)))
        if nbytes < 0:
            _lib.check(nbytes if nbytes > -1000 else -1000 - nbytes, "llampc_lookback_tick_workspace_bytes")
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        t.workspace, t.workspace_bytes = self.workspace.data_ptr(), nbytes
        self._tick = t
        self._tick_ref = C.byref(t)
        self._L = L
        self._stream_dev = torch.cuda.device(dev)
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._idx_out = np.zeros(max(self.Kt, 1), dtype=np.int64)
        self._score_out = np.zeros(max(self.Kt, 1), dtype=np.float64)
        self._nvalid = C.c_int(0)
        self._idx_out_p, self._score_out_p = self._idx_out.ctypes.data, self._score_out.ctypes.data
        self._nvalid_p = C.addressof(self._nvalid)
        self._lf_shared, self._lr_shared = float(self.bank.lf_shared), float(self.bank.lr_shared)
        get_dev = getattr(torch._C, "_cuda_getDevice", None)
        get_raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        self._get_device = get_dev if get_dev is not None else torch.cuda.current_device
        self._get_raw_stream = get_raw if get_raw is not None else (lambda i: torch.cuda.current_stream(i).cuda_stream)

    def plan(self):
        """The library's choices for this object's scoring launch: dict(kernel, split, sine, launches, grid)."""
        d = _lib.LookbackDesc()
        t = self._tick
        d.bank, d.N, d.Npad, d.idx_offset = t.bank, t.N, t.Npad, t.idx_offset
        d.geom_shared, d.sin_arg_max, d.sine = t.geom_shared, t.sin_arg_max, t.sine
        d.hist, d.W, d.n_vehicles, d.hist_stride_rows, d.Ts = t.hist, t.W, 1, t.W, t.Ts
        d.mode = _lib.LB_ROLLING if self.rolling else _lib.LB_RECOMPUTE
        d.emit, d.err_ring, d.row32_h = 1, t.err_ring, (self._r32_base if self.rolling else None)
        d.K, d.avg_err, d.out = (min(max(self.Kt, 1), _lib.LIST_LEN)), t.avg_err, t.result
        d.kernel, d.split = t.kernel, t.split
        p = _lib.LookbackPlan()
        with self._stream_dev:
            _lib.check(self._L.llampc_lookback_plan(C.byref(d), C.byref(p)), "llampc_lookback_plan")
        return {"kernel": _lib.KERNEL_NAMES[p.kernel], "split": p.split, "sine": _lib.SIN_NAMES[p.sine],
                "launches": p.launches, "grid": (p.grid_x, p.grid_y), "block": p.block}

    @property
    def sine_name(self):
        return self.plan()["sine"]

    @property
    def fast_sin(self):
        """True when the scoring launch runs the SFU (MUFU.SIN) tyre sine."""
        return self.plan()["sine"] == _lib.SIN_NAMES[_lib.SIN_SFU]

    def __del__(self):
        try:                                                     # frees the CUDA graph the C tick attached to the struct
            self._L.llampc_lookback_tick_release(self._tick_ref)
        except Exception:
            pass

    # ------------------------------------------------------------------ history ring
    def _pack_row(self, slot, x_k, u_k, x_k1):
        self._xk[:] = x_k
        self._uk[:] = u_k
        self._xk1[:4] = np.asarray(x_k1)[:4]
        r32 = self._r32_base + slot * (_lib.HIST_ROW * 4)
        r64 = self._r64_base + slot * (_lib.HIST64_ROW * 8)
        rc = self._L.llampc_hist_row_pack_h(self._xk_p, self._uk_p, self._xk1_p, self.Ts, self.bank.lf_shared,
                                            self.bank.lr_shared, r32, r64)
        if rc:
            _lib.check(rc, "llampc_hist_row_pack_h")
        return r32, r64

    def load_window(self, x_k, u_k, x_k1):
        """Replace the whole ring by W transitions: x_k (W,6), u_k (W,2), x_k1 (W,>=4) (oldest first)."""
        if self.rolling:                                         # the (W, N) error ring would stay empty
            raise _lib.LlampcError("load_window() only fills the history ring; in mode='rolling' push the W transitions "
                                   "(or use mode='recompute')")
        x_k, u_k, x_k1 = np.asarray(x_k), np.asarray(u_k), np.asarray(x_k1)
        if x_k.shape[0] != self.W:
            raise ValueError("load_window needs exactly W transitions")
        for j in range(self.W):
            self._pack_row(j, x_k[j], u_k[j], x_k1[j])
        torch = self.torch
        self.hist.copy_(torch.from_numpy(self.rows32_h))
        self.hist64.copy_(torch.from_numpy(self.rows64_h))
        self.window_count, self._next_slot = self.W, 0
        vx, vy, w = (np.abs(self.rows32_h[:, i]) for i in (6, 7, 8))
        self._hard[:] = (vx < 0.6) | (vy + 0.06 * w > 0.4 * vx)
        self._tick.n_hard = int(self._hard.sum())

    # ------------------------------------------------------------------ per-tick API
    def push(self, x_k, u_k, x_k1):
        """One MPC tick: (x_k, u_k) -> measured x_k1.  Returns (best_idx, topk_idx, best_err); all None
        while fewer than W transitions have been pushed (rt.py:357)."""
        slot = self._next_slot
        self._next_slot = (slot + 1) % self.W
        self.window_count = min(self.window_count + 1, self.W)
        t = self._tick
        filling = self.window_count < self.W
        if filling or (self.group is not None and self._peer is None):
            r32, r64 = self._pack_row(slot, x_k, u_k, x_k1)
            if filling:
                torch = self.torch
                self.hist[slot].copy_(torch.from_numpy(self.rows32_h[slot]))
                if not self.rolling:                             # the tick is not called while the window fills: flag the row here
                    vx, vy, w = (abs(float(v)) for v in self.rows32_h[slot, 6:9])
                    hard = int(vx < 0.6 or vy + 0.06 * w > 0.4 * vx)
                    t.n_hard += hard - int(self._hard[slot])
                    self._hard[slot] = hard
                if self.rolling:                                 # store the error column, no decision yet
                    t.row32_h, t.row64_h, t.slot, t.rolling = r32, (r64 if self.n_refine > 0 else None), slot, 2
                    with self._stream_dev:
                        rc = self._L.llampc_lookback_tick(self._tick_ref, torch.cuda.current_stream().cuda_stream)
                    t.rolling = 1
                    if rc:
                        _lib.check(rc, "llampc_lookback_tick")
                    torch.cuda.current_stream().synchronize()    # row64_h is reused by the next push
                else:
                    self.hist64[slot].copy_(torch.from_numpy(self.rows64_h[slot]))
                return None, None, None
            t.row32_h, t.row64_h, t.slot = r32, (r64 if self.n_refine > 0 else None), slot
            return self._run_tick()
        # single-GPU steady state: one FFI crossing (row packing, tick, decode all happen in C)
        self._xk[:] = x_k
        self._uk[:] = u_k
        self._xk1[:4] = x_k1[:4]
        t.row32_h = self._r32_base + slot * (_lib.HIST_ROW * 4)
        t.row64_h = self._r64_base + slot * (_lib.HIST64_ROW * 8)     # scratch even without re-score
        t.slot = slot
        if self._peer is not None:
            t.peer_seq = self._peer.next_seq()                   # same count on every rank: one per decided tick
        # raw accessors: torch.cuda.current_device() / current_stream() cost several microseconds of a ~75 us tick
        if self._get_device() != self._dev_index:
            self.torch.cuda.set_device(self._dev_index)
        rc = self._L.llampc_lookback_push(self._tick_ref, self._xk_p, self._uk_p, self._xk1_p, self._lf_shared,
                                          self._lr_shared, self._idx_out_p, self._score_out_p, self._nvalid_p,
                                          self._get_raw_stream(self._dev_index))
        if rc:
            _lib.check(rc, "llampc_lookback_push")
        n = self._nvalid.value
        if n == 0:                                               # every score is NaN (np.argmin would return the first NaN)
            return None, np.zeros(0, dtype=np.int64), float("nan")
        idx = self._idx_out[:n]
        return int(idx[0]), idx[:self.K].copy(), float(self._score_out[0])

    def push_async(self, x_k, u_k, x_k1):
        """Enqueue one tick and return immediately; fetch the decision later with ``collect()``.  The host is free in
        between (in the reference loop: the next tick's planner + NMPC solve), so the look-back latency leaves the
        critical path.  Single GPU or NVLink peer exchange only; while the window fills it degrades to ``push``."""
        if self.group is not None and self._peer is None:
            raise _lib.LlampcError("push_async needs the peer-memory exchange (or a single GPU)")
        self._async_out = None
        if self.window_count + 1 < self.W:
            self._async_out = self.push(x_k, u_k, x_k1)
            return
        slot = self._next_slot
        self._next_slot = (slot + 1) % self.W
        self.window_count = min(self.window_count + 1, self.W)
        r32, r64 = self._pack_row(slot, x_k, u_k, x_k1)
        t = self._tick
        t.row32_h, t.row64_h, t.slot = r32, r64, slot
        if self._peer is not None:
            t.peer_seq = self._peer.next_seq()
        t.sync = 0
        torch = self.torch
        with self._stream_dev:
            rc = self._L.llampc_lookback_tick(self._tick_ref, torch.cuda.current_stream().cuda_stream)
        t.sync = 1
        if rc:
            _lib.check(rc, "llampc_lookback_tick")
        self._async_pending = True

    def collect(self):
        """Decision of the tick enqueued by ``push_async``: (best_idx, topk_idx, best_err)."""
        if not getattr(self, "_async_pending", False):
            return self._async_out if self._async_out is not None else (None, None, None)
        self._async_pending = False
        torch = self.torch
        with self._stream_dev:
            rc = self._L.llampc_lookback_finish(self._tick_ref, torch.cuda.current_stream().cuda_stream)
        if rc:
            _lib.check(rc, "llampc_lookback_finish")
        rc = self._L.llampc_lookback_decode(self._tick_ref, self._idx_out_p, self._score_out_p, self._nvalid_p)
        if rc:
            _lib.check(rc, "llampc_lookback_decode")
        n = self._nvalid.value
        if n == 0:
            return None, np.zeros(0, dtype=np.int64), float("nan")
        idx = self._idx_out[:n]
        return int(idx[0]), idx[:self.K].copy(), float(self._score_out[0])

    def replay(self, x_k, u_k, x_k1, depth=4):
        """Push T consecutive transitions -- x_k (T,6), u_k (T,2), x_k1 (T,>=4), oldest first, e.g. a recorded run -- and return
        the list of T decisions, each what ``push`` would have returned.  The ticks are PIPELINED: up to ``depth`` of them are
        in flight (every tick still carries its own row to the device in the launch parameters and hands its own result back
        through its own mapped pinned slot), so the host packs and enqueues tick t + 1 .. t + depth while the GPU works on
        tick t.  Same decisions as T calls of ``push`` (the look-back's inputs are measurements: no tick depends on the result
        of an earlier one).  Falls back to a loop over ``push`` where the zero-copy hand-off is not in use."""
        x_k, u_k, x_k1 = np.asarray(x_k, dtype=np.float64), np.asarray(u_k, dtype=np.float64), np.asarray(x_k1, dtype=np.float64)
        T = x_k.shape[0]
        out = []
        t = self._tick
        piped = bool(t.zero_copy) and self.n_refine > 0 and (self.group is None or self._peer is not None) and depth > 1
        i = 0
        while i < T and (self.window_count + 1 < self.W or not piped):
            out.append(self.push(x_k[i], u_k[i], x_k1[i]))
            i += 1
        if i >= T:
            return out
        torch = self.torch
        if len(getattr(self, "_slots", ())) < depth:
            self._slots = [torch.zeros(self.result_h.numel(), dtype=torch.int64, pin_memory=True) for _ in range(depth)]
        slot_ptrs = (C.c_void_p * depth)(*[s_.data_ptr() for s_ in self._slots[:depth]])
        n = T - i
        xs = np.ascontiguousarray(x_k[i:, :6])
        us = np.ascontiguousarray(u_k[i:, :2])
        x1 = np.ascontiguousarray(x_k1[i:])
        kt = max(self.Kt, 1)
        idx = np.zeros((n, kt), dtype=np.int64)
        sc = np.zeros((n, kt), dtype=np.float64)
        nv = np.zeros(n, dtype=np.int32)
        seq = C.c_uint(self._peer.seq) if self._peer is not None else None
        first = self._next_slot
        with self._stream_dev:
            rc = self._L.llampc_lookback_replay(self._tick_ref, xs.ctypes.data, us.ctypes.data, x1.ctypes.data, x1.shape[1], n,
                                                first, self._lf_shared, self._lr_shared, self._r32_base, self._r64_base,
                                                C.cast(slot_ptrs, C.c_void_p), depth,
                                                C.addressof(seq) if seq is not None else None, idx.ctypes.data, sc.ctypes.data,
                                                nv.ctypes.data, torch.cuda.current_stream().cuda_stream)
        if seq is not None:
            self._peer.seq = int(seq.value)
        self._next_slot = (first + n) % self.W
        self.window_count = self.W
        if rc:
            _lib.check(rc, "llampc_lookback_replay")
        for r in range(n):
            m = int(nv[r])
            if m == 0:
                out.append((None, np.zeros(0, dtype=np.int64), float("nan")))
            else:
                out.append((int(idx[r, 0]), idx[r, :min(self.K, m)].copy(), float(sc[r, 0])))
        return out

    def evaluate(self):
        """Score the window currently in the ring (after load_window); same return as push."""
        if self.rolling:
            raise _lib.LlampcError("evaluate()/load_window() re-integrate the window: use mode='recompute'")
        if self.window_count < self.W:
            return None, None, None
        t = self._tick
        t.row32_h, t.row64_h, t.slot = None, None, 0
        return self._run_tick()

    def _run_tick(self):
        torch = self.torch
        Kt = self.Kt
        dist = self.group is not None and self._peer is None     # NCCL exchange after the tick
        self._tick.sync = 0 if dist else 1
        if self._peer is not None:                               # NVLink exchange inside the tick's last kernel
            self._tick.peer_seq = self._peer.next_seq()
        with self._stream_dev:
            rc = self._L.llampc_lookback_tick(self._tick_ref, torch.cuda.current_stream().cuda_stream)
        if rc:
            _lib.check(rc, "llampc_lookback_tick")
        if dist:
            return self._finish_distributed()
        keys = self._res_keys
        if Kt == 0:
            err, idx = decode_keys(keys[:1])
            return int(idx[0]), np.zeros(0, dtype=np.int64), float(err[0])
        # finalists arrive ordered by score (fp64 when re-scored), ties by index, NaN / padding last
        idx = (keys[1:1 + Kt] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        scores = self._res_errs[1 + Kt:1 + 2 * Kt]
        n_ok = int(np.count_nonzero(scores == scores))
        return int(idx[0]), idx[:min(self.K, n_ok)].copy(), float(scores[0])

    def _finish_distributed(self):
        """Multi-GPU exchange, on the device: one MIN all-reduce of the packed key (K = 0) or one all-gather of every
        rank's finalists (keys + fp64 scores), then a single copy to the host."""
        import torch.distributed as td
        torch, Kt = self.torch, self.Kt
        if Kt == 0:
            k = _dist.minloc_allreduce(self.result[:1], self.group)
            err, idx = decode_keys(np.array([k], dtype=np.uint64))
            return int(idx[0]), np.zeros(0, dtype=np.int64), float(err[0])
        world = td.get_world_size(self.group)
        if getattr(self, "_gather_buf", None) is None:
            self._gather_buf = torch.empty((world, 2 * Kt), dtype=torch.int64, device=self.bank.device)
        td.all_gather_into_tensor(self._gather_buf, self.result[1:1 + 2 * Kt], group=self.group)
        allr = self._gather_buf.cpu().numpy()                      # synchronises the stream
        keys = allr[:, :Kt].reshape(-1).view(np.uint64)
        idx = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
        if self.n_refine > 0:
            scores = allr[:, Kt:].reshape(-1).view(np.float64)
        else:
            scores = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32).astype(np.float64)
            scores[keys == np.uint64(0xFFFFFFFFFFFFFFFF)] = np.nan
        order = np.lexsort((idx, scores))                          # NaN sorts last
        scores, idx = scores[order], idx[order]
        n_ok = int(np.count_nonzero(scores == scores))
        return int(idx[0]), idx[:min(self.K, n_ok)].copy(), float(scores[0])

    # ------------------------------------------------------------------ bank replacement
    def set_bank(self, bank_params):
        """Swap the model bank (same number of candidates), e.g. after ``ModelBank.generate`` re-centred it on the
        current best model.  In recompute mode the next tick simply scores the new bank over the window already in the
        history ring; in rolling mode the per-tick error ring belongs to the old bank, so the window refills."""
        bank = bank_params if isinstance(bank_params, ModelBank) else ModelBank(bank_params, self.bank.device)
        if bank.N != self.bank.N or bank.device != self.bank.device:
            raise ValueError("set_bank needs a bank of the same size on the same device")
        self.bank = bank
        t = self._tick
        t.bank, t.geom_shared, t.sin_arg_max = bank.packed.data_ptr(), int(bank.geom_shared), float(bank.sin_arg_max)
        if self.n_refine > 0:
            t.bank64 = bank.bank64.data_ptr()
        if self.rolling:
            self.err_ring.zero_()
            self.window_count = 0
        elif self.window_count:
            # the history rows carry stage-1 slip angles computed with the OLD bank's lf, lr: re-pack them
            raise_if = bank.geom_shared and (bank.lf_shared, bank.lr_shared) != (self._geom[0], self._geom[1])
            if raise_if or not bank.geom_shared:
                self.window_count = 0                            # geometry changed: start a fresh window
        self._geom = (bank.lf_shared, bank.lr_shared)
        self._lf_shared, self._lr_shared = float(bank.lf_shared), float(bank.lr_shared)

    # ------------------------------------------------------------------ inspection
    def avg_errors(self):
        """(N,) fp32 window-mean errors of the last tick (avg_errors of rt.py:357) as float64 ndarray."""
        return self.avg_err.cpu().numpy().astype(np.float64)

    def best_key_value(self):
        """Packed fp32 key (float_bits(avg_err) << 32 | index) of the block arg-min of the last tick."""
        return int(self._res_keys[0])
