"""Look-back adaptation step on the GPU.

The reference has no function for this: the logic is inline in the experiment scripts
(run_nmpc_orca_llampc_rt.py:347-366; same block in ..._nrt.py:421-441, ..._nrt_avg_runs.py:446-466).
``LookBack.push`` is that block: it takes the transition measured at one MPC tick, scores every candidate
model of the bank over the last W transitions and returns the arg-min and the K best candidates.

Window semantics (= the reference): W independent one-step RK4 predictions, each re-anchored at the measured
state; no decision until W transitions have been seen (rt.py:354-357).  The window is recomputed from the
device-resident history ring every tick (stateless w.r.t. the bank, so the bank may be replaced at any time).
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


class LookBack:
    """bank_params: dict of the 14 ``Dynamic`` parameters (scalar or (N,) arrays) or a ``ModelBank``.

    W        look-back window length (LookBack_W, rt.py:67)
    K        number of best candidates returned (smoothing_mu_over_mod = 10, rt.py:69,360)
    refine   re-score the max(K, refine) best fp32 candidates in fp64 on the device and order them by the
             fp64 score: the returned indices and errors are then exact in the reference's arithmetic
             (0 = fp32 scores only).  max(K, refine) <= 16 uses the fused two-launch tick (K1 writes per-CTA
             sorted lists, a K-way merge kernel finishes); larger values use the stand-alone top-K kernel.
    fast_sin evaluate the tyre sine with MUFU.SIN (SFU) instead of the FMA-pipe polynomial: ~17 % faster; measured
             worst per-candidate score error 2.2e-5 (C1) / 8.8e-6 (C2) / 3.8e-5 (sigma = 2 bank) instead of
             8.7e-6 / 9.8e-6 / 6.7e-5 relative -- both inside the 1e-4 tolerance, and the fp64 re-score makes the
             returned indices exact either way.  Default: on (strict mode: fast_sin=False or LLAMPC_FAST_SIN=0).
    mode     "recompute" (default): every tick re-integrates the whole W-row window from the history ring (N*W RK4
             steps, stateless w.r.t. the bank); "rolling": the reference's own bookkeeping (rt.py:352-354) -- only the
             newest transition is integrated and its error column replaces the oldest one in a device-resident
             (W, N) ring, the window mean is re-summed (N steps + N*W*4 bytes per tick)
    balanced recompute mode with max(K, refine) <= 16: run the work-balanced kernel K1b (equal contiguous ranges of
             (candidate group, window row) units over exactly SMs x resident CTAs persistent CTAs, top-K finished by an
             in-kernel tree of warp merges) instead of K1 + list merge.  Default: on (LLAMPC_BALANCED=0 disables).
    idx_offset / group   multi-GPU: this rank's bank is the slice starting at global index idx_offset;
             `group` is a torch.distributed process group (None = single GPU)
    """

    def __init__(self, bank_params, W, Ts=0.02, K=10, refine=16, device=None, idx_offset=0, group=None, split=0,
                 mode="recompute", fast_sin=None, balanced=None):
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
        if fast_sin is None:
            fast_sin = os.environ.get("LLAMPC_FAST_SIN", "1") == "1"
        self.fast_sin = bool(fast_sin)
        if balanced is None:
            balanced = os.environ.get("LLAMPC_BALANCED", "1") == "1"
        # bit 5 (+32) of `split` selects the MUFU.SIN tyre sine in K1 (include/llampc_b200.h)
        self.idx_offset, self.group, self.split = int(idx_offset), group, int(split) | (32 if self.fast_sin else 0)
        if mode not in ("recompute", "rolling"):
            raise ValueError("mode must be 'recompute' or 'rolling'")
        self.rolling = mode == "rolling"
        N = self.bank.N
        L = _lib.lib()
        self.hist = torch.zeros((self.W, _lib.HIST_ROW), dtype=torch.float32, device=dev)
        self.hist64 = torch.zeros((self.W, _lib.HIST64_ROW), dtype=torch.float64, device=dev)
        self.avg_err = torch.empty(N, dtype=torch.float32, device=dev)
        self.best_key = torch.empty(1, dtype=torch.int64, device=dev)
        self.best_key.fill_(-1)                                   # armed once; the merge kernel re-arms it every tick
        n_lists = L.llampc_lookback_num_lists(N, self.W, self.split)
        if self.rolling:
            n_lists = max(n_lists, (N + 127) // 128)
        self.fused = self.Kt <= _lib.LIST_LEN and 0 < n_lists <= 8192
        if self.rolling and not self.fused:
            raise ValueError("rolling mode needs max(K, refine) <= %d and N <= 1,048,576" % _lib.LIST_LEN)
        self.err_ring = torch.zeros((self.W, self.bank.Npad), dtype=torch.float32, device=dev) if self.rolling else None
        self.n_lists = n_lists
        self.cta_lists = torch.empty(max(1, n_lists) * _lib.LIST_LEN, dtype=torch.int64, device=dev) if self.fused else None
        if not self.fused:
            ctas = L.llampc_topk_scratch_ctas(N)
            self.topk_scratch = torch.empty(max(1, ctas * max(self.Kt, 1)), dtype=torch.int64, device=dev)
            self.topk_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        # result: best key | Kt finalist keys | Kt fp64 scores  (the merge kernel writes LIST_LEN + 1 words)
        words = max(2 + 2 * self.Kt, _lib.LIST_LEN + 2)           # + 1 spare word: zero-copy sequence flag
        self._peer = None
        if group is not None and self.n_refine > 0 and os.environ.get("LLAMPC_PEER_GATHER", "1") == "1":
            import torch.distributed as td
            if td.get_backend(group) == "nccl":
                try:                                             # NVLink peer-memory finalist gather (no NCCL per tick)
                    w = td.get_world_size(group)
                    self._peer = _dist.PeerExchange(group, dev, words=2 * w * (2 * self.Kt + 1))
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
        t.geom_shared, t.split, t.idx_offset = int(self.bank.geom_shared), self.split, self.idx_offset
        t.avg_err, t.best_key = self.avg_err.data_ptr(), self.best_key.data_ptr()
        t.K, t.n_refine = self.K, self.n_refine
        t.cta_lists = self.cta_lists.data_ptr() if self.fused else None
        if not self.fused:
            t.topk_scratch, t.topk_counter = self.topk_scratch.data_ptr(), self.topk_counter.data_ptr()
        if self.n_refine > 0:
            t.bank64, t.hist64 = self.bank.bank64.data_ptr(), self.hist64.data_ptr()
        t.result, t.result_h = self.result.data_ptr(), self.result_h.data_ptr()
        t.sync = 1
        self.ticket = torch.zeros(2, dtype=torch.int32, device=dev)
        t.ticket = self.ticket.data_ptr()
        t.zero_copy = int(os.environ.get("LLAMPC_ZERO_COPY", "1") == "1")
        if self._peer is not None:
            t.zero_copy = 1
            t.peer_bufs, t.peer_world, t.peer_rank = self._peer.peer_ptrs.data_ptr(), self._peer.world, self._peer.rank
        if self.rolling:
            t.err_ring, t.rolling = self.err_ring.data_ptr(), 1
        # K1b: work-balanced persistent kernel with the in-kernel tree merge (one launch per tick)
        self.balanced = bool(balanced) and self.fused and not self.rolling and self.Kt > 0
        if self.balanced:
            nbytes = int(L.llampc_lookback_balanced_workspace_bytes(N, self.W))
            if nbytes <= 0:
                _lib.check(nbytes if nbytes > -1000 else -1000 - nbytes, "llampc_lookback_balanced_workspace_bytes")
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
        x_k, u_k, x_k1 = np.asarray(x_k), np.asarray(u_k), np.asarray(x_k1)
        if x_k.shape[0] != self.W:
            raise ValueError("load_window needs exactly W transitions")
        for j in range(self.W):
            self._pack_row(j, x_k[j], u_k[j], x_k1[j])
        torch = self.torch
        self.hist.copy_(torch.from_numpy(self.rows32_h))
        self.hist64.copy_(torch.from_numpy(self.rows64_h))
        self.window_count, self._next_slot = self.W, 0

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
        t.bank, t.geom_shared = bank.packed.data_ptr(), int(bank.geom_shared)
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
