"""Look-back adaptation step on the GPU.

The reference has no function for this: the logic is inline in the experiment scripts
(run_nmpc_orca_llampc_rt.py:347-366; same block in ..._nrt.py:421-441, ..._nrt_avg_runs.py:446-466).
``LookBack.push`` is that block: it takes the transition measured at one MPC tick, scores every candidate
model of the bank over the last W transitions and returns the arg-min and the K best candidates.

Window semantics (= the reference): W independent one-step RK4 predictions, each re-anchored at the measured
state; no decision until W transitions have been seen (rt.py:354-357).  The window is recomputed from the
device-resident history ring every tick (stateless w.r.t. the bank, so the bank may be replaced at any time).
"""
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
    idx_offset / group   multi-GPU: this rank's bank is the slice starting at global index idx_offset;
             `group` is a torch.distributed process group (None = single GPU)
    """

    def __init__(self, bank_params, W, Ts=0.02, K=10, refine=16, device=None, idx_offset=0, group=None, split=0):
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
        self.idx_offset, self.group, self.split = int(idx_offset), group, int(split)
        N = self.bank.N
        L = _lib.lib()
        self.hist = torch.zeros((self.W, _lib.HIST_ROW), dtype=torch.float32, device=dev)
        self.hist64 = torch.zeros((self.W, _lib.HIST64_ROW), dtype=torch.float64, device=dev)
        self.avg_err = torch.empty(N, dtype=torch.float32, device=dev)
        self.best_key = torch.empty(1, dtype=torch.int64, device=dev)
        ctas = L.llampc_topk_scratch_ctas(N)
        self.topk_scratch = torch.empty(max(1, ctas * max(self.Kt, 1)), dtype=torch.int64, device=dev)
        self.topk_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.topk_keys = torch.empty(1 + max(self.Kt, _lib.LIST_LEN), dtype=torch.int64, device=dev)
        n_lists = L.llampc_lookback_num_lists(N, self.W, self.split)
        self.fused = self.Kt <= _lib.LIST_LEN and 0 < n_lists <= 8192
        self.cta_lists = torch.empty(max(1, n_lists) * _lib.LIST_LEN, dtype=torch.int64, device=dev) if self.fused else None
        self.best_key.fill_(-1)                                   # armed once; the merge kernel re-arms it every tick
        self.refine_err = torch.empty(max(self.Kt, 1), dtype=torch.float64, device=dev)
        self.rows32_h = torch.zeros((self.W, _lib.HIST_ROW), dtype=torch.float32, pin_memory=True)
        self.rows64_h = torch.zeros((self.W, _lib.HIST64_ROW), dtype=torch.float64, pin_memory=True)
        self.out_keys_h = torch.zeros(1 + max(self.Kt, 1), dtype=torch.int64, pin_memory=True)
        self.out_err_h = torch.zeros(max(self.Kt, 1), dtype=torch.float64, pin_memory=True)
        self._out_keys_np = self.out_keys_h.numpy().view(np.uint64)
        self._out_err_np = self.out_err_h.numpy()
        self.window_count = 0
        self._next_slot = 0
        t = _lib.Tick()
        t.bank, t.N, t.Npad = self.bank.packed.data_ptr(), N, self.bank.Npad
        t.hist, t.W, t.Ts = self.hist.data_ptr(), self.W, self.Ts
        t.geom_shared, t.split, t.idx_offset = int(self.bank.geom_shared), self.split, self.idx_offset
        t.avg_err, t.best_key = self.avg_err.data_ptr(), self.best_key.data_ptr()
        t.K, t.n_refine = self.K, self.n_refine
        t.cta_lists = self.cta_lists.data_ptr() if self.fused else None
        t.topk_scratch, t.topk_counter = self.topk_scratch.data_ptr(), self.topk_counter.data_ptr()
        t.topk_keys = self.topk_keys.data_ptr()
        if self.n_refine > 0:
            t.bank64, t.hist64 = self.bank.bank64.data_ptr(), self.hist64.data_ptr()
            t.refine_err64, t.out_err64_h = self.refine_err.data_ptr(), self.out_err_h.data_ptr()
        t.out_keys_h = self.out_keys_h.data_ptr()
        t.sync = 1
        self._tick = t
        self._L = L

    # ------------------------------------------------------------------ history ring
    def _pack_row(self, slot, x_k, u_k, x_k1):
        x_k = np.ascontiguousarray(x_k, dtype=np.float64)
        u_k = np.ascontiguousarray(u_k, dtype=np.float64)
        x_k1 = np.ascontiguousarray(x_k1, dtype=np.float64)
        if x_k.shape != (6,) or u_k.shape != (2,) or x_k1.shape[0] < 4:
            raise ValueError("push expects x_k (6,), u_k (2,), x_k1 (>=4,)")
        r32 = self.rows32_h.data_ptr() + slot * _lib.HIST_ROW * 4
        r64 = self.rows64_h.data_ptr() + slot * _lib.HIST64_ROW * 8
        _lib.check(self._L.llampc_hist_row_pack_h(x_k.ctypes.data, u_k.ctypes.data, x_k1.ctypes.data, self.Ts,
                                                  self.bank.lf_shared, self.bank.lr_shared, r32, r64),
                   "llampc_hist_row_pack_h")
        return r32, r64

    def load_window(self, x_k, u_k, x_k1):
        """Replace the whole ring by W transitions: x_k (W,6), u_k (W,2), x_k1 (W,>=4) (oldest first)."""
        x_k, u_k, x_k1 = np.asarray(x_k), np.asarray(u_k), np.asarray(x_k1)
        if x_k.shape[0] != self.W:
            raise ValueError("load_window needs exactly W transitions")
        for j in range(self.W):
            self._pack_row(j, x_k[j], u_k[j], x_k1[j])
        self.hist.copy_(self.rows32_h, non_blocking=True)
        self.hist64.copy_(self.rows64_h, non_blocking=True)
        self.window_count, self._next_slot = self.W, 0

    # ------------------------------------------------------------------ per-tick API
    def push(self, x_k, u_k, x_k1):
        """One MPC tick: (x_k, u_k) -> measured x_k1.  Returns (best_idx, topk_idx, best_err); all None
        while fewer than W transitions have been pushed (rt.py:357)."""
        slot = self._next_slot
        r32, r64 = self._pack_row(slot, x_k, u_k, x_k1)
        self._next_slot = (slot + 1) % self.W
        self.window_count = min(self.window_count + 1, self.W)
        if self.window_count < self.W:
            self.hist[slot].copy_(self.rows32_h[slot], non_blocking=True)
            self.hist64[slot].copy_(self.rows64_h[slot], non_blocking=True)
            return None, None, None
        t = self._tick
        t.row32_h, t.row64_h, t.slot = r32, (r64 if self.n_refine > 0 else None), slot
        return self._run_tick()

    def evaluate(self):
        """Score the window currently in the ring (after load_window); same return as push."""
        if self.window_count < self.W:
            return None, None, None
        t = self._tick
        t.row32_h, t.row64_h, t.slot = None, None, 0
        return self._run_tick()

    def _run_tick(self):
        torch = self.torch
        with torch.cuda.device(self.bank.device):
            _lib.check(self._L.llampc_lookback_tick(self._tick, _lib.stream_ptr(torch)), "llampc_lookback_tick")
        keys = self._out_keys_np
        if self.Kt == 0:
            err, idx = decode_keys(keys[:1])
            best, topk, best_err = int(idx[0]), idx[:0], float(err[0])
            if self.group is not None:
                k = _dist.minloc_allreduce(self.topk_keys[:1], self.group)
                err, idx = decode_keys(np.array([k], dtype=np.uint64))
                best, best_err = int(idx[0]), float(err[0])
            return best, topk, best_err
        err32, idx = decode_keys(keys[1:1 + self.Kt])
        scores = self._out_err_np[:self.Kt].copy() if self.n_refine > 0 else err32.astype(np.float64)
        if self.group is not None:
            scores, idx = _dist.gather_finalists(scores, idx, self.group, self.bank.device)
        order = np.lexsort((idx, scores))            # by score, ties by index (np.argmin / stable argsort)
        order = order[~np.isnan(scores[order])] if np.isnan(scores).any() else order
        topk = idx[order[:self.K]] if self.K > 0 else idx[:0]
        b = order[0]
        return int(idx[b]), topk, float(scores[b])

    # ------------------------------------------------------------------ inspection
    def avg_errors(self):
        """(N,) fp32 window-mean errors of the last tick (avg_errors of rt.py:357) as float64 ndarray."""
        return self.avg_err.cpu().numpy().astype(np.float64)

    def best_key_value(self):
        return int(np.uint64(self._out_keys_np[0]))
