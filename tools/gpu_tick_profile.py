"""GPU diagnostic: 12 LookBack.push ticks in recompute and in rolling mode (C2 shape) -- run under
`ncu --metrics gpu__time_duration.sum` to get the per-kernel device times of one e2e tick."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.mpc import LookBack
from oracle import llampc_oracle as orc
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
bank = orc.make_bank(65536, 1, variation=orc.RT_VARIATION + (("mass", 0.15),))
for mode in ("recompute", "rolling"):
    lb = LookBack(bank, W=50, Ts=Ts, K=10, refine=16, mode=mode)
    for t in range(500, 550):
        lb.push(S[:, t], U[:, t], S[:, t + 1])
    torch.cuda.synchronize()
    lat = []
    for t in range(550, 562):
        a = time.perf_counter()
        lb.push(S[:, t], U[:, t], S[:, t + 1])
        lat.append(time.perf_counter() - a)
    print(mode, "push p50 %.1f us" % (np.median(lat) * 1e6))

# ---- device time of each stage of a tick, warm, CUDA events (not under ncu)
from llampc_b200 import _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
lb = LookBack(bank, W=50, Ts=Ts, K=10, refine=16, mode="recompute")
ts = np.arange(500, 550)
lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
lb.evaluate()
lbr = LookBack(bank, W=50, Ts=Ts, K=10, refine=16, mode="rolling")
for t in range(500, 551):
    lbr.push(S[:, t], U[:, t], S[:, t + 1])
row = np.ascontiguousarray(lbr.rows32_h[0])
n_lists = L.llampc_lookback_num_lists(65536, 50, 0)
ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
stages = {
    "K1 (lists + argmin)": lambda: L.llampc_lookback_window_f32(lb.bank.packed.data_ptr(), 65536, lb.bank.Npad, lb.hist.data_ptr(), 50, 1, 50, Ts, lb.avg_err.data_ptr(), lb.best_key.data_ptr(), lb.cta_lists.data_ptr(), 0, 1, lb.split, st),
    "K1 + in-kernel merge K=16": lambda: L.llampc_lookback_window_topk_f32(lb.bank.packed.data_ptr(), 65536, lb.bank.Npad, lb.hist.data_ptr(), 50, 1, 50, Ts, lb.avg_err.data_ptr(), lb.best_key.data_ptr(), lb.cta_lists.data_ptr(), 0, 1, lb.split, 16, ticket.data_ptr(), lb.result.data_ptr(), st),
    "K1 + in-kernel merge K=1": lambda: L.llampc_lookback_window_topk_f32(lb.bank.packed.data_ptr(), 65536, lb.bank.Npad, lb.hist.data_ptr(), 50, 1, 50, Ts, lb.avg_err.data_ptr(), lb.best_key.data_ptr(), lb.cta_lists.data_ptr(), 0, 1, lb.split, 1, ticket.data_ptr(), lb.result.data_ptr(), st),
    "K1 + in-kernel merge K=8": lambda: L.llampc_lookback_window_topk_f32(lb.bank.packed.data_ptr(), 65536, lb.bank.Npad, lb.hist.data_ptr(), 50, 1, 50, Ts, lb.avg_err.data_ptr(), lb.best_key.data_ptr(), lb.cta_lists.data_ptr(), 0, 1, lb.split, 8, ticket.data_ptr(), lb.result.data_ptr(), st),
    "K1 no avg_err no lists": lambda: L.llampc_lookback_window_f32(lb.bank.packed.data_ptr(), 65536, lb.bank.Npad, lb.hist.data_ptr(), 50, 1, 50, Ts, None, lb.best_key.data_ptr(), None, 0, 1, lb.split, st),
    "merge kernel 1024 lists K=16": lambda: L.llampc_topk_merge_lists(lb.cta_lists.data_ptr(), n_lists, 1, 16, lb.best_key.data_ptr(), lb.result.data_ptr(), st),
    "refine 16 finalists": lambda: L.llampc_refine_f64(lb.bank.bank64.data_ptr(), 65536, lb.hist64.data_ptr(), 50, Ts, lb.result[1:].data_ptr(), 16, 0, lb.result[17:].data_ptr(), st),
    "K1r rolling": lambda: L.llampc_lookback_rolling_f32(lbr.bank.packed.data_ptr(), 65536, lbr.bank.Npad, row.ctypes.data, 3, 50, Ts, lbr.err_ring.data_ptr(), lbr.avg_err.data_ptr(), lbr.best_key.data_ptr(), lbr.cta_lists.data_ptr(), 0, 1, 1, st),
    "merge kernel 512 lists K=16": lambda: L.llampc_topk_merge_lists(lbr.cta_lists.data_ptr(), 512, 1, 16, lbr.best_key.data_ptr(), lbr.result.data_ptr(), st),
    "D2H 264 B + sync": lambda: (lb.result_h.copy_(lb.result, non_blocking=True), torch.cuda.current_stream().synchronize()),
}
for name, fn in stages.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(50):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("%-32s device %.2f us   host-side %.2f us per call" % (name, a.elapsed_time(b) * 20, (time.perf_counter() - t0) * 2e4))
# host-side pieces of push()
t0 = time.perf_counter()
for i in range(200):
    lb._pack_row(0, S[:, 600], U[:, 600], S[:, 601])
print("host row packing %.2f us" % ((time.perf_counter() - t0) * 5e3))
