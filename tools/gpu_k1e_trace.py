"""GPU diagnostic: per-warp timeline of the look-back tick on the equal-share kernel K1e, from the trace build of the
library (`make -C llampc_b200/csrc trace`; run with LLAMPC_LIB=llampc_b200/libllampc_b200_trace.so).
    LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_trace.so python tools/gpu_k1e_trace.py [N] [W]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                     # noqa: E402
from llampc_b200 import _lib                                     # noqa: E402
from llampc_b200.bank import ModelBank                            # noqa: E402
from llampc_b200.mpc.lookback import LookbackLaunch               # noqa: E402
from oracle import llampc_oracle as orc                           # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_C2
W = int(sys.argv[2]) if len(sys.argv) > 2 else bench.W_C2
L = _lib.lib()
dbg = C.CDLL(_lib.LIB_PATH).llampc_debug_k1e_trace
S, U = bench.synthetic_history(W + 8, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
bank = ModelBank(bench.make_bank(N, seed=1))
rows = np.zeros((W, 20), dtype=np.float32)
for j in range(W):
    xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, j], U[:, j], S[:, j + 1]))
    L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, bench.TS, bank.lf_shared, bank.lr_shared,
                             rows[j].ctypes.data, None)
hist = torch.from_numpy(rows).cuda()
lb = LookbackLaunch(bank, hist, W, bench.TS, K=10, kernel="k1e")
WPC = lb.plan.block // 32
n_w = min(lb.plan.grid_x * WPC, 8192)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
buf = np.zeros((n_w, 6), dtype=np.uint64)
print("N=%d W=%d: %d CTAs (%s), %.2f warp-steps per warp" % (N, W, lb.plan.grid_x, lb.kernel_name, (N + 63) // 64 * W / (lb.plan.grid_x * WPC)))
for rep in range(5):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lb.launch()
    b.record()
    torch.cuda.synchronize()
    assert dbg(buf.ctypes.data, n_w) == 0
    if rep < 2:
        continue
    t = buf[:, :4].astype(np.int64)
    if rep == 4:
        np.save(os.path.join(ROOT, "gpurun_out", "k1e_trace_%d_%d.npy" % (N, W)), buf)
    t0 = t[:, 0].min()
    rel = (t - t0) * 1e-3
    w0 = np.arange(n_w) % WPC == 0
    rows_us = rel[:, 2] - rel[:, 1]
    pct = lambda v: "min %.2f p10 %.2f p50 %.2f p90 %.2f max %.2f" % (v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max())
    print("event %.2f us | start %s" % (a.elapsed_time(b) * 1e3, pct(rel[:, 0])))
    print("   prologue %s" % pct(rel[:, 1] - rel[:, 0]))
    print("   rows     %s" % pct(rows_us))
    print("   rows end %s" % pct(rel[:, 2]))
    print("   exit (warp 0 of every CTA) %s" % pct(rel[w0, 3]))
    wi = np.arange(n_w) % WPC
    print("   rows by warp-in-CTA: " + " ".join("%.1f" % rows_us[wi == w].mean() for w in range(WPC)))
