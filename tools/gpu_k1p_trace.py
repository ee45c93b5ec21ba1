"""GPU diagnostic: per-CTA timeline of the C2 look-back tick (K1p + merge tree), from the trace build of the library
(`make -C llampc_b200/csrc trace`; run with LLAMPC_LIB=llampc_b200/libllampc_b200_trace.so).  Prints where the launch
spends its time: CTA start ramp, history staging, RK4 rows, CTA selection, merge tree, and the idle tail.
    LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_trace.so python tools/gpu_k1p_trace.py [N] [W]"""
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
dbg = C.CDLL(_lib.LIB_PATH).llampc_debug_k1p_trace
S, U = bench.synthetic_history(W + 8, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
bank = ModelBank(bench.make_bank(N, seed=1))
rows = np.zeros((W, 20), dtype=np.float32)
for j in range(W):
    xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, j], U[:, j], S[:, j + 1]))
    L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, bench.TS, bank.lf_shared, bank.lr_shared,
                             rows[j].ctypes.data, None)
hist = torch.from_numpy(rows).cuda()
lb = LookbackLaunch(bank, hist, W, bench.TS, K=10, kernel="k1p")
n_cta = lb.plan.grid_x
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
buf = np.zeros((n_cta, 6), dtype=np.uint64)
agg = []
for rep in range(6):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lb.launch()
    b.record()
    torch.cuda.synchronize()
    assert dbg(buf.ctypes.data, n_cta) == 0
    t = buf[:, :5].astype(np.int64)
    t0 = t[:, 0].min()
    rel = (t - t0) * 1e-3                                        # us since the first CTA started
    sm = buf[:, 5].astype(np.int64)
    if rep < 2:
        continue
    order = np.argsort(rel[:, 0])
    first_wave = rel[:, 0] < 3.0
    agg.append(dict(event_us=a.elapsed_time(b) * 1e3, span=rel[:, 4].max(), start_p50=np.median(rel[first_wave, 0]),
                    start_max_first_wave=rel[first_wave, 0].max(), n_first_wave=int(first_wave.sum()),
                    prologue=np.median(rel[:, 1] - rel[:, 0]), prologue_first=np.median((rel[:, 1] - rel[:, 0])[first_wave]),
                    rows=np.median(rel[:, 2] - rel[:, 1]), rows_first=np.median((rel[:, 2] - rel[:, 1])[first_wave]),
                    rows_second=np.median((rel[:, 2] - rel[:, 1])[~first_wave]) if (~first_wave).any() else 0.0,
                    select=np.median(rel[:, 3] - rel[:, 2]), tree=np.median(rel[:, 4] - rel[:, 3]),
                    tree_max=(rel[:, 4] - rel[:, 3]).max(), last_rows_end=rel[:, 2].max(), last_exit=rel[:, 4].max(),
                    second_wave_start_p50=np.median(rel[~first_wave, 0]) if (~first_wave).any() else 0.0,
                    ctas_per_sm_max=int(np.bincount(sm).max()), ctas_per_sm_min=int(np.bincount(sm, minlength=148).min())))
print("N=%d W=%d: %d CTAs (%s split %d)" % (N, W, n_cta, lb.kernel_name, lb.plan.split))
for k in agg[0]:
    print("  %-24s %s" % (k, "  ".join("%8.2f" % d[k] for d in agg)))
