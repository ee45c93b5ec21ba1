"""GPU diagnostic (run under gpurun): LookBack.push against the bare C call llampc_lookback_push with prebuilt
arguments -- the difference is the Python cost of a tick."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.mpc import LookBack                # noqa: E402
from oracle import llampc_oracle as orc             # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
bank = orc.make_bank(65536, 1, variation=orc.RT_VARIATION + (("mass", 0.15),))
for mode in ("recompute", "rolling"):
    lb = LookBack(bank, W=50, Ts=Ts, K=10, refine=16, mode=mode)
    for t in range(500, 560):
        lb.push(S[:, t], U[:, t], S[:, t + 1])
    lat_py, lat_c = [], []
    for t in range(560, 760):
        a = time.perf_counter()
        lb.push(S[:, t], U[:, t], S[:, t + 1])
        lat_py.append(time.perf_counter() - a)
    L, tk = lb._L, lb._tick
    st = torch.cuda.current_stream().cuda_stream
    for t in range(760, 960):
        slot = lb._next_slot
        lb._next_slot = (slot + 1) % lb.W
        lb._xk[:] = S[:, t]; lb._uk[:] = U[:, t]; lb._xk1[:4] = S[:4, t + 1]
        tk.row32_h = lb._r32_base + slot * 80
        tk.row64_h = lb._r64_base + slot * 96
        tk.slot = slot
        a = time.perf_counter()
        L.llampc_lookback_push(lb._tick_ref, lb._xk_p, lb._uk_p, lb._xk1_p, lb._lf_shared, lb._lr_shared, lb._idx_out_p,
                               lb._score_out_p, lb._nvalid_p, st)
        lat_c.append(time.perf_counter() - a)
    print("%-10s LookBack.push p50 %.1f us   bare llampc_lookback_push p50 %.1f us" % (
        mode, np.median(lat_py) * 1e6, np.median(lat_c) * 1e6), flush=True)
