"""GPU diagnostic: LookBack.replay (pipelined ticks, one C loop) against a loop over push, per-tick wall time by depth."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                     # noqa: E402
from llampc_b200.mpc import LookBack                             # noqa: E402
from oracle import llampc_oracle as orc                           # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_C2
W, T = bench.W_C2, 200
S, U = bench.synthetic_history(W + 3 * T + 8, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
bank = bench.make_bank(N, seed=1)
for mode in ("recompute", "rolling"):
    lb = LookBack(bank, W=W, Ts=bench.TS, K=10, refine=16, mode=mode)
    for t in range(W + 5):
        lb.push(S[:, t], U[:, t], S[:, t + 1])
    t0 = time.perf_counter()
    for t in range(W + 5, W + 5 + T):
        lb.push(S[:, t], U[:, t], S[:, t + 1])
    print("%-9s push loop            %6.1f us per tick" % (mode, (time.perf_counter() - t0) / T * 1e6))
    ts = np.arange(W + 5 + T, W + 5 + 2 * T)
    for depth in (1, 2, 3, 4, 8):
        lb.replay(S[:, ts[:8]].T, U[:, ts[:8]].T, S[:, ts[:8] + 1].T, depth=depth)
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        t0 = time.perf_counter()
        lb.replay(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T, depth=depth)
        wall = time.perf_counter() - t0
        eb.record()
        torch.cuda.synchronize()
        print("%-9s replay depth %d       %6.1f us per tick (device span %6.1f us per tick)" % (mode, depth, wall / T * 1e6, ea.elapsed_time(eb) * 1e3 / T))
