"""GPU diagnostic: where the fixed cost of a one-launch tick comes from.  Event-to-event time of (a) a trivial framework
kernel, (b) the C1 tick (1,024 x 20: ~1.5 us of RK4 rows), (c) the C2 tick, each timed right after a 256 MiB L2 flush (the
bench's protocol), after a small unrelated kernel, and back to back with itself."""
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

L = _lib.lib()


def make(N, W):
    S, U = bench.synthetic_history(W + 8, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
    bank = ModelBank(bench.make_bank(N, seed=1))
    rows = np.zeros((W, 20), dtype=np.float32)
    for j in range(W):
        xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, j], U[:, j], S[:, j + 1]))
        L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, bench.TS, bank.lf_shared, bank.lr_shared,
                                 rows[j].ctypes.data, None)
    return LookbackLaunch(bank, torch.from_numpy(rows).cuda(), W, bench.TS, K=10)


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
small = torch.zeros(1024, device="cuda")


def timed(fn, before, reps=30):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        before()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in evs[3:]]) * 1e3
    return "%6.2f us (min %6.2f)" % (np.median(t), t.min())


c1, c2 = make(1024, 20), make(65536, 50)
cases = [("trivial kernel (1,024-element add)", lambda: small.add_(1.0)), ("C1 tick 1,024 x 20 (%s)" % c1.kernel_name, c1.launch),
         ("C2 tick 65,536 x 50 (%s)" % c2.kernel_name, c2.launch)]
for name, fn in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print("%-40s after L2 flush %s | after a small kernel %s | back to back %s"
          % (name, timed(fn, lambda: flush.fill_(1)), timed(fn, lambda: small.mul_(1.0)), timed(fn, fn)))
