"""GPU diagnostic: device time of one llampc_lookback_launch configuration (CUDA events, L2 flushed between launches).

    python tools/gpu_launch_timing.py N W V [mode] [kernel] [sine] [reps] [split]
      N candidates, W window rows, V vehicles; mode recompute|rolling; kernel auto|k1|k1p|k1b|k1r|k1v; sine auto|sfu|strict
Prints the library's plan, the mean / min device time and candidate-steps/s (recompute: N*W*V per launch, rolling: N*V)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200 import _lib                                     # noqa: E402
from llampc_b200.bank import ModelBank                            # noqa: E402
from llampc_b200.mpc.lookback import LookbackLaunch               # noqa: E402
from oracle import llampc_oracle as orc                           # noqa: E402  (bank construction only)

N, W, V = (int(a) for a in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "recompute"
kernel = sys.argv[5] if len(sys.argv) > 5 else "auto"
sine = sys.argv[6] if len(sys.argv) > 6 else "auto"
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 30
split = int(sys.argv[8]) if len(sys.argv) > 8 else 0
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
L = _lib.lib()
bank = ModelBank(orc.make_bank(N, seed=0, variation=orc.RT_VARIATION + (("mass", 0.15),)))
rows = np.zeros((V, W, 20), dtype=np.float32)
one = np.zeros((W, 20), dtype=np.float32)
for j in range(W):
    t = 600 + j
    xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, t], U[:, t], S[:, t + 1]))
    L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, Ts, bank.lf_shared, bank.lr_shared,
                             one[j].ctypes.data, None)
rows[:] = one[None]
hist = torch.from_numpy(rows).cuda()
ring = torch.zeros((V, _lib.ring_rows(W), bank.Npad), dtype=torch.float32, device="cuda") if mode == "rolling" else None
lb = LookbackLaunch(bank, hist, W, Ts, K=10, n_vehicles=V, mode=mode, err_ring=ring, kernel=kernel, split=split,
                    fast_sin={"auto": None, "sfu": True, "strict": False}[sine])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(W if mode == "rolling" else 3):
    lb.launch(slot=i % W) if mode == "rolling" else lb.launch()
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for i, (a, b) in enumerate(evs):
    flush.add_(1)
    a.record()
    lb.launch(slot=i % W) if mode == "rolling" else lb.launch()
    b.record()
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in evs])
steps = N * V * (1 if mode == "rolling" else W)
print("N=%d W=%d V=%d %s: kernel %s split %d sine %s grid (%d,%d)x%d launches %d | %.1f us mean, %.1f us min | %.3e steps/s (%.2f of 1.142e11)"
      % (N, W, V, mode, lb.kernel_name, lb.plan.split, lb.sine_name, lb.plan.grid_x, lb.plan.grid_y, lb.plan.block,
         lb.plan.launches, ms.mean() * 1e3, ms.min() * 1e3, steps / (ms.mean() * 1e-3), steps / (ms.mean() * 1e-3) / 1.142e11))
