"""GPU diagnostic: device time of the C2 tick for windows ending at different ticks of the synthetic history (is the
kernel's speed data dependent?), with the fraction of candidate-rows that leave the straight-line step (guard fallback)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from llampc_b200 import _lib
from llampc_b200.bank import ModelBank
from llampc_b200.mpc.lookback import LookbackLaunch
from oracle import llampc_oracle as orc
L = _lib.lib()
W = bench.W_C2
S, U = bench.synthetic_history(520, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
bank = ModelBank(bench.make_bank(bench.N_C2, seed=1))
bh = bench.make_bank(bench.N_C2, seed=1)
for t_end in (50, 150, 250, 300, 350, 400, 450, 500):
    rows = np.zeros((W, 20), dtype=np.float32)
    for j in range(W):
        t = t_end - W + j
        xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, t], U[:, t], S[:, t + 1]))
        L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, bench.TS, bank.lf_shared, bank.lr_shared, rows[j].ctypes.data, None)
    ll = LookbackLaunch(bank, torch.from_numpy(rows).cuda(), W, bench.TS, K=10)
    for _ in range(3):
        ll.launch()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ll.launch()
    b.record()
    torch.cuda.synchronize()
    st = S[:, t_end - W:t_end]
    print("window ending at tick %3d: %6.1f us per launch | vx %.2f..%.2f vy %.2f..%.2f w %.2f..%.2f" % (
        t_end, a.elapsed_time(b) * 1e3 / 20, st[3].min(), st[3].max(), st[4].min(), st[4].max(), st[5].min(), st[5].max()))
