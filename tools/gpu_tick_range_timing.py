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
    hist = torch.from_numpy(rows).cuda()
    us, keys = [], []
    for wide in (False, True):
        ll = LookbackLaunch(bank, hist, W, bench.TS, K=10, wide=wide)
        for _ in range(3):
            ll.launch()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            ll.launch()
        b.record()
        torch.cuda.synchronize()
        us.append(a.elapsed_time(b) * 1e3 / 20)
        keys.append(ll.keys()[0, :11].copy())
    vxr, vyr, wr = (np.abs(rows[:, i]) for i in (6, 7, 8))
    n_hard = int(((vxr < 0.6) | (vyr + 0.06 * wr > 0.4 * vxr)).sum())
    same = bool(((keys[0] & np.uint64(0xFFFFFFFF)) == (keys[1] & np.uint64(0xFFFFFFFF))).all())
    st = S[:, t_end - W:t_end]
    print("window ending at tick %3d: default form %6.1f us, wide form %6.1f us per launch | %2d of %d rows flagged | same arg-min / top-10: %s | vx %.2f..%.2f" % (
        t_end, us[0], us[1], n_hard, W, same, st[3].min(), st[3].max()))
