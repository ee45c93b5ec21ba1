"""GPU probe: does an event record between two back-to-back launches break programmatic dependent launch?  Per-tick time of
the C2 tick (one bank copy per tick) with pdl on, without / with a CUDA event recorded on the stream between the launches."""
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
S, U = bench.synthetic_history(W + 8, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
bank = ModelBank(bench.make_bank(bench.N_C2, seed=1))
rows = np.zeros((W, 20), dtype=np.float32)
for j in range(W):
    xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, j], U[:, j], S[:, j + 1]))
    L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, bench.TS, bank.lf_shared, bank.lr_shared, rows[j].ctypes.data, None)
hist = torch.from_numpy(rows).cuda()
T = 100
copies = [bank.packed.clone() for _ in range(T + 5)]
side = torch.cuda.Stream()
for pdl in (False, True):
    ll = LookbackLaunch(bank, hist, W, bench.TS, K=10, pdl=pdl)
    for mode in ("plain", "event record between launches", "event record + side stream waiting on it"):
        evs = [torch.cuda.Event() for _ in range(T + 5)]
        for i in range(5):
            ll.desc.bank = copies[i].data_ptr(); ll.launch()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(T):
            ll.desc.bank = copies[5 + i].data_ptr()
            ll.launch()
            if mode != "plain":
                evs[i].record()
                if mode.endswith("waiting on it"):
                    side.wait_event(evs[i])
        b.record()
        torch.cuda.synchronize()
        print("pdl=%-5s %-42s %6.2f us per tick" % (pdl, mode, a.elapsed_time(b) * 1e3 / T))
