"""GPU diagnostic (trace build, LLAMPC_LIB=...libllampc_b200_trace.so): warp-steps of the C2 tick that take the guard fallback,
for windows ending at different ticks of the synthetic history (K1e counts them per warp)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from llampc_b200 import _lib
from llampc_b200.bank import ModelBank
from llampc_b200.mpc.lookback import LookbackLaunch
from oracle import llampc_oracle as orc
L = _lib.lib()
dbg = C.CDLL(_lib.LIB_PATH).llampc_debug_k1e_trace
W = bench.W_C2
S, U = bench.synthetic_history(520, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
bank = ModelBank(bench.make_bank(bench.N_C2, seed=1))
for t_end in (250, 350, 400, 450):
    rows = np.zeros((W, 20), dtype=np.float32)
    for j in range(W):
        t = t_end - W + j
        xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, t], U[:, t], S[:, t + 1]))
        L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, bench.TS, bank.lf_shared, bank.lr_shared, rows[j].ctypes.data, None)
    ll = LookbackLaunch(bank, torch.from_numpy(rows).cuda(), W, bench.TS, K=10, kernel="k1e")
    ll.launch(); torch.cuda.synchronize()
    nw = ll.plan.grid_x * ll.plan.block // 32
    buf = np.zeros((nw, 6), dtype=np.uint64)
    assert dbg(buf.ctypes.data, nw) == 0
    fb = buf[:, 5].astype(np.int64).sum()
    print("window ending at tick %d: %d of %d warp-steps took the fallback" % (t_end, fb, (bench.N_C2 // 64) * W))
