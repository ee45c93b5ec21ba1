"""Minimal ncu target: a few launches of the C2 look-back tick (K1p + merge tree), the C3 look-ahead rollout (K2p), the C4
recompute look-back (K1p over candidate tiles x vehicles), the rolling K1v tick and the planner, each after warm-up.
    ncu --set full --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/prof python tools/gpu_profile_target.py [what ...]
what: c2 c3 c4 k1v planner (default: all)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                     # noqa: E402
from llampc_b200 import _lib                                     # noqa: E402
from llampc_b200.bank import ModelBank                            # noqa: E402
from llampc_b200.mpc import LookAhead                             # noqa: E402
from llampc_b200.mpc.lookback import LookbackLaunch               # noqa: E402
from oracle import llampc_oracle as orc                           # noqa: E402  (synthetic inputs only)

what = set(sys.argv[1:]) or {"c2", "c3", "c4", "k1v", "planner"}
L = _lib.lib()
S, U = bench.synthetic_history(bench.W_C2 + 40, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
TS = bench.TS


def rows_for(bank, W, V=1):
    rows = np.zeros((V, W, 20), dtype=np.float32)
    for j in range(W):
        xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, j], U[:, j], S[:, j + 1]))
        L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, TS, bank.lf_shared, bank.lr_shared,
                                 rows[0, j].ctypes.data, None)
    rows[:] = rows[0][None]
    return torch.from_numpy(rows).cuda()


if "c2" in what:
    bank = ModelBank(bench.make_bank(bench.N_C2, seed=1))
    lb = LookbackLaunch(bank, rows_for(bank, 50)[0], 50, TS, K=10)
    for _ in range(4):
        lb.launch()
    torch.cuda.synchronize()
if "c3" in what:
    M, K, H, t0 = 16384, 32, 20, 50
    rng = np.random.RandomState(3)
    Useq = U[:, t0:t0 + H].T[None] + np.stack([0.1 * rng.randn(K, H), 0.05 * rng.randn(K, H)], axis=-1)
    Useq[..., 0] = np.clip(Useq[..., 0], -0.1, 1.0)
    Useq[..., 1] = np.clip(Useq[..., 1], -0.35, 0.35)
    la = LookAhead(bench.make_bank(M, seed=2), Ts=TS)
    plan = la.plan(S[:, t0], Useq, S[:2, t0:t0 + H + 1], U[:, t0 - 1])
    for _ in range(4):
        plan.run()
    torch.cuda.synchronize()
if "c4" in what or "k1v" in what:
    bank = ModelBank(bench.make_bank_rt(1024, seed=0))
    V = 4096
    hist = rows_for(bank, 20, V)
    if "c4" in what:
        lb4 = LookbackLaunch(bank, hist, 20, TS, K=10, n_vehicles=V)
        for _ in range(3):
            lb4.launch()
        torch.cuda.synchronize()
    if "k1v" in what:
        ring = torch.zeros((V, _lib.ring_rows(20), bank.Npad), dtype=torch.float32, device="cuda")
        lbv = LookbackLaunch(bank, hist, 20, TS, K=10, n_vehicles=V, mode="rolling", err_ring=ring)
        for i in range(24):
            lbv.launch(slot=i % 20, emit=1)
        torch.cuda.synchronize()
if "planner" in what:
    from llampc_b200.tracks import RacelineTable
    rl = np.load(os.path.join(ROOT, "tests", "golden", "raceline_ethzmobil.npz"))
    tab = RacelineTable(rl["x"], rl["y"], rl["speeds"], rl["mus"])
    V = 4096
    r4 = np.random.RandomState(4)
    start = r4.randint(0, 400, V)
    st = np.zeros((V, 6))
    st[:, 0] = 0.6 * rl["x"][start + 1] + 0.4 * rl["x"][start + 2]
    st[:, 1] = 0.6 * rl["y"][start + 1] + 0.4 * rl["y"][start + 2]
    st[:, 3] = 1.0
    for _ in range(4):
        tab.plan(st, start, np.full(V, 0.8), 20, TS, 0.9)
    torch.cuda.synchronize()
print("profile target done:", sorted(what))
