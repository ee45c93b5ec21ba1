"""GPU diagnostic: device time of the C3 look-ahead exactly as bench.py builds it (the 16,384 best-adapted models of a
65,536-candidate bank x 32 sequences x 20 steps, L2 flushed), for the library named by LLAMPC_LIB (kernel experiments)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                     # noqa: E402
from llampc_b200.mpc import LookAhead, LookBack                   # noqa: E402
from oracle import llampc_oracle as orc                           # noqa: E402

TS, W = bench.TS, bench.W_C2
S, U = bench.synthetic_history(W + 40, lambda p, x, u: orc.rk6_step(p, x, u, 0, TS))
M, K, H, t0 = 16384, 32, 20, W
rng = np.random.RandomState(3)
Useq = U[:, t0:t0 + H].T[None] + np.stack([0.1 * rng.randn(K, H), 0.05 * rng.randn(K, H)], axis=-1)
Useq[..., 0] = np.clip(Useq[..., 0], -0.1, 1.0)
Useq[..., 1] = np.clip(Useq[..., 1], -0.35, 0.35)
xref = S[:2, t0:t0 + H + 1]
big = bench.make_bank(bench.N_C2, seed=2)
lbm = LookBack(big, W=W, Ts=TS, K=10, refine=0)
ts = np.arange(0, W)
lbm.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
lbm.evaluate()
keep = np.argsort(lbm.avg_errors(), kind="stable")[:M]
bank = {k: (v[keep] if np.ndim(v) else v) for k, v in big.items()}
la = LookAhead(bank, Ts=TS)
plan = la.plan(S[:, t0], Useq, xref, U[:, t0 - 1])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in evs:
    flush.fill_(1)
    a.record()
    plan.run()
    b.record()
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in evs])
J, bk = plan.fetch()
sub = {k: (bank[k][:128] if np.ndim(bank[k]) else bank[k]) for k in orc.PARAM_NAMES}
Jr, bkr = orc.lookahead_rollout(sub, S[:, t0], Useq, xref, U[:, t0 - 1], TS)
Jp, _ = orc.lookahead_rollout(sub, S[:, t0] * (1 + 1e-7), Useq, xref, U[:, t0 - 1], TS)
rel, sens = np.abs(J[:128] - Jr) / Jr, np.abs(Jp - Jr) / Jr
print("%-40s C3: %.1f us mean %.1f us min  %.3e steps/s (%.3f of 1.142e11) | max rel J err %.2e where the oracle's own sensitivity < 1e-5; "
      "worst err/sens elsewhere %.1f" % (os.path.basename(os.environ.get("LLAMPC_LIB", "default")), ms.mean() * 1e3, ms.min() * 1e3,
                                        M * K * H / (ms.mean() * 1e-3), M * K * H / (ms.mean() * 1e-3) / 1.142e11,
                                        rel[sens < 1e-5].max(), (rel[sens >= 1e-5] / sens[sens >= 1e-5]).max() if (sens >= 1e-5).any() else 0.0))
