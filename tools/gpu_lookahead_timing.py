"""GPU diagnostic: look-ahead kernel (config C3: 16,384 x 32 x 20) timing and accuracy for the loaded library."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.mpc import LookAhead
from oracle import llampc_oracle as orc
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
M, K, H, t0 = 16384, 32, 20, 600
rng = np.random.RandomState(3)
Useq = U[:, t0:t0 + H].T[None] + np.stack([0.1 * rng.randn(K, H), 0.05 * rng.randn(K, H)], axis=-1)
Useq[..., 0] = np.clip(Useq[..., 0], -0.1, 1.0); Useq[..., 1] = np.clip(Useq[..., 1], -0.35, 0.35)
xref = S[:2, t0:t0 + H + 1]
bank = orc.make_bank(M, seed=2)  # raw bank (includes spinning candidates); bench.py uses adapted models
la = LookAhead(bank, Ts=Ts)
plan = la.plan(S[:, t0], Useq, xref, U[:, t0 - 1])
for _ in range(3): plan.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): plan.run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
J, bk = plan.fetch()
sub = {k: (bank[k][:256] if np.ndim(bank[k]) else bank[k]) for k in orc.PARAM_NAMES}
Jr, bkr = orc.lookahead_rollout(sub, S[:, t0], Useq, xref, U[:, t0 - 1], Ts)
print("%s: %.1f us  %.3e steps/s  max rel J err %.2e  best_k agree %d/256" % (
    os.environ.get("LLAMPC_LIB", "default"), ms * 1e3, M * K * H / ms * 1e3, np.max(np.abs(J[:256] - Jr) / Jr), int((bk[:256] == bkr).sum())))

rel = np.abs(J[:256] - Jr) / Jr
pm = rel.max(axis=1)
planG = la.plan(S[:, t0], Useq, xref, U[:, t0 - 1], _force_general=True)
planG.run()
JG, _ = planG.fetch()
relG = (np.abs(JG[:256] - Jr) / Jr).max(axis=1)
print("general path: max rel %.2e; fast path: max rel %.2e; p50 %.2e p99 %.2e" % (relG.max(), pm.max(), np.median(pm), np.percentile(pm, 99)))
for m in np.argsort(-pm)[:6]:
    print("model %3d rel fast %.2e general %.2e  Bf %.2f Cf %.2f Df %.3f Br %.2f Cr %.2f Dr %.3f  J range %.3f..%.3f" % (
        m, pm[m], relG[m], bank["Bf"][m], bank["Cf"][m], bank["Df"][m], bank["Br"][m], bank["Cr"][m], bank["Dr"][m], Jr[m].min(), Jr[m].max()))
# fp32 noise amplification in the oracle itself: perturb x0 by 1e-7 relative
Jp, _ = orc.lookahead_rollout(sub, S[:, t0] * (1 + 1e-7), Useq, xref, U[:, t0 - 1], Ts)
print("oracle sensitivity: max rel change of J for a 1e-7 relative perturbation of x0: %.2e (worst model %d)" % (
    (np.abs(Jp - Jr) / Jr).max(), int(np.argmax((np.abs(Jp - Jr) / Jr).max(axis=1)))))
