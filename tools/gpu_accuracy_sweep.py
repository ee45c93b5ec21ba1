"""GPU diagnostic: worst-case relative score error of K1 (SFU sine and strict mode) against the fp64 oracle over many
windows of the recorded dataset and two window lengths; also counts arg-min / top-10 disagreements of the fp32 ranking
(before the fp64 re-score)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.mpc import LookBack
from oracle import llampc_oracle as orc
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
N = int(os.environ.get("SWEEP_N", "4096"))      # >= 8192 exercises the packed kernel K1p
bank = orc.make_bank(N, seed=21, variation=orc.RT_VARIATION + (("mass", 0.15),))
for W in (10, 50):
    worst = {True: 0.0, False: 0.0}
    where = {True: None, False: None}
    bad_rank = {True: 0, False: 0}
    lbs = {fs: LookBack(bank, W=W, Ts=Ts, K=10, refine=0, fast_sin=fs) for fs in (True, False)}
    ticks = list(range(W + 5, 1790, 35))
    for t_end in ticks:
        ref = np.mean(orc.window_errors(bank, S, U, t_end, W, Ts), axis=1)
        order = np.argsort(ref, kind="stable")[:10]
        ts = np.arange(t_end - W + 1, t_end + 1)
        for fs, lb in lbs.items():
            lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
            best, topk, _ = lb.evaluate()
            rel = np.abs(lb.avg_errors() - ref) / ref
            if rel.max() > worst[fs]:
                worst[fs], where[fs] = rel.max(), (t_end, int(np.argmax(rel)), ref[np.argmax(rel)], ref.min())
            bad_rank[fs] += int(best != order[0] or list(topk) != list(order))
    for fs in (True, False):
        print("N=%d W=%2d %-28s windows %d  worst rel err %.2e at (tick, cand, score, best score) %s  fp32-ranking mismatches %d" % (
            N, W, "SFU sine (default)" if fs else "strict polynomial", len(ticks), worst[fs], where[fs], bad_rank[fs]), flush=True)
