"""Replay a recorded closed-loop trajectory through the look-back step with the reference script's own settings
(run_nmpc_orca_llampc_rt.py:52-74: N_MODELS = 5000, LookBack_W = 10, top-10 friction estimate, 20-tick smoothing,
alpha = 0.08), i.e. the per-tick body rt.py:326-366 without the IPOPT solve.

    python examples/replay_lookback.py [path/to/history.npz]      # default: tests/golden/ethz_history.npz

Prints the model switches and the friction estimate; `replay()` is also used by the parity test, which runs the
same loop with the NumPy oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_MODELS, LookBack_W, smoothing_mu, smoothing_mu_over_mod, mu_alpha = 5000, 10, 20, 10, 0.08


def replay(S, U, Ts, bank, n_ticks, make_lookback, make_mu):
    """The reference's tick body: returns (current_model_idx per tick, MU_pred per tick)."""
    lookback = make_lookback(bank, LookBack_W, Ts, smoothing_mu_over_mod)
    mu_est = make_mu(bank["mass"] if np.ndim(bank["mass"]) == 0 else bank["mass"][0])
    current_model_idx, MU_pred = -1, 1.0
    idx_hist, mu_hist = [], []
    ind_best = None
    for idt in range(n_ticks):
        if ind_best is not None:                                   # rt.py:326-344 (uses the previous tick's top-K)
            MU_pred = mu_est.update(bank["Dr"][ind_best], bank["Df"][ind_best])
        best, topk, _ = lookback.push(S[:, idt], U[:, idt], S[:, idt + 1])      # rt.py:347-360
        if best is not None:
            ind_best = np.asarray(topk)
            if best != current_model_idx:                          # rt.py:362-366
                current_model_idx = best
        idx_hist.append(current_model_idx)
        mu_hist.append(MU_pred)
    return np.array(idx_hist), np.array(mu_hist)


def main():
    from llampc_b200.bank import make_bank
    from llampc_b200.mpc import LookBack, MuEstimator
    from llampc_b200.params import ORCA
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ethz_history.npz")
    g = np.load(path)
    S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
    bank = make_bank(ORCA(), N_MODELS, rng=np.random.RandomState(0))
    idx, mu = replay(S, U, Ts, bank, U.shape[1] - 1,
                     lambda b, W, ts, K: LookBack(b, W=W, Ts=ts, K=K),
                     lambda m: MuEstimator(mass=m, smoothing_mu=smoothing_mu, alpha=mu_alpha))
    switches = np.flatnonzero(np.diff(idx)) + 1
    print("ticks %d, model switches %d, final model %d (Df %.4f Dr %.4f), final mu estimate %.4f" % (
        len(idx), len(switches), idx[-1], bank["Df"][idx[-1]], bank["Dr"][idx[-1]], mu[-1]))
    for t in switches[:10]:
        print("  tick %4d -> model %4d  mu %.4f" % (t, idx[t], mu[t]))


if __name__ == "__main__":
    main()
