"""Replay a recorded closed-loop trajectory through the look-back step with the reference script's own settings
(run_nmpc_orca_llampc_rt.py:52-74: N_MODELS = 5000, LookBack_W = 10, top-10 friction estimate, 20-tick smoothing,
alpha = 0.08), i.e. the per-tick body rt.py:278-282 + :326-366 without the IPOPT solve.

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


def replay(S, U, Ts, bank, n_ticks, make_lookback, make_mu, W=LookBack_W, K=smoothing_mu_over_mod):
    """The reference's tick body in its own order (rt.py:278-282 planner arguments, :326-344 friction estimate from the
    PREVIOUS tick's top-K, :346-366 look-back from tick 1 on).  `make_mu(mass, lf, lr, W)` returns an object with
    ``planner_args(idt)``, ``tick(idt, best_Dr, best_Df)``, ``mu_display``.  Returns a dict of per-tick arrays."""
    lookback = make_lookback(bank, W, Ts, K)
    scalar = lambda k: float(bank[k] if np.ndim(bank[k]) == 0 else bank[k][0])
    mu_est = make_mu(scalar("mass"), scalar("lf"), scalar("lr"), W)
    current_model_idx, ind_best_KM = 0, None
    out = {k: [] for k in ("current_model_idx", "MU_pred", "MU_preds", "planner_mu", "planner_scale", "ind_best_KM")}
    for idt in range(n_ticks):
        kw = mu_est.planner_args(idt)                               # what ConstantSpeed would receive, rt.py:278-282
        out["planner_mu"].append(kw.get("curr_mu", 1.0))
        out["planner_scale"].append(kw.get("scale", 1.0))
        if idt <= W:                                               # rt.py:326-330
            mu = mu_est.tick(idt)
        else:                                                      # rt.py:331-344
            mu = mu_est.tick(idt, bank["Dr"][ind_best_KM], bank["Df"][ind_best_KM])
        if idt > 0:                                                # rt.py:346: the transition of tick 0 is skipped
            best, topk, _ = lookback.push(S[:, idt], U[:, idt], S[:, idt + 1])      # rt.py:349-360
            if best is not None:
                ind_best_KM = np.asarray(topk)
                if best != current_model_idx:                      # rt.py:362-366
                    current_model_idx = best
        out["current_model_idx"].append(current_model_idx)
        out["MU_pred"].append(np.nan if mu is None else mu)
        out["MU_preds"].append(mu_est.mu_display)
        out["ind_best_KM"].append(np.full(K, -1) if ind_best_KM is None else ind_best_KM.copy())
    return {k: np.asarray(v) for k, v in out.items()}


def main():
    from llampc_b200.bank import make_bank
    from llampc_b200.mpc import LookBack, MuEstimator
    from llampc_b200.params import ORCA
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ethz_history.npz")
    g = np.load(path)
    S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
    bank = make_bank(ORCA(), N_MODELS, rng=np.random.RandomState(0))
    r = replay(S, U, Ts, bank, U.shape[1] - 1,
               lambda b, W, ts, K: LookBack(b, W=W, Ts=ts, K=K),
               lambda m, lf, lr, W: MuEstimator(mass=m, lf=lf, lr=lr, W=W, smoothing_mu=smoothing_mu, alpha=mu_alpha))
    idx, mu = r["current_model_idx"], r["MU_pred"]
    switches = np.flatnonzero(np.diff(idx)) + 1
    print("ticks %d, model switches %d, final model %d (Df %.4f Dr %.4f), final MU_pred %.4f (logged value %.4f)" % (
        len(idx), len(switches), idx[-1], bank["Df"][idx[-1]], bank["Dr"][idx[-1]], mu[-1], r["MU_preds"][-1]))
    for t in switches[:10]:
        print("  tick %4d -> model %4d  MU_pred %.4f" % (t, idx[t], mu[t]))


if __name__ == "__main__":
    main()
