"""Golden replay of the reference's friction-estimate / look-back block, made by EXECUTING THE REFERENCE'S OWN LINES.

    python tests/golden/make_golden_mu.py        (build container only: needs /root/reference)

The block is inline, module-level code of llampc/mpc/run_nmpc_orca_llampc_rt.py (the script cannot be imported: it
builds IPOPT solvers at import time), so this generator reads the script's source AT GENERATION TIME, slices

    :103-113   class ExponentialSmoother
    :278-282   the planner call (which friction value / speed scale ConstantSpeed receives, and from which tick on)
    :326-344   friction estimate (warm-up seeds, 20-tick moving average MU_pred, smoothed x 0.95 display value MU_preds)
    :346-366   look-back (evaluate_models_vectorized, error window, arg-min, top-10)

and exec()s them, tick by tick, over the recorded closed-loop dataset with a seeded bank built by the reference's
``Dynamic``.  Nothing of the reference's source is written to the repository: only the per-tick outputs are saved to
``mu_replay.npz`` (MU_pred, MU_preds, current_model_idx, ind_best_KM, and the (curr_mu, scale) the planner was called
with).  ``ConstantSpeed`` is replaced by a recorder with the reference's own signature defaults (planner.py:12).
"""
import copy
import inspect
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_adapter as ra  # noqa: E402
from oracle import llampc_oracle as orc  # noqa: E402

ref = ra.load()
SCRIPT = os.path.join(ref.root, "llampc/mpc/run_nmpc_orca_llampc_rt.py")
with open(SCRIPT) as f:
    SRC = f.read().split("\n")


def lines(lo, hi, must_contain):
    """source lines lo..hi (1-based, inclusive), dedented; the anchors guard against a shifted file"""
    chunk = "\n".join(SRC[lo - 1:hi])
    for m in must_contain:
        assert m in chunk, "reference script changed: %r not in lines %d-%d" % (m, lo, hi)
    return textwrap.dedent(chunk)


SMOOTHER = lines(103, 113, ["class ExponentialSmoother", "self.alpha * new_value"])
PLANNER = lines(278, 282, ["if idt > LookBack_W+1:", "curr_mu=MU_pred, scale=v_factor"])
MU_BLOCK = lines(326, 344, ["if idt <= LookBack_W:", "9.8 * params['lr']", "MU_preds.append(smoother.update(MU_pred)*.95)"])
LOOKBACK = lines(346, 366, ["if idt > 0:", "evaluate_models_vectorized(", "ind_best_KM = avg_errors.argsort()"])


def replay(N_MODELS, LookBack_W, seed, t0, n_ticks, smoothing_mu=20, smoothing_mu_over_mod=10, mu_init=1.0,
           mu_alpha=0.08, v_factor=0.9):
    d = np.load(os.path.join(HERE, "ethz_history.npz"))
    S, U, Ts = d["states"], d["inputs"], float(d["Ts"])
    rng = np.random.RandomState(seed)
    params = ref.ORCA(control='pwm')
    MODEL_BANK = []
    for _ in range(N_MODELS):                     # the construction of rt.py:162-179 from a seeded generator
        pv = params.copy()
        for name, sigma in orc.RT_VARIATION:
            pv[name] *= (1 + sigma * rng.randn())
        MODEL_BANK.append(ref.Dynamic(**pv))
    params_pass = tuple(np.array([getattr(m, k) for m in MODEL_BANK]) for k in ("Bf", "Cf", "Df", "Br", "Cr", "Dr"))
    sig = inspect.signature(ref.ConstantSpeed)
    calls = []

    def ConstantSpeed(**kw):                      # records what the planner would have received (planner.py:12 defaults)
        calls.append((kw.get("curr_mu", sig.parameters["curr_mu"].default), kw.get("scale", sig.parameters["scale"].default)))
        return None, kw["projidx"], 0.0

    ns = dict(np=np, copy=copy, params=params, model=ref.Dynamic(**params), mu_init=mu_init, LookBack_W=LookBack_W,
              smoothing_mu=smoothing_mu, smoothing_mu_over_mod=smoothing_mu_over_mod, v_factor=v_factor,
              MODEL_BANK=MODEL_BANK, N_MODELS=N_MODELS, params_pass=params_pass, Ts=Ts,
              evaluate_models_vectorized=ref.evaluate_models_vectorized, ConstantSpeed=ConstantSpeed,
              states=S[:, t0:], inputs=U[:, t0:], error_windows=np.zeros((N_MODELS, LookBack_W)), window_count=0,
              Drs_preds=[], Dfs_preds=[], MUs=[], MU_preds=[], model_switches=[], model_mses=[], chosen_models=[],
              current_model_idx=0, projidx=0, track=None, horizon=20)
    exec(SMOOTHER, ns)
    ns["smoother"] = ns["ExponentialSmoother"](alpha=mu_alpha)
    out = {k: [] for k in ("MU_pred", "MU_preds", "current_model_idx", "ind_best_KM", "planner_mu", "planner_scale",
                           "Drs_preds", "Dfs_preds")}
    for idt in range(n_ticks):
        ns["idt"] = idt
        ns["x0"] = ns["states"][:, idt]
        exec(PLANNER, ns)
        exec(MU_BLOCK, ns)
        exec(LOOKBACK, ns)
        out["MU_pred"].append(ns.get("MU_pred", np.nan))
        out["MU_preds"].append(ns["MU_preds"][-1])
        out["current_model_idx"].append(int(ns["current_model_idx"]))
        out["ind_best_KM"].append(np.asarray(ns["ind_best_KM"]) if "ind_best_KM" in ns
                                  else np.full(smoothing_mu_over_mod, -1))
        out["planner_mu"].append(calls[-1][0])
        out["planner_scale"].append(calls[-1][1])
        out["Drs_preds"].append(ns["Drs_preds"][-1])
        out["Dfs_preds"].append(ns["Dfs_preds"][-1])
    res = {k: np.asarray(v) for k, v in out.items()}
    res.update(N_MODELS=N_MODELS, LookBack_W=LookBack_W, seed=seed, t0=t0, n_ticks=n_ticks, smoothing_mu=smoothing_mu,
               smoothing_mu_over_mod=smoothing_mu_over_mod, mu_init=mu_init, mu_alpha=mu_alpha, v_factor=v_factor,
               Df_bank=params_pass[2], Dr_bank=params_pass[5], mass=params["mass"], lf=params["lf"], lr=params["lr"])
    return res


if __name__ == "__main__":
    # the script's own settings (rt.py:66-72: W = 10, 20-tick moving average, top-10) on a 512-model bank, 160 ticks of
    # the recorded dataset from tick 300; and a W = 25 case where the W + 1 seeds outnumber the 20-tick average
    a = replay(N_MODELS=512, LookBack_W=10, seed=11, t0=300, n_ticks=160)
    b = replay(N_MODELS=256, LookBack_W=25, seed=12, t0=900, n_ticks=90)
    merged = {"a_" + k: v for k, v in a.items()}
    merged.update({"b_" + k: v for k, v in b.items()})
    np.savez_compressed(os.path.join(HERE, "mu_replay.npz"), **merged)
    for tag, r in (("a", a), ("b", b)):
        W = int(r["LookBack_W"])
        print(tag, "W", W, "first MU_pred tick", int(np.argmax(~np.isnan(r["MU_pred"]))),
              "first planner tick with MU_pred", int(np.argmax(r["planner_scale"] != 1.0)),
              "MU_pred[W+1..W+4]", r["MU_pred"][W + 1:W + 5], "MU_preds", r["MU_preds"][W + 1:W + 5])
