"""Golden vectors for the planner (ConstantSpeed) made by RUNNING THE REFERENCE (build container only):
    python tests/golden/make_golden_planner.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import reference_adapter as ra  # noqa: E402

ref = ra.load()
Ts, H = 0.02, 20
out = {}
rng = np.random.RandomState(42)
for name, trk in (("ethz", ref.ETHZ(reference='optimal', longer=True)), ("ethzmobil", ref.ETHZMobil(reference='optimal', longer=True))):
    n = trk.raceline.shape[1]
    cases, xrefs, pouts, vrs = [], [], [], []
    mus = np.asarray(trk.mus)
    mu_choices = [mus[0] - 0.1, mus[-1] + 0.2, mus[0], mus[-1], 0.83, float(0.5 * (mus[1] + mus[2]))]
    for c in range(40):
        pid = int(rng.randint(0, n - 12))
        if c in (0, 1):
            pid = n - 12                                        # walk past the end of the arc-length table (modulo)
        j = pid + int(rng.randint(1, 8))
        p = trk.raceline[:, j] + 0.03 * rng.randn(2)            # a point near the raceline inside the search window
        v0 = float(rng.uniform(0.0, 3.0)) if c != 2 else 0.0
        mu = float(mu_choices[c % len(mu_choices)]) if c < 12 else float(rng.uniform(mus[0] - 0.05, mus[-1] + 0.05))
        scale = 0.9 if c % 3 else 1.0
        xref, pout, vr = ref.ConstantSpeed(x0=p, v0=v0, track=trk, N=H, Ts=Ts, projidx=pid, scale=scale, curr_mu=mu)
        cases.append([p[0], p[1], v0, pid, mu, scale])
        xrefs.append(xref); pouts.append(pout); vrs.append(vr)
    out[name + "_cases"] = np.array(cases)
    out[name + "_xref"] = np.array(xrefs)
    out[name + "_projidx"] = np.array(pouts)
    out[name + "_vr"] = np.array(vrs)
    # spline coefficients of the reference objects (x spline and the first / last speed splines)
    out[name + "_sx_b"] = np.array(trk.spline.sx.b); out[name + "_sx_c"] = np.array(trk.spline.sx.c); out[name + "_sx_d"] = np.array(trk.spline.sx.d)
    out[name + "_v0_b"] = np.array(trk.spline_v[0].b); out[name + "_vlast_c"] = np.array(trk.spline_v[-1].c)
np.savez_compressed(os.path.join(HERE, "planner_kat.npz"), Ts=Ts, H=H, **out)
print("written", {k: v.shape for k, v in out.items()})
