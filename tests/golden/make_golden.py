"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Everything written here is an output of reference functions on seeded inputs, or reference DATA
(the recorded closed-loop dataset / raceline tables, MIT-licensed, LICENSE.md:1-3); no reference
source code is copied.  The GPU box has no /root/reference, so tests there read these files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_adapter as ra  # noqa: E402
from oracle import llampc_oracle as orc  # noqa: E402

ref = ra.load()
Ts = 0.02


def ref_bank(n, seed, variation=orc.RT_VARIATION):
    """Bank drawn exactly like run_nmpc_orca_llampc_rt.py:162-179 but from a seeded RandomState."""
    rng = np.random.RandomState(seed)
    params = ref.ORCA(control='pwm')
    models = []
    for _ in range(n):
        pv = params.copy()
        for name, sigma in variation:
            pv[name] *= (1 + sigma * rng.randn())
        models.append(ref.Dynamic(**pv))
    pp = tuple(np.array([getattr(m, k) for m in models]) for k in ("Bf", "Cf", "Df", "Br", "Cr", "Dr"))
    return models, pp


# ---------------------------------------------------------------- dataset (reference DATA)
d = np.load(os.path.join(ref.root, "llampc/data/DYN-GPMPC-NOCONS-with_var_speedsETHZ.npz"))
S = d["states"][:6].astype(np.float64)
U = d["inputs"].astype(np.float64)
np.savez_compressed(os.path.join(HERE, "ethz_history.npz"), states=S, inputs=U, Ts=Ts)

# ---------------------------------------------------------------- nominal known-answer vectors
p = ref.ORCA(control='pwm')
m = ref.Dynamic(**p)
x = np.array([0.1, 0.2, 0.3, 1.5, 0.05, 0.4])
u = np.array([0.5, 0.1])
kat = dict(
    x=x, u=u, Ts=Ts,
    f=m._diffequation(None, x, u),
    f_batch=m._diffequation_batch(None, x[None], u[None])[0],
    forces=np.array(m.calc_forces(x, u, return_slip=True)),
    rk4=m._integrate_batch(x[None], u[None], 0, Ts)[0],
    rk6=m._integrate(x, u, 0, Ts),
)
xs, dxs = m.sim_continuous(S[:, 600], U[:, 600:605], np.arange(6) * Ts)
kat.update(sim_x=xs, sim_dxdt=dxs)
np.savez(os.path.join(HERE, "kat_nominal.npz"), **kat)

# ---------------------------------------------------------------- look-back, config C1 (N=1024, W=20)
N, W, K = 1024, 20, 10
models, pp = ref_bank(N, seed=0)
out = dict(params=np.stack(pp), W=W, K=K, Ts=Ts, ticks=np.array([100, 600, 1100, 1600]))
for t_end in out["ticks"]:
    # the rt.py:347-366 block, driven by the recorded states/inputs, from an empty window
    error_windows = np.zeros((N, W))
    window_count = 0
    for idt in range(t_end - W + 1, t_end + 1):
        pred = ref.evaluate_models_vectorized(models, N, S[:, idt], U[:, idt], Ts, pp)
        errors = np.mean((pred - S[0:4, idt + 1]) ** 2, axis=1)
        error_windows = np.roll(error_windows, -1, axis=1)
        error_windows[:, -1] = errors
        window_count = min(window_count + 1, W)
    assert window_count >= W
    avg = np.mean(error_windows, axis=1)
    out["pred_%d" % t_end] = pred                       # one-step predictions of the last tick (N,4)
    out["avg_%d" % t_end] = avg
    out["best_%d" % t_end] = np.argmin(avg)
    out["topk_%d" % t_end] = avg.argsort()[:K]
np.savez_compressed(os.path.join(HERE, "lookback_c1.npz"), **out)

# ---------------------------------------------------------------- all-14-parameter broadcast
rng = np.random.RandomState(7)
Nv = 256
pv = {k: p[k] * (1 + 0.1 * rng.randn(Nv)) for k in orc.PARAM_NAMES}
mv = ref.Dynamic(**pv)
xb = np.tile(S[:, 900], (Nv, 1))
ub = np.tile(U[:, 900], (Nv, 1))
np.savez_compressed(os.path.join(HERE, "vary14.npz"), tick=900, Ts=Ts,
                    **{"p_" + k: pv[k] for k in orc.PARAM_NAMES},
                    rk4=mv._integrate_batch(xb, ub, 0, Ts), f=mv._diffequation_batch(None, xb, ub))

# ---------------------------------------------------------------- look-ahead KAT
track = ref.ETHZ(reference='optimal', longer=True)
t0, H = 600, 20
x0 = S[:, t0]
dist = np.hypot(track.raceline[0] - x0[0], track.raceline[1] - x0[1])
near = int(np.argmin(dist))
xref, projidx_out, vr = ref.ConstantSpeed(x0=x0[:2], v0=x0[3], track=track, N=H, Ts=Ts, projidx=near - 2,
                                          curr_mu=0.83, scale=.9)
M, Kc = 64, 8
models_la, pp_la = ref_bank(M, seed=2)
rng = np.random.RandomState(3)
u_nom = U[:, t0:t0 + H].T                                  # (H,2)
eps = np.stack([0.1 * rng.randn(Kc, H), 0.05 * rng.randn(Kc, H)], axis=-1)
eps[0] = 0.0                                               # sequence 0 = recorded inputs
Useq = u_nom[None] + eps
Useq[..., 0] = np.clip(Useq[..., 0], p["min_pwm"], p["max_pwm"])
Useq[..., 1] = np.clip(Useq[..., 1], p["min_steer"], p["max_steer"])
uprev = U[:, t0 - 1]
bm = ref.Dynamic(**{**{k: p[k] for k in orc.PARAM_NAMES},
                    **dict(zip(("Bf", "Cf", "Df", "Br", "Cr", "Dr"), [np.repeat(a, Kc) for a in pp_la]))})
xb = np.tile(x0, (M * Kc, 1))
Uf = np.broadcast_to(Useq[None], (M, Kc, H, 2)).reshape(M * Kc, H, 2)
Jt = np.zeros(M * Kc)
Ja = np.zeros(M * Kc)
Rw = (5e-3, 1.0)
for h in range(H):
    du = Uf[:, h] - (uprev[None] if h == 0 else Uf[:, h - 1])
    Ja += Rw[0] * du[:, 0] ** 2 + Rw[1] * du[:, 1] ** 2
    xb = bm._integrate_batch(xb, Uf[:, h], 0, Ts)
    e = xb[:, :2] - xref[:, h + 1][None]
    Jt += e[:, 0] ** 2 + e[:, 1] ** 2
# nominal model on the recorded inputs (the SURVEY section 4 KAT numbers)
xn = x0[None].copy()
Jn = 0.0
for h in range(H):
    xn = m._integrate_batch(xn, U[:, t0 + h][None], 0, Ts)
    Jn += float(np.sum((xn[0, :2] - xref[:, h + 1]) ** 2))
np.savez_compressed(os.path.join(HERE, "lookahead_kat.npz"), t0=t0, H=H, Ts=Ts, x0=x0, uprev=uprev, xref=xref,
                    projidx_in=near - 2, projidx_out=projidx_out, vr=vr, params=np.stack(pp_la), U=Useq,
                    J=(Jt + Ja).reshape(M, Kc), x_final=xb.reshape(M, Kc, 6), J_track_nominal=Jn,
                    x_final_nominal=xn[0])

# ---------------------------------------------------------------- raceline tables (reference DATA)
for name, trk in (("ethz", track), ("ethzmobil", ref.ETHZMobil(reference='optimal', longer=True))):
    np.savez_compressed(os.path.join(HERE, "raceline_%s.npz" % name), x=trk.x_raceline, y=trk.y_raceline,
                        speeds=np.asarray(trk.v_raceline), mus=np.asarray(trk.mus), s=np.asarray(trk.spline.s))

print("golden vectors written to", HERE)
print("KAT rk4", kat["rk4"].tolist())
print("lookahead nominal J_track", Jn, "vr", vr, "projidx", near - 2, projidx_out)
for t_end in out["ticks"]:
    print(t_end, int(out["best_%d" % t_end]), out["topk_%d" % t_end].tolist(), float(out["avg_%d" % t_end].min()))
