"""CPU oracle for the LLA-MPC look-back / look-ahead hot path (float64, NumPy).

TEST INFRASTRUCTURE ONLY.  This file is a restatement of the reference's algorithm,
used as the checker in ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under ``llampc_b200/`` (the product)
imports it; the product path fails loudly when the CUDA library is missing.

Parity pinning: the reference (tianhao-stan-wu/LLA-MPC) ships no tests, golden vectors or
known-answer fixtures of its own ("parity unpinned" by the reference's test-suite).  The
oracle is pinned instead by
  * ``tests/golden/*.npz`` -- outputs of the reference's own functions, generated in the build
    container by importing ``/root/reference`` (script: ``tests/golden/make_golden.py``),
  * ``tests/test_oracle_vs_reference.py`` -- bit-for-bit float64 comparison against the imported
    reference whenever ``/root/reference`` exists.

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# llampc/params/orca.py:13-35 -- ORCA 1:43 nominal parameters and input limits
# --------------------------------------------------------------------------------------
PARAM_NAMES = ("lf", "lr", "mass", "Iz", "Bf", "Br", "Cf", "Cr", "Df", "Dr", "Cm1", "Cm2", "Cr0", "Cr2")


def orca_params():
    """llampc/params/orca.py:13-84 (pwm control branch)."""
    return {
        "lf": 0.029, "lr": 0.033, "mass": 0.041, "Iz": 27.8e-6,
        "Bf": 2.579, "Br": 3.3852, "Cf": 1.2, "Cr": 1.2691, "Df": 0.192, "Dr": 0.1737,
        "Cm1": 0.287, "Cm2": 0.0545, "Cr0": 0.0518, "Cr2": 0.00035,
        "max_pwm": 1.0, "min_pwm": -0.1, "max_steer": 0.35, "min_steer": -0.35, "max_steer_vel": 5.0,
    }


# --------------------------------------------------------------------------------------
# llampc/models/dynamic.py:117-154 -- tire / drivetrain forces (pwm + Pacejka branch :141-149)
# --------------------------------------------------------------------------------------
def calc_forces_batch(p, x_batch, u_batch, return_slip=False):
    steer = u_batch[:, 1]
    vx = x_batch[:, 3]
    vy = x_batch[:, 4]
    omega = x_batch[:, 5]
    pwm = u_batch[:, 0]
    Frx = (p["Cm1"] - p["Cm2"] * vx) * pwm - p["Cr0"] - p["Cr2"] * (vx ** 2)
    alphaf = steer - np.arctan2((p["lf"] * omega + vy), np.abs(vx))
    alphar = np.arctan2((p["lr"] * omega - vy), np.abs(vx))
    Ffy = p["Df"] * np.sin(p["Cf"] * np.arctan(p["Bf"] * alphaf))
    Fry = p["Dr"] * np.sin(p["Cr"] * np.arctan(p["Br"] * alphar))
    if return_slip:
        return Ffy, Frx, Fry, alphaf, alphar
    return Ffy, Frx, Fry


# --------------------------------------------------------------------------------------
# llampc/models/dynamic.py:98-115 -- 6-state right-hand side, batched
# --------------------------------------------------------------------------------------
def diffequation_batch(p, x_batch, u_batch):
    psi = x_batch[:, 2]
    vx = x_batch[:, 3]
    vy = x_batch[:, 4]
    omega = x_batch[:, 5]
    Ffy, Frx, Fry = calc_forces_batch(p, x_batch, u_batch)
    return np.stack([
        vx * np.cos(psi) - vy * np.sin(psi),
        vx * np.sin(psi) + vy * np.cos(psi),
        omega,
        1 / p["mass"] * (Frx - Ffy * np.sin(u_batch[:, 1])) + vy * omega,
        1 / p["mass"] * (Fry + Ffy * np.cos(u_batch[:, 1])) - vx * omega,
        1 / p["Iz"] * (Ffy * p["lf"] * np.cos(u_batch[:, 1]) - Fry * p["lr"]),
    ], axis=1)


# --------------------------------------------------------------------------------------
# llampc/utils/rk6.py:50-68 via llampc/models/model.py:32-40 -- one classic RK4 step, batched
# --------------------------------------------------------------------------------------
def rk4_step_batch(p, x_batch, u_batch, t_start, t_end):
    h = t_end - t_start
    k1 = h * diffequation_batch(p, x_batch, u_batch)
    k2 = h * diffequation_batch(p, x_batch + k1 / 2, u_batch)
    k3 = h * diffequation_batch(p, x_batch + k2 / 2, u_batch)
    k4 = h * diffequation_batch(p, x_batch + k3, u_batch)
    return x_batch + (k1 + 2 * k2 + 2 * k3 + k4) / 6


# --------------------------------------------------------------------------------------
# llampc/utils/rk6.py:13-28 via llampc/models/model.py:18-30 -- plant step (6-stage RKF45 weights)
# scalar right-hand side: llampc/models/dynamic.py:76-96,156-193
# --------------------------------------------------------------------------------------
def diffequation(p, x, u):
    x = np.asarray(x, dtype=np.float64)
    psi, vx, vy, omega = x[2], x[3], x[4], x[5]
    pwm, steer = u[0], u[1]
    Frx = (p["Cm1"] - p["Cm2"] * vx) * pwm - p["Cr0"] - p["Cr2"] * (vx ** 2)
    alphaf = steer - np.arctan2((p["lf"] * omega + vy), abs(vx))
    alphar = np.arctan2((p["lr"] * omega - vy), abs(vx))
    Ffy = p["Df"] * np.sin(p["Cf"] * np.arctan(p["Bf"] * alphaf))
    Fry = p["Dr"] * np.sin(p["Cr"] * np.arctan(p["Br"] * alphar))
    dxdt = np.zeros(6)
    dxdt[0] = vx * np.cos(psi) - vy * np.sin(psi)
    dxdt[1] = vx * np.sin(psi) + vy * np.cos(psi)
    dxdt[2] = omega
    dxdt[3] = 1 / p["mass"] * (Frx - Ffy * np.sin(steer)) + vy * omega
    dxdt[4] = 1 / p["mass"] * (Fry + Ffy * np.cos(steer)) - vx * omega
    dxdt[5] = 1 / p["Iz"] * (Ffy * p["lf"] * np.cos(steer) - Fry * p["lr"])
    return dxdt


_RK6_GAMMA = np.asarray([16 / 135, 0, 6656 / 12825, 28561 / 56430, -9 / 50, 2 / 55])


def rk6_step(p, x, u, t_start, t_end):
    """One plant step, llampc/utils/rk6.py:17-27 (the tableau literals are restated verbatim
    because bit-for-bit float64 parity depends on the exact operation order)."""
    y0 = np.asarray(x, dtype=np.float64)
    h = t_end - t_start
    f = lambda y: diffequation(p, y, u)
    k1 = h * f(y0)
    k2 = h * f(y0 + k1 / 4)
    k3 = h * f(y0 + 3 / 32 * k1 + 9 / 32 * k2)
    k4 = h * f(y0 + 1932 / 2197 * k1 - 7200 / 2197 * k2 + 7296 / 2197 * k3)
    k5 = h * f(y0 + 439 / 216 * k1 - 8 * k2 + 3680 / 513 * k3 - 845 / 4104 * k4)
    k6 = h * f(y0 - 8 / 27 * k1 + 2 * k2 - 3544 / 2565 * k3 + 1859 / 4104 * k4 - 11 / 40 * k5)
    K = np.asarray([k1, k2, k3, k4, k5, k6])
    return y0 + _RK6_GAMMA @ K


def sim_continuous(p, x0, u, t):
    """llampc/models/dynamic.py:59-74 -- plant simulation over n steps (RK6)."""
    n_steps = u.shape[1]
    x = np.zeros([6, n_steps + 1])
    dxdt = np.zeros([6, n_steps + 1])
    dxdt[:, 0] = diffequation(p, x0, [0, 0])
    x[:, 0] = x0
    for ids in range(1, n_steps + 1):
        x[:, ids] = rk6_step(p, x[:, ids - 1], u[:, ids - 1], t[ids - 1], t[ids])
        dxdt[:, ids] = diffequation(p, x[:, ids], u[:, ids - 1])
    return x, dxdt


# --------------------------------------------------------------------------------------
# llampc/mpc/evaluate_models_vectorized.py:4-23 -- the drop-in boundary function
# --------------------------------------------------------------------------------------
def evaluate_models_vectorized(shared, n_models, current_state, input_val, Ts, params):
    """`shared` plays the role of ``models[0]`` (:16-19): mass, lf, lr, Iz, Cm1, Cm2, Cr0, Cr2 come
    from it; `params` is the 6-tuple (Bfs, Cfs, Dfs, Brs, Crs, Drs) of (N,) arrays (:8)."""
    Bfs, Cfs, Dfs, Brs, Crs, Drs = params
    n_models = len(Bfs)
    x0_batch = np.tile(current_state, (n_models, 1))
    u_batch = np.tile(input_val, (n_models, 1))
    p = {k: shared[k] for k in ("mass", "lf", "lr", "Iz", "Cm1", "Cm2", "Cr0", "Cr2")}
    p.update(Bf=Bfs, Cf=Cfs, Df=Dfs, Br=Brs, Cr=Crs, Dr=Drs)
    return np.vstack(rk4_step_batch(p, x0_batch, u_batch, 0, Ts))[:, 0:4]


def onestep_predict(bank, state, input_val, Ts):
    """Generalisation of the boundary function: every one of the 14 parameters may be a scalar or an
    (N,) array (verified in SURVEY.md section 8 quirk 5 to broadcast in the reference)."""
    n = bank_size(bank)
    x0_batch = np.tile(np.asarray(state, dtype=np.float64), (n, 1))
    u_batch = np.tile(np.asarray(input_val, dtype=np.float64), (n, 1))
    return rk4_step_batch(bank, x0_batch, u_batch, 0, Ts)


def bank_size(bank):
    n = 1
    for k in PARAM_NAMES:
        v = np.asarray(bank[k])
        if v.ndim == 1:
            n = max(n, v.shape[0])
    return n


# --------------------------------------------------------------------------------------
# llampc/mpc/run_nmpc_orca_llampc_rt.py:145-179 -- model-bank construction
# --------------------------------------------------------------------------------------
RT_VARIATION = (("Br", 0.2), ("Cr", 0.1), ("Dr", 0.5), ("Bf", 0.2), ("Cf", 0.1), ("Df", 0.5))


def make_bank(n_models, seed, variation=RT_VARIATION, nominal=None):
    """Same draw order as the reference loop (for each model, for each entry of variation_dict in
    dict order, one randn), but from a seeded RandomState (the reference uses the unseeded global RNG)."""
    rng = np.random.RandomState(seed)
    nominal = orca_params() if nominal is None else nominal
    z = rng.randn(n_models, len(variation))
    bank = {k: nominal[k] for k in PARAM_NAMES}
    for j, (name, sigma) in enumerate(variation):
        bank[name] = nominal[name] * (1 + sigma * z[:, j])
    return bank


# --------------------------------------------------------------------------------------
# llampc/mpc/run_nmpc_orca_llampc_rt.py:347-366 -- look-back scoring and selection
# --------------------------------------------------------------------------------------
def onestep_errors(bank, x_k, u_k, x_k1, Ts):
    """rt.py:349 -- mean squared one-step prediction error over states x, y, psi, vx."""
    pred = onestep_predict(bank, x_k, u_k, Ts)[:, 0:4]
    return np.mean((pred - np.asarray(x_k1, dtype=np.float64)[0:4]) ** 2, axis=1)


def window_errors(bank, states, inputs, t_end, W, Ts):
    """(N, W) matrix whose column j holds the rt.py:349 errors of tick ``t_end - W + 1 + j``;
    this is what ``error_windows`` contains after tick ``t_end`` (rt.py:352-353)."""
    cols = [onestep_errors(bank, states[:, t], inputs[:, t], states[:, t + 1], Ts)
            for t in range(t_end - W + 1, t_end + 1)]
    return np.stack(cols, axis=1)


def select(avg_errors, K=10):
    """rt.py:357-360 -- argmin (first index on ties) and the K best indices."""
    return int(np.argmin(avg_errors)), avg_errors.argsort()[:K]


class LookBackOracle:
    """Stateful restatement of the per-tick block rt.py:347-366 (np.roll window and all)."""

    def __init__(self, bank, W, Ts, K=10):
        self.bank, self.W, self.Ts, self.K = bank, W, Ts, K
        self.error_windows = np.zeros((bank_size(bank), W))
        self.window_count = 0

    def push(self, x_k, u_k, x_k1):
        errors = onestep_errors(self.bank, x_k, u_k, x_k1, self.Ts)
        self.error_windows = np.roll(self.error_windows, -1, axis=1)
        self.error_windows[:, -1] = errors
        self.window_count = min(self.window_count + 1, self.W)
        if self.window_count >= self.W:
            avg_errors = np.mean(self.error_windows, axis=1)
            best, topk = select(avg_errors, self.K)
            return best, topk, avg_errors
        return None, None, None


# --------------------------------------------------------------------------------------
# llampc/mpc/run_nmpc_orca_llampc_rt.py:326-344 -- friction estimate from the K best models, and which value the
# planner call :278-282 receives.  Pinned against tests/golden/mu_replay.npz (the reference's own lines executed over the
# recorded dataset by tests/golden/make_golden_mu.py).
# --------------------------------------------------------------------------------------
class MuEstimatorOracle:
    """Per-tick restatement.  ``tick(idt, ind_best_KM, Dr_bank, Df_bank)`` is the block :326-344 of tick ``idt``;
    ``MU_pred`` is the raw moving average of :341 (NaN until tick W + 1), ``MU_preds`` the logged list of :330/:344."""

    def __init__(self, mass, lf, lr, W, smoothing_mu=20, alpha=0.08, mu_init=1.0, v_factor=0.9):
        self.mass, self.lf, self.lr, self.W = mass, lf, lr, W
        self.smoothing_mu, self.alpha, self.mu_init, self.v_factor = smoothing_mu, alpha, mu_init, v_factor
        self.Drs_preds, self.Dfs_preds, self.MU_preds = [], [], []
        self.MU_pred = np.nan
        self.smooth_value = None                # ExponentialSmoother state, rt.py:103-113

    def tick(self, idt, ind_best_KM=None, Dr_bank=None, Df_bank=None):
        if idt <= self.W:                                                   # :326-330 (g = 9.8 here)
            self.Drs_preds.append(self.mu_init * self.mass * 9.8 * self.lr / (self.lf + self.lr))
            self.Dfs_preds.append(self.mu_init * self.mass * 9.8 * self.lf / (self.lf + self.lr))
            self.MU_preds.append(self.mu_init)
        else:                                                               # :331-344 (g = 9.81 here)
            bestKDr = [Dr_bank[i] for i in ind_best_KM]
            bestKDf = [Df_bank[i] for i in ind_best_KM]
            self.Drs_preds.append(np.mean(bestKDr))
            self.Dfs_preds.append(np.mean(bestKDf))
            self.MU_pred = (np.mean(np.array(self.Drs_preds)[-self.smoothing_mu:])
                            + np.mean(np.array(self.Dfs_preds)[-self.smoothing_mu:])) / (9.81 * self.mass)
            if self.smooth_value is None:
                self.smooth_value = self.MU_pred
            else:
                self.smooth_value = self.alpha * self.MU_pred + (1 - self.alpha) * self.smooth_value
            self.MU_preds.append(self.smooth_value * .95)
        return self.MU_pred

    def planner_mu_scale(self, idt):
        """(curr_mu, scale) of the ConstantSpeed call of tick idt, rt.py:278-282 (defaults of planner.py:12 before W + 2)."""
        return (self.MU_pred, self.v_factor) if idt > self.W + 1 else (1., 1.)


# --------------------------------------------------------------------------------------
# Look-ahead: Model._integrate_batch chained H times (model.py:32-40) + the NMPC objective
# llampc/mpc/nmpc.py:48,66-71,111 with Q, P, R from run_nmpc_orca_llampc_rt.py:60-62
# --------------------------------------------------------------------------------------
def lookahead_rollout(bank, x0, U, xref, uprev, Ts, Q=(1.0, 1.0), R=(5e-3, 1.0), P=(0.0, 0.0),
                      return_traj=False):
    """bank: M candidate models; U: (K, H, 2) control sequences shared by all models or (M, K, H, 2);
    xref: (2, H+1); returns J (M, K) and per-model argmin over K.

    J = sum_{h=1..H} (p_h-xref_h)^T Q (p_h-xref_h) + (p_H-xref_H)^T P (p_H-xref_H)
        + sum_{h=0..H-1} du_h^T R du_h,   du_0 = u_0 - uprev  (nmpc.py:66-71)."""
    M = bank_size(bank)
    U = np.asarray(U, dtype=np.float64)
    if U.ndim == 3:
        U = np.broadcast_to(U[None], (M,) + U.shape)
    _, K, H, _ = U.shape
    flat = {k: (np.repeat(np.asarray(bank[k], dtype=np.float64), K) if np.ndim(bank[k]) == 1 else bank[k])
            for k in PARAM_NAMES}
    x = np.tile(np.asarray(x0, dtype=np.float64), (M * K, 1))
    Uf = U.reshape(M * K, H, 2)
    J_track = np.zeros(M * K)
    J_act = np.zeros(M * K)
    traj = [x.copy()]
    for h in range(H):
        du = Uf[:, h, :] - (np.asarray(uprev, dtype=np.float64)[None, :] if h == 0 else Uf[:, h - 1, :])
        J_act += R[0] * du[:, 0] ** 2 + R[1] * du[:, 1] ** 2
        x = rk4_step_batch(flat, x, Uf[:, h, :], 0, Ts)
        e = x[:, 0:2] - xref[:, h + 1][None, :]
        J_track += Q[0] * e[:, 0] ** 2 + Q[1] * e[:, 1] ** 2
        if return_traj:
            traj.append(x.copy())
    e = x[:, 0:2] - xref[:, H][None, :]
    J_track += P[0] * e[:, 0] ** 2 + P[1] * e[:, 1] ** 2
    J = (J_track + J_act).reshape(M, K)
    best_k = np.argmin(J, axis=1)
    if return_traj:
        return J, best_k, np.stack(traj, axis=1).reshape(M, K, H + 1, 6)
    return J, best_k
