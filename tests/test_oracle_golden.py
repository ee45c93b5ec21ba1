"""The oracle restatement must reproduce the reference's own outputs (tests/golden, made by
tests/golden/make_golden.py from /root/reference) bit-for-bit in float64."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import llampc_oracle as orc


def test_kat_nominal_bit_exact():
    g = load_golden("kat_nominal.npz")
    p = orc.orca_params()
    x, u, Ts = g["x"], g["u"], float(g["Ts"])
    assert np.array_equal(orc.diffequation(p, x, u), g["f"])
    assert np.array_equal(orc.diffequation_batch(p, x[None], u[None])[0], g["f_batch"])
    forces = np.array(orc.calc_forces_batch(p, x[None], u[None], return_slip=True))[:, 0]
    assert np.array_equal(forces, g["forces"])
    assert np.array_equal(orc.rk4_step_batch(p, x[None], u[None], 0, Ts)[0], g["rk4"])
    assert np.array_equal(orc.rk6_step(p, x, u, 0, Ts), g["rk6"])


def test_kat_literals_from_survey():
    """SURVEY.md section 4 literals (computed from the reference by the survey)."""
    p = orc.orca_params()
    x = np.array([0.1, 0.2, 0.3, 1.5, 0.05, 0.4])
    u = np.array([0.5, 0.1])
    f = orc.diffequation(p, x, u)
    assert f.tolist() == [1.4182287233553421, 0.4910471344482896, 0.4, 1.1562355111119065,
                          -0.20549569970320886, 57.52663373383829]
    rk6 = orc.rk6_step(p, x, u, 0, 0.02)
    np.testing.assert_allclose(rk6, [0.12855058624451512, 0.20998681089444135, 0.3173689655134326,
                                     1.5234401271629654, 0.03543879866412555, 1.2425832175917786], rtol=0, atol=1e-16)


def test_plant_sim_continuous(history):
    g = load_golden("kat_nominal.npz")
    S, U, Ts = history
    xs, dxs = orc.sim_continuous(orc.orca_params(), S[:, 600], U[:, 600:605], np.arange(6) * Ts)
    assert np.array_equal(xs, g["sim_x"])
    assert np.array_equal(dxs, g["sim_dxdt"])


def test_dataset_rk4_floor(history):
    """nominal-model one-step MSE at dataset tick 600 = 4.69e-11 (RK4-vs-RK6 floor)."""
    S, U, Ts = history
    e = orc.onestep_errors(orc.orca_params(), S[:, 600], U[:, 600], S[:, 601], Ts)
    np.testing.assert_allclose(e[0], 4.6925049599691874e-11, rtol=1e-9)


def test_bank_draw_order_matches_reference():
    g = load_golden("lookback_c1.npz")
    bank = orc.make_bank(1024, seed=0)
    ref = g["params"]                                   # rows Bf, Cf, Df, Br, Cr, Dr
    for row, k in enumerate(("Bf", "Cf", "Df", "Br", "Cr", "Dr")):
        assert np.array_equal(bank[k], ref[row]), k


def test_lookback_c1_bit_exact(history):
    g = load_golden("lookback_c1.npz")
    S, U, Ts = history
    W, K = int(g["W"]), int(g["K"])
    bank = orc.make_bank(1024, seed=0)
    shared = orc.orca_params()
    pp = tuple(bank[k] for k in ("Bf", "Cf", "Df", "Br", "Cr", "Dr"))
    for t_end in g["ticks"]:
        t_end = int(t_end)
        pred = orc.evaluate_models_vectorized(shared, 1024, S[:, t_end], U[:, t_end], Ts, pp)
        assert np.array_equal(pred, g["pred_%d" % t_end])
        lb = orc.LookBackOracle(bank, W, Ts, K)
        for idt in range(t_end - W + 1, t_end + 1):
            best, topk, avg = lb.push(S[:, idt], U[:, idt], S[:, idt + 1])
        assert np.array_equal(avg, g["avg_%d" % t_end])
        assert best == int(g["best_%d" % t_end])
        assert np.array_equal(topk, g["topk_%d" % t_end])
        # stateless form used by the GPU path: mean over the (N, W) error matrix
        ew = orc.window_errors(bank, S, U, t_end, W, Ts)
        assert np.array_equal(np.mean(ew, axis=1), avg)


def test_vary14_bit_exact(history):
    g = load_golden("vary14.npz")
    S, U, Ts = history
    t = int(g["tick"])
    bank = {k: g["p_" + k] for k in orc.PARAM_NAMES}
    n = orc.bank_size(bank)
    xb, ub = np.tile(S[:, t], (n, 1)), np.tile(U[:, t], (n, 1))
    assert np.array_equal(orc.diffequation_batch(bank, xb, ub), g["f"])
    assert np.array_equal(orc.onestep_predict(bank, S[:, t], U[:, t], Ts), g["rk4"])


def test_lookahead_kat_bit_exact():
    g = load_golden("lookahead_kat.npz")
    bank = orc.orca_params()
    for row, k in enumerate(("Bf", "Cf", "Df", "Br", "Cr", "Dr")):
        bank[k] = g["params"][row]
    J, best_k, traj = orc.lookahead_rollout(bank, g["x0"], g["U"], g["xref"], g["uprev"], float(g["Ts"]),
                                            return_traj=True)
    assert np.array_equal(J, g["J"])
    assert np.array_equal(traj[:, :, -1, :], g["x_final"])
    assert np.array_equal(best_k, np.argmin(g["J"], axis=1))
    # SURVEY section 4 look-ahead literal
    np.testing.assert_allclose(float(g["J_track_nominal"]), 0.2169732391849316, rtol=0, atol=1e-15)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_mu_estimator_oracle_matches_reference_replay(tag):
    """mu_replay.npz holds the outputs of the reference's OWN lines rt.py:278-282 / :326-366 executed tick by tick over
    the recorded dataset (tests/golden/make_golden_mu.py).  The oracle must reproduce, from tick 0: the raw moving
    average MU_pred (what ConstantSpeed receives), the logged smoothed x 0.95 value MU_preds, and the tick from which
    the planner is fed (idt > W + 1), including the W + 1 warm-up seeds (g = 9.8) inside the 20-tick average."""
    g = load_golden("mu_replay.npz")
    G = {k[2:]: g[k] for k in g.files if k.startswith(tag + "_")}
    W = int(G["LookBack_W"])
    est = orc.MuEstimatorOracle(mass=float(G["mass"]), lf=float(G["lf"]), lr=float(G["lr"]), W=W,
                                smoothing_mu=int(G["smoothing_mu"]), alpha=float(G["mu_alpha"]),
                                mu_init=float(G["mu_init"]), v_factor=float(G["v_factor"]))
    for idt in range(int(G["n_ticks"])):
        mu, scale = est.planner_mu_scale(idt)
        assert scale == G["planner_scale"][idt]
        assert mu == G["planner_mu"][idt]
        prev = G["ind_best_KM"][idt - 1] if idt > 0 else None          # top-10 of the previous tick's look-back
        est.tick(idt, prev, G["Dr_bank"], G["Df_bank"])
        if idt <= W:
            assert np.isnan(G["MU_pred"][idt]) and np.isnan(est.MU_pred)
        else:
            assert est.MU_pred == G["MU_pred"][idt]
        assert est.MU_preds[-1] == G["MU_preds"][idt]
        assert est.Drs_preds[-1] == G["Drs_preds"][idt] and est.Dfs_preds[-1] == G["Dfs_preds"][idt]
    # the two values are NOT interchangeable: the display value sits ~5 % below the planner's
    assert np.nanmax(np.abs(G["MU_pred"] - G["MU_preds"])) > 0.03


def test_lookback_oracle_matches_reference_replay_selection():
    """The same replay pins the stateful look-back oracle (np.roll window, arg-min, top-10, first decision at idt = W
    because the reference skips the transition of tick 0, rt.py:346)."""
    g = load_golden("mu_replay.npz")
    h = load_golden("ethz_history.npz")
    S, U, Ts = h["states"], h["inputs"], float(h["Ts"])
    W, t0, n, seed, N = (int(g["a_" + k]) for k in ("LookBack_W", "t0", "n_ticks", "seed", "N_MODELS"))
    bank = orc.make_bank(N, seed=seed)
    assert np.array_equal(bank["Dr"], g["a_Dr_bank"]) and np.array_equal(bank["Df"], g["a_Df_bank"])
    lb = orc.LookBackOracle(bank, W, Ts, K=10)
    cur = 0
    for idt in range(n):
        if idt > 0:
            best, topk, _ = lb.push(S[:, t0 + idt], U[:, t0 + idt], S[:, t0 + idt + 1])
            if best is not None:
                cur = best
                assert np.array_equal(topk, g["a_ind_best_KM"][idt])
            else:
                assert idt < W
        assert cur == g["a_current_model_idx"][idt]
