"""The oracle restatement must reproduce the reference's own outputs (tests/golden, made by
tests/golden/make_golden.py from /root/reference) bit-for-bit in float64."""
import numpy as np

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


def test_mu_estimator_matches_inline_formula():
    est = orc.MuEstimatorOracle(mass=0.041)
    rng = np.random.RandomState(0)
    Drs, Dfs, smooth = [], [], None
    for _ in range(50):
        dr, df = 0.17 + 0.01 * rng.randn(10), 0.19 + 0.01 * rng.randn(10)
        Drs.append(np.mean(dr)); Dfs.append(np.mean(df))
        mu = (np.mean(np.array(Drs)[-20:]) + np.mean(np.array(Dfs)[-20:])) / (9.81 * 0.041)
        smooth = mu if smooth is None else 0.08 * mu + 0.92 * smooth
        assert est.update(dr, df) == smooth * .95
