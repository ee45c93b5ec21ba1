"""The planner oracle must reproduce the reference's ConstantSpeed / spline outputs (golden file made from the
reference by tests/golden/make_golden_planner.py) bit-for-bit, and the imported reference when it is present."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import planner_oracle as po
from oracle import reference_adapter as ra


@pytest.fixture(scope="module")
def tracks():
    out = {}
    for name in ("ethz", "ethzmobil"):
        r = load_golden("raceline_%s.npz" % name)
        out[name] = po.RacelineOracle(r["x"], r["y"], r["speeds"], r["mus"])
    return out


@pytest.mark.parametrize("name", ["ethz", "ethzmobil"])
def test_constant_speed_golden_bit_exact(tracks, name):
    g = load_golden("planner_kat.npz")
    trk = tracks[name]
    Ts, H = float(g["Ts"]), int(g["H"])
    for c, case in enumerate(g[name + "_cases"]):
        x0, v0, pid, mu, scale = case[:2], case[2], int(case[3]), case[4], case[5]
        xref, pout, vr = po.constant_speed(x0, v0, trk, H, Ts, pid, scale=scale, curr_mu=mu)
        assert np.array_equal(xref, g[name + "_xref"][c]), c
        assert pout == int(g[name + "_projidx"][c])
        assert vr == g[name + "_vr"][c]


@pytest.mark.parametrize("name", ["ethz", "ethzmobil"])
def test_spline_coefficients_golden(tracks, name):
    g = load_golden("planner_kat.npz")
    trk = tracks[name]
    assert np.array_equal(np.array(trk.spline.sx.b), g[name + "_sx_b"])
    assert np.array_equal(np.array(trk.spline.sx.c), g[name + "_sx_c"])
    assert np.array_equal(np.array(trk.spline.sx.d), g[name + "_sx_d"])
    assert np.array_equal(np.array(trk.spline_v[0].b), g[name + "_v0_b"])
    assert np.array_equal(np.array(trk.spline_v[-1].c), g[name + "_vlast_c"])


def test_survey_lookahead_kat(tracks):
    """SURVEY.md section 4 literal: ETHZ, t=600 of the dataset, projidx_in=254 -> 255, vr = 2.0034526840506954."""
    g = load_golden("lookahead_kat.npz")
    xref, pout, vr = po.constant_speed(g["x0"][:2], g["x0"][3], tracks["ethz"], 20, 0.02, int(g["projidx_in"]),
                                       scale=.9, curr_mu=0.83)
    assert pout == 255 == int(g["projidx_out"])
    assert vr == 2.0034526840506954 == float(g["vr"])
    assert np.array_equal(xref, g["xref"])


@pytest.mark.skipif(not (ra.available() and ra.has_data()), reason="reference tree with its raceline data not present (GPU box)")
def test_against_imported_reference_random(tracks):
    ref = ra.load()
    trk_ref = ref.ETHZMobil(reference='optimal', longer=True)
    trk = tracks["ethzmobil"]
    rng = np.random.RandomState(5)
    for _ in range(25):
        pid = int(rng.randint(0, 480))
        p = trk_ref.raceline[:, pid + int(rng.randint(1, 8))] + 0.05 * rng.randn(2)
        v0, mu = float(rng.uniform(0, 3)), float(rng.uniform(0.4, 1.1))
        a = ref.ConstantSpeed(x0=p, v0=v0, track=trk_ref, N=20, Ts=0.02, projidx=pid, scale=.9, curr_mu=mu)
        b = po.constant_speed(p, v0, trk, 20, 0.02, pid, scale=.9, curr_mu=mu)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2]
