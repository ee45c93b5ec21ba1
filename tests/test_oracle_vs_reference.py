"""When the reference tree is present (build container), the restatement must equal the reference's own
functions bit-for-bit in float64 on random inputs, not only on the committed golden vectors."""
import numpy as np
import pytest

from oracle import llampc_oracle as orc
from oracle import reference_adapter as ra

pytestmark = pytest.mark.skipif(not ra.available(), reason="no reference tree (/root/reference or baseline/_ref)")


def test_rhs_and_rk4_random_banks():
    ref = ra.load()
    rng = np.random.RandomState(11)
    p = ref.ORCA(control='pwm')
    for trial in range(5):
        n = 257
        bank = {k: p[k] for k in orc.PARAM_NAMES}
        for k, s in (("Bf", .2), ("Cf", .1), ("Df", .5), ("Br", .2), ("Cr", .1), ("Dr", .5), ("mass", .15)):
            bank[k] = p[k] * (1 + s * rng.randn(n))
        x = np.column_stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(-30, 30, n),
                             rng.uniform(0.05, 3.5, n), rng.uniform(-1, 1, n), rng.uniform(-6, 6, n)])
        u = np.column_stack([rng.uniform(-0.1, 1, n), rng.uniform(-0.35, 0.35, n)])
        m = ref.Dynamic(**bank)
        assert np.array_equal(m._diffequation_batch(None, x, u), orc.diffequation_batch(bank, x, u))
        assert np.array_equal(m._integrate_batch(x, u, 0, 0.02), orc.rk4_step_batch(bank, x, u, 0, 0.02))
        ff = m.calc_forces_batch(x, u, return_slip=True)
        gg = orc.calc_forces_batch(bank, x, u, return_slip=True)
        for a, b in zip(ff, gg):
            assert np.array_equal(a, b)


def test_evaluate_models_vectorized_and_plant(history):
    ref = ra.load()
    S, U, Ts = history
    p = ref.ORCA(control='pwm')
    bank = orc.make_bank(500, seed=5)
    pp = tuple(bank[k] for k in ("Bf", "Cf", "Df", "Br", "Cr", "Dr"))
    m0 = ref.Dynamic(**p)
    for t in (5, 333, 1234):
        a = ref.evaluate_models_vectorized([m0] * 500, 500, S[:, t], U[:, t], Ts, pp)
        b = orc.evaluate_models_vectorized(orc.orca_params(), 500, S[:, t], U[:, t], Ts, pp)
        assert np.array_equal(a, b)
        assert np.array_equal(m0._integrate(S[:, t], U[:, t], 0, Ts), orc.rk6_step(orc.orca_params(), S[:, t], U[:, t], 0, Ts))


def test_orca_constants():
    ref = ra.load()
    p = ref.ORCA(control='pwm')
    q = orc.orca_params()
    for k in q:
        assert p[k] == q[k], k
