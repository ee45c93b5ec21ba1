"""CPU-side checks: the C-ABI library loads and exports every symbol include/llampc_b200.h declares, and the
host packing functions (no GPU needed) produce the documented layouts."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from llampc_b200 import _lib
from oracle import llampc_oracle as orc


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "llampc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(llampc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.PROTOTYPES, "ctypes prototype missing for %s" % n
    assert lib.llampc_abi_version() == 6
    lookback = [n for n in names if n.startswith('llampc_lookback_')]
    assert {'llampc_lookback_launch', 'llampc_lookback_tick', 'llampc_lookback_push'} <= set(lookback)
    # three look-back entry points + helpers (plan / finish / decode / release / sizes); the six overlapping launch
    # functions of ABI v4 are gone
    assert not [n for n in lookback if 'window' in n or 'rolling' in n or 'balanced' in n]
    assert b'peer' in lib.llampc_error_string(-4)
    assert b"aligned" in lib.llampc_error_string(-2)


def test_no_compute_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from llampc_b200.mpc import LookBack, LookAhead
    from llampc_b200.mpc.montecarlo import MonteCarlo
    from llampc_b200.mpc.evaluate_models_vectorized import evaluate_models_vectorized
    from llampc_b200.models import Dynamic
    from llampc_b200.params import ORCA
    bank = orc.make_bank(16, 0)
    with pytest.raises(_lib.LlampcError):
        LookBack(bank, W=4)
    with pytest.raises(_lib.LlampcError):
        LookAhead(bank, Ts=0.02)
    with pytest.raises(_lib.LlampcError):
        MonteCarlo(bank, None, np.zeros((2, 6)), np.zeros(2, dtype=np.int32), orc.orca_params(), np.zeros(2))
    with pytest.raises(_lib.LlampcError):                     # the reference's own call (rt.py:349): no CPU fallback either
        evaluate_models_vectorized([Dynamic(**ORCA())] * 16, 16, np.zeros(6), np.zeros(2), 0.02,
                                   tuple(bank[k] for k in ("Bf", "Cf", "Df", "Br", "Cr", "Dr")))


def _pack(params, N, Npad):
    ptrs = (C.c_void_p * 14)(*[params[k].ctypes.data for k in _lib.PARAM_NAMES])
    flags = (C.c_int * 14)(*[int(params[k].ndim == 1) for k in _lib.PARAM_NAMES])
    out = np.zeros((4, Npad, 4), dtype=np.float32)
    arg_max = C.c_float(-1.0)
    rc = _lib.lib().llampc_bank_pack_h(C.cast(ptrs, C.c_void_p), C.cast(flags, C.c_void_p), N, Npad, out.ctypes.data,
                                       C.addressof(arg_max))
    assert rc == 0
    # bound of the tyre-sine argument C atan(B alpha): max(|Cf|, |Cr|) pi/2 (sine mode "auto")
    want = max(np.abs(params["Cf"]).max(), np.abs(params["Cr"]).max()) * np.pi / 2
    assert arg_max.value == np.float32(want)
    return out


def test_bank_pack_layout():
    bank = orc.make_bank(300, seed=4)
    params = {k: np.array(bank[k], dtype=np.float64) for k in _lib.PARAM_NAMES}
    out = _pack(params, 300, 384)
    f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32)
    assert np.array_equal(out[0, :300, 0], f32(bank["Bf"]))
    assert np.array_equal(out[0, :300, 3], f32(bank["Br"]))
    assert np.array_equal(out[1, :300, 1], f32(bank["Dr"]))
    assert np.all(out[1, :300, 2] == np.float32(1.0 / bank["mass"]))
    assert np.all(out[2, :300, 1] == np.float32(bank["lf"] / bank["Iz"]))
    assert np.all(out[2, :300, 2] == np.float32(bank["lr"] / bank["Iz"]))
    assert np.all(out[3, :300, 0] == np.float32(bank["Cm2"]))
    # padding rows repeat the last candidate
    assert np.array_equal(out[:, 300:, :], np.broadcast_to(out[:, 299:300, :], (4, 84, 4)))
    assert _lib.lib().llampc_bank_pack_h(None, None, 1, 1, None, None) == -1


def test_hist_row_pack(history):
    S, U, Ts = history
    p = orc.orca_params()
    t = 777
    row = np.zeros(20, dtype=np.float32)
    r64 = np.zeros(12)
    xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, t], U[:, t], S[:, t + 1]))
    rc = _lib.lib().llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, Ts, p["lf"], p["lr"],
                                           row.ctypes.data, r64.ctypes.data)
    assert rc == 0
    psi, vx, vy, w = xk[2:6]
    f32 = np.float32
    assert row[0] == f32(np.sin(psi)) and row[1] == f32(np.cos(psi))
    assert row[2] == f32(np.sin(psi + Ts * w / 2)) and row[5] == f32(np.cos(psi + Ts * w))
    assert row[6] == f32(vx) and row[8] == f32(w) and row[9] == f32(uk[0]) and row[11] == f32(np.sin(uk[1]))
    xd0 = vx * np.cos(psi) - vy * np.sin(psi)
    np.testing.assert_allclose(row[13], (xk1[0] - xk[0]) - Ts / 6 * xd0, rtol=1e-6)
    np.testing.assert_allclose(row[15], (xk1[2] - xk[2]) - Ts * w, rtol=1e-6, atol=1e-12)
    dvx = xk1[3] - xk[3]
    assert float(row[16]) + float(row[17]) == pytest.approx(dvx, rel=1e-13)
    _, _, _, af, ar = orc.calc_forces_batch(p, xk[None], uk[None], return_slip=True)
    assert row[18] == f32(af[0]) and row[19] == f32(ar[0])
    assert np.array_equal(r64, np.concatenate([xk, uk, xk1[:4]]))
    # geometry not shared -> slip slots zero
    _lib.lib().llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, Ts, float("nan"), float("nan"),
                                      row.ctypes.data, None)
    assert row[18] == 0 and row[19] == 0


def test_orca_matches_oracle_constants():
    from llampc_b200.params import ORCA
    p, q = ORCA(), orc.orca_params()
    for k in q:
        assert p[k] == q[k]
    assert p["max_inputs"] == [1.0, 0.35] and p["min_rates"] == [None, -5.0]
    with pytest.raises(NotImplementedError):
        ORCA(control="torque")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_mu_estimator_matches_reference_replay(tag):
    """Host MuEstimator against the golden made by the reference's own lines (rt.py:278-282, :326-344): raw MU_pred for
    the planner, smoothed x 0.95 for logging, W + 1 warm-up seeds, planner fed from tick W + 2 on -- to 1e-12 from tick 0."""
    from conftest import load_golden
    from llampc_b200.mpc.mu_estimator import MuEstimator
    g = load_golden("mu_replay.npz")
    G = {k[2:]: g[k] for k in g.files if k.startswith(tag + "_")}
    W = int(G["LookBack_W"])
    est = MuEstimator(mass=float(G["mass"]), lf=float(G["lf"]), lr=float(G["lr"]), W=W, smoothing_mu=int(G["smoothing_mu"]),
                      alpha=float(G["mu_alpha"]), mu_init=float(G["mu_init"]), v_factor=float(G["v_factor"]))
    for idt in range(int(G["n_ticks"])):
        kw = est.planner_args(idt)
        assert kw.get("scale", 1.0) == G["planner_scale"][idt]
        np.testing.assert_allclose(kw.get("curr_mu", 1.0), G["planner_mu"][idt], rtol=1e-12, atol=0)
        top = G["ind_best_KM"][idt - 1] if idt > 0 else None
        mu = est.tick(idt, None if idt <= W else G["Dr_bank"][top], None if idt <= W else G["Df_bank"][top])
        if idt <= W:
            assert mu is None and est.MU_pred is None
        else:
            np.testing.assert_allclose(mu, G["MU_pred"][idt], rtol=1e-12, atol=0)
        np.testing.assert_allclose(est.mu_display, G["MU_preds"][idt], rtol=1e-12, atol=0)
    with pytest.raises(ValueError):
        est.tick(W + 50)                                            # past warm-up the top-K parameters are required


def test_raceline_table_coefficients_match_reference_splines():
    """Thomas-algorithm spline tables == the reference's dense-solve Spline coefficients (golden, from the reference)."""
    from conftest import load_golden
    from llampc_b200.tracks import RacelineTable
    g = load_golden("planner_kat.npz")
    for name in ("ethz", "ethzmobil"):
        r = load_golden("raceline_%s.npz" % name)
        tab = RacelineTable(r["x"], r["y"], r["speeds"], r["mus"])
        np.testing.assert_allclose(tab.s, r["s"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(tab.coef[:, 1], g[name + "_sx_b"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(tab.coef[:, 2], g[name + "_sx_c"][:-1], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(tab.coef[:, 3], g[name + "_sx_d"], rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(tab.coef[:, 8 + 1], g[name + "_v0_b"], rtol=1e-9, atol=1e-10)
        nmu = len(r["mus"])
        np.testing.assert_allclose(tab.coef[:, 8 + 4 * (nmu - 1) + 2], g[name + "_vlast_c"][:-1], rtol=1e-9, atol=1e-8)


def test_make_bank_draw_order_matches_reference_golden():
    """Product-side bank construction == the reference's loop (golden params of lookback_c1.npz, seed 0)."""
    from conftest import load_golden
    from llampc_b200.bank import make_bank
    from llampc_b200.params import ORCA
    g = load_golden("lookback_c1.npz")
    bank = make_bank(ORCA(), 1024, rng=np.random.RandomState(0))
    for row, k in enumerate(("Bf", "Cf", "Df", "Br", "Cr", "Dr")):
        assert np.array_equal(bank[k], g["params"][row]), k
    assert bank["mass"] == 0.041


def test_ctypes_prototypes_match_header_arity():
    """Every prototype in _lib.PROTOTYPES has as many arguments as the declaration in include/llampc_b200.h."""
    text = open(os.path.join(ROOT, "include", "llampc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = dict(re.findall(r"\b(llampc_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S))
    assert set(decls) == set(_lib.PROTOTYPES)
    for name, params in decls.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.PROTOTYPES[name][1]), (name, n, len(_lib.PROTOTYPES[name][1]))


def test_tick_struct_layout_matches_c():
    lib = _lib.lib()
    assert lib.llampc_tick_sizeof() == C.sizeof(_lib.Tick)
    for which, field in enumerate(("Ts", "avg_err", "result_h", "peer_seq", "rolling", "workspace")):
        assert lib.llampc_tick_offsetof(which) == getattr(_lib.Tick, field).offset, field
    assert lib.llampc_lookback_desc_sizeof() == C.sizeof(_lib.LookbackDesc)


def test_lookback_desc_argument_errors_need_no_gpu():
    """llampc_lookback_plan validates the descriptor before touching the device: argument errors come back as LLAMPC_E_*."""
    lib = _lib.lib()
    d, p = _lib.LookbackDesc(), _lib.LookbackPlan()
    assert lib.llampc_lookback_plan(None, C.byref(p)) == -1
    assert lib.llampc_lookback_plan(C.byref(d), C.byref(p)) == -1             # no bank
    buf = np.zeros(64, dtype=np.float32)
    base = buf.ctypes.data + (-buf.ctypes.data) % 16
    d.bank, d.N, d.Npad, d.n_vehicles, d.W = base, 8, 8, 1, 2000
    assert lib.llampc_lookback_plan(C.byref(d), C.byref(p)) == -3             # W beyond LLAMPC_MAX_W
    d.W, d.K = 4, 17
    assert lib.llampc_lookback_plan(C.byref(d), C.byref(p)) == -3             # K beyond LLAMPC_LIST_LEN
    d.K, d.bank = 4, base + 4
    assert lib.llampc_lookback_plan(C.byref(d), C.byref(p)) == -2             # bank not 16-byte aligned
    d.bank = base
    assert lib.llampc_lookback_plan(C.byref(d), C.byref(p)) == -1             # K > 0 without out
    assert lib.llampc_lookback_launch(None, None) == -1


def test_threshold_filter_selection_argument():
    """The selection rule of K1v (lookback_rolling_vehicle_kernel) restated in NumPy: keys are split
    into groups of 32 (warp x candidate slot), T = the K-th smallest group minimum (all ones when fewer than K groups hold a
    key), and the top-K of the survivors {key <= T} must be the top-K of all keys -- for any bank size, with ties in the
    score word, missing candidates (all-ones keys) and K up to LLAMPC_LIST_LEN."""
    rng = np.random.RandomState(7)
    EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)
    for trial in range(300):
        n = int(rng.choice([1, 2, 5, 31, 32, 33, 300, 512, 777, 1024, 1500, 2048]))
        K = int(rng.randint(1, 17))
        passes = (n + 1023) // 1024
        scores = rng.randint(0, 50 if trial % 3 == 0 else 2 ** 31, size=n).astype(np.uint64)    # trial % 3 == 0: many ties
        keys = np.full(passes * 1024, EMPTY, dtype=np.uint64)
        keys[:n] = (scores << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        # thread t of pass j owns candidates j*1024 + 4t .. + 3; group (j, warp, q) = the q-th candidates of a warp's threads
        grid = keys.reshape(passes, 8, 32, 4)
        minima = np.sort(grid.min(axis=2).reshape(-1))
        T = minima[K - 1]
        survivors = np.sort(keys[(keys <= T) & (keys != EMPTY)])
        want = np.sort(keys[:n])[:K]
        assert len(survivors) >= min(K, n)
        assert np.array_equal(survivors[:K], want), (n, K)


def test_equal_share_partition_argument():
    """The work partition of K1e (lookback_equal_kernel) restated in Python: the (group of 64
    candidates) x (window row) space of L = G W steps is cut into n_warps contiguous ranges [w L / n, (w + 1) L / n); only
    L / 4 warps take a share on tiny problems.  Checked for every shape: the ranges tile the space exactly, every sharing warp
    has work, the closed forms the kernel uses for 'first / last warp of a group' -- ((x + 1) n - 1) / L -- name the true
    owners, every warp between them holds a piece of the group (so the arrival count of a group is wl - wf + 1), and the
    scratch bound maxch holds."""
    EQ_WARPS, SMS = 16, 148
    for G in list(range(1, 24)) + [100, 1024, 2048]:
        for W in list(range(1, 12)) + [20, 50, 64]:
            L = G * W
            grid = max(1, min(SMS, (L + 4 * EQ_WARPS - 1) // (4 * EQ_WARPS)))
            n = grid * EQ_WARPS
            if n > L // 4:
                n = max(1, L // 4)
            lo = [w * L // n for w in range(n + 1)]
            assert lo[0] == 0 and lo[-1] == L and all(b > a for a, b in zip(lo, lo[1:]))
            owner = np.repeat(np.arange(n), np.diff(lo))
            per = max(L // n, 1)
            maxch = min((W + per - 1) // per + 1, n)
            for g in range(G):
                x0 = g * W
                wf, wl = ((x0 + 1) * n - 1) // L, ((x0 + W) * n - 1) // L
                assert owner[x0] == wf and owner[x0 + W - 1] == wl
                assert wl - wf + 1 <= maxch
                assert np.array_equal(np.unique(owner[x0:x0 + W]), np.arange(wf, wl + 1))
