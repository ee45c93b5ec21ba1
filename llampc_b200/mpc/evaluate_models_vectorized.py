"""Drop-in for llampc/mpc/evaluate_models_vectorized.py:4-23 (same name, arguments and return value).

    pred = evaluate_models_vectorized(MODEL_BANK, N_MODELS, states[:, idt], inputs[:, idt], Ts, params_pass)

returns the (N, 4) float64 one-step RK4 predictions [x, y, psi, vx] of every candidate.  As in the reference,
`n_models` is ignored, `params` is the 6-tuple (Bfs, Cfs, Dfs, Brs, Crs, Drs) and every other parameter is
taken from ``models[0]`` (:16-19).  The packed device bank is cached between ticks (keyed on the parameter values).
"""
import numpy as np

from .. import _lib
from ..bank import ModelBank

_cache = {}


def _shared_of(m0):
    return {k: getattr(m0, k) for k in ("mass", "lf", "lr", "Iz", "Cm1", "Cm2", "Cr0", "Cr2")}


def _bank_for(models, params):
    """Packed device bank for (params, models[0]); rebuilt whenever ANY value differs from the cached call.  The cache is
    keyed on the array CONTENTS (compared against private copies: ~0.3 ms at 65,536 candidates, next to a 2 MB result
    copy), not on sums or identities, so a bank that was permuted, resampled or edited in place never reuses a stale one."""
    Bfs, Cfs, Dfs, Brs, Crs, Drs = params
    m0 = models[0]
    arrs = [np.asarray(a, dtype=np.float64) for a in (Bfs, Cfs, Dfs, Brs, Crs, Drs)]
    shared = {k: np.asarray(v, dtype=np.float64) for k, v in _shared_of(m0).items()}
    hit = _cache.get("bank")
    if hit is not None:
        c_arrs, c_shared, bank = hit
        if (all(a.shape == c.shape and np.array_equal(a, c, equal_nan=True) for a, c in zip(arrs, c_arrs))
                and all(np.array_equal(shared[k], c_shared[k], equal_nan=True) for k in shared)):
            return bank
    p = dict(shared)
    p.update(Bf=arrs[0], Cf=arrs[1], Df=arrs[2], Br=arrs[3], Cr=arrs[4], Dr=arrs[5])
    bank = ModelBank(p)
    if bank.N != len(models) and all(a.ndim == 0 for a in arrs):
        raise ValueError("params must be per-candidate arrays")
    _cache["bank"] = ([a.copy() for a in arrs], {k: v.copy() for k, v in shared.items()}, bank)
    return bank


def invalidate_cache():
    """Drop the cached device bank (frees its HBM)."""
    _cache.clear()


def evaluate_models_vectorized(models, n_models, current_state, input_val, Ts, params):
    torch = _lib.require_cuda()
    bank = _bank_for(models, params)
    return onestep(bank, current_state, input_val, Ts, cols=4)


def onestep(bank, state, input_val, Ts, cols=6):
    """One RK4 step of every model of `bank` from the shared (state, input); (N, cols) float64."""
    torch = _lib.require_cuda()
    x = torch.as_tensor(np.ascontiguousarray(state, dtype=np.float64)[:6]).to(bank.device)
    u = torch.as_tensor(np.ascontiguousarray(input_val, dtype=np.float64)[:2]).to(bank.device)
    out = torch.empty((bank.N, cols), dtype=torch.float64, device=bank.device)
    with torch.cuda.device(bank.device):
        _lib.check(_lib.lib().llampc_rk4_batch_f32(bank.packed.data_ptr(), bank.N, bank.Npad, x.data_ptr(), 1,
                                                   u.data_ptr(), 1, float(Ts), out.data_ptr(), cols,
                                                   _lib.stream_ptr(torch)), "llampc_rk4_batch_f32")
    return out.cpu().numpy()
