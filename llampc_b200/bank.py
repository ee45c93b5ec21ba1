"""Model bank on the device: the 14 Dynamic parameters of N candidate models, packed for the kernels.

Replaces the reference's list of N ``Dynamic`` objects plus the six gathered (N,) arrays
(run_nmpc_orca_llampc_rt.py:145-179) and the array-parameter ``Dynamic`` that
``evaluate_models_vectorized`` rebuilds every tick (evaluate_models_vectorized.py:13-22).
"""
import ctypes as C

import numpy as np

from . import _lib

PARAM_NAMES = _lib.PARAM_NAMES


def _as_param(v):
    a = np.asarray(v, dtype=np.float64)
    if a.ndim > 1:
        raise ValueError("model parameters must be scalars or 1-D arrays")
    return np.ascontiguousarray(a) if a.ndim == 1 else a.copy()      # ascontiguousarray would promote 0-d to 1-d


class ModelBank:
    """params: dict with the 14 keys of ``Dynamic.__init__`` (scalars broadcast, arrays are per candidate).
    Extra keys (limits etc.) are ignored, like ``Dynamic(**ORCA())`` ignores them."""

    @classmethod
    def generate(cls, center, sigmas, n_models, seed=0, device=None):
        """Draw a bank ON THE DEVICE: every parameter named in `sigmas` (dict name -> relative sigma) of every model is
        centre x (1 + sigma randn) (the construction of run_nmpc_orca_llampc_rt.py:145-179 with a counter-based
        Philox generator); `center` is a parameter dict (e.g. ``ORCA()`` or one selected candidate, which re-centres
        the bank on it).  Nothing is materialised on the host until ``.params`` is read."""
        torch = _lib.require_cuda()
        self = cls.__new__(cls)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        unknown = set(sigmas) - set(PARAM_NAMES)
        if unknown:
            raise ValueError("unknown parameters in sigmas: %s" % sorted(unknown))
        c = np.array([float(center[k]) for k in PARAM_NAMES], dtype=np.float64)
        sg = np.array([float(sigmas.get(k, 0.0)) for k in PARAM_NAMES], dtype=np.float64)
        self.N = int(n_models)
        self.Npad = (self.N + 127) // 128 * 128
        self.geom_shared = sg[0] == 0.0 and sg[1] == 0.0
        self.lf_shared = float(c[0]) if self.geom_shared else float("nan")
        self.lr_shared = float(c[1]) if self.geom_shared else float("nan")
        self.packed = torch.empty((4, self.Npad, 4), dtype=torch.float32, device=self.device)
        self._bank64 = torch.empty((_lib.NPARAM, self.N), dtype=torch.float64, device=self.device)
        arg_max = torch.zeros(1, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().llampc_bank_generate_f32(c.ctypes.data, sg.ctypes.data, self.N, self.Npad, int(seed),
                                                           self.packed.data_ptr(), self._bank64.data_ptr(),
                                                           arg_max.data_ptr(), _lib.stream_ptr(torch)),
                       "llampc_bank_generate_f32")
        # max(|Cf|, |Cr|) pi/2 over the bank: bound of the tyre-sine argument (sine mode "auto", include/llampc_b200.h)
        self.sin_arg_max = float(arg_max.item())
        self._params = None
        self._varied = sg != 0.0
        return self

    @property
    def params(self):
        if self._params is None:                                 # generated on the device: fetch on first use
            full = self._bank64.cpu().numpy()
            self._params = {k: (full[j].copy() if self._varied[j] else np.array(full[j, 0]))
                            for j, k in enumerate(PARAM_NAMES)}
        return self._params

    def __init__(self, params, device=None):
        torch = _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        missing = [k for k in PARAM_NAMES if k not in params or params[k] is None]
        if missing:
            raise ValueError("model bank needs all Pacejka parameters; missing %s "
                             "(the linear-tyre 'approx' branch of Dynamic is not on the LLA-MPC path)" % missing)
        self._params = {k: _as_param(params[k]) for k in PARAM_NAMES}
        sizes = {a.shape[0] for a in self._params.values() if a.ndim == 1}
        if len(sizes) > 1:
            raise ValueError("per-candidate parameter arrays differ in length: %s" % sorted(sizes))
        self.N = sizes.pop() if sizes else 1
        self.Npad = (self.N + 127) // 128 * 128
        self.geom_shared = self.params["lf"].ndim == 0 and self.params["lr"].ndim == 0
        self.lf_shared = float(self.params["lf"]) if self.geom_shared else float("nan")
        self.lr_shared = float(self.params["lr"]) if self.geom_shared else float("nan")

        ptrs = (C.c_void_p * _lib.NPARAM)(*[self.params[k].ctypes.data for k in PARAM_NAMES])
        flags = (C.c_int * _lib.NPARAM)(*[int(self.params[k].ndim == 1) for k in PARAM_NAMES])
        packed_h = torch.empty((4, self.Npad, 4), dtype=torch.float32, pin_memory=True)
        arg_max = C.c_float(0.0)
        _lib.check(_lib.lib().llampc_bank_pack_h(C.cast(ptrs, C.c_void_p), C.cast(flags, C.c_void_p), self.N, self.Npad,
                                                 packed_h.data_ptr(), C.addressof(arg_max)), "llampc_bank_pack_h")
        # max(|Cf|, |Cr|) pi/2 over the bank: bound of the tyre-sine argument (sine mode "auto", include/llampc_b200.h)
        self.sin_arg_max = float(arg_max.value)
        self.packed = packed_h.to(self.device, non_blocking=False)
        self._bank64 = None

    @property
    def bank64(self):
        """[14][N] doubles on the device (fp64 finalist re-score)."""
        if self._bank64 is None:
            torch = _lib.require_cuda()
            full = np.stack([np.broadcast_to(self.params[k], (self.N,)) for k in PARAM_NAMES])
            self._bank64 = torch.from_numpy(np.ascontiguousarray(full)).to(self.device)
        return self._bank64

    def param(self, name, idx):
        a = self.params[name]
        return a[idx] if a.ndim == 1 else np.broadcast_to(a, np.shape(idx)).copy()

    def shard(self, rank, world):
        """Contiguous candidate slice [lo, hi) of rank `rank` out of `world`."""
        per = (self.N + world - 1) // world
        lo = min(rank * per, self.N)
        hi = min(lo + per, self.N)
        return lo, hi


def shard_params(params, lo, hi):
    """Slice every per-candidate array of a parameter dict to [lo, hi)."""
    out = {}
    for k in PARAM_NAMES:
        a = np.asarray(params[k], dtype=np.float64)
        out[k] = a[lo:hi] if a.ndim == 1 else a
    return out


# variation of run_nmpc_orca_llampc_rt.py:153-158 (name, sigma) in the reference's dict order
RT_VARIATION = (("Br", 0.2), ("Cr", 0.1), ("Dr", 0.5), ("Bf", 0.2), ("Cf", 0.1), ("Df", 0.5))


def make_bank(nominal, n_models, variation=RT_VARIATION, rng=None):
    """Model-bank construction of run_nmpc_orca_llampc_rt.py:145-179: every varied parameter of every model is the
    nominal value times (1 + sigma * randn), drawn model by model in the order of `variation` (the reference's dict
    order), so a seeded ``np.random.RandomState`` reproduces the reference's bank draw for draw.  Returns a parameter
    dict for ``LookBack`` / ``ModelBank`` (varied entries are (n_models,) arrays, the rest scalars)."""
    rng = np.random if rng is None else rng
    z = rng.randn(n_models, len(variation))
    bank = {k: nominal[k] for k in PARAM_NAMES}
    for j, (name, sigma) in enumerate(variation):
        bank[name] = nominal[name] * (1 + sigma * z[:, j])
    return bank
