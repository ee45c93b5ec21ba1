"""``Dynamic`` -- the 6-state dynamic bicycle with Pacejka tyres, GPU-backed.

Same constructor and batch methods as the reference class (llampc/models/dynamic.py:22-154 and its base
llampc/models/model.py:12-40): parameters may be scalars or (N,) arrays; ``_integrate_batch`` is one classic
RK4 step (rk6.py:50-68), ``_integrate`` / ``sim_continuous`` use the 6-stage plant integrator (rk6.py:13-28).
Every method launches kernels of libllampc_b200.so; nothing is computed in NumPy.

Not provided (outside the LLA-MPC look-back/look-ahead path, they belong to the CasADi/IPOPT side):
``casadi``, ``casadi_parametric``, ``sim_discrete``, ``linearize``, the linear-tyre ``approx`` branch,
``input_acc`` and ``carla`` variants.
"""
import numpy as np

from .. import _lib
from ..bank import ModelBank, PARAM_NAMES


class Dynamic:

    def __init__(self, lf, lr, mass, Iz, Cf, Cr, Bf=None, Br=None, Df=None, Dr=None,
                 Cm1=None, Cm2=None, Cr0=None, Cr2=None, input_acc=False, carla=False, **kwargs):
        self.lf, self.lr, self.mass, self.Iz = lf, lr, mass, Iz
        self.dr = lr / (lf + lr)
        self.Cf, self.Cr, self.Bf, self.Br, self.Df, self.Dr = Cf, Cr, Bf, Br, Df, Dr
        self.Cm1, self.Cm2, self.Cr0, self.Cr2 = Cm1, Cm2, Cr0, Cr2
        self.approx = Bf is None or Br is None or Df is None or Dr is None
        self.input_acc, self.carla = input_acc, carla
        self.n_states, self.n_inputs = 6, 2
        if self.approx or input_acc or carla:
            raise NotImplementedError("llampc_b200.Dynamic implements the pwm / Pacejka model used by LLA-MPC only")
        self._bank = None

    # ------------------------------------------------------------------ device bank
    def _params(self):
        return {k: getattr(self, k) for k in PARAM_NAMES}

    def bank(self):
        if self._bank is None:
            self._bank = ModelBank(self._params())
        return self._bank

    def _batch_call(self, fn_name, x_batch, u_batch, cols, Ts=None):
        torch = _lib.require_cuda()
        bank = self.bank()
        x = np.ascontiguousarray(x_batch, dtype=np.float64)
        u = np.ascontiguousarray(u_batch, dtype=np.float64)
        if x.ndim == 1:
            x = x[None]
        if u.ndim == 1:
            u = u[None]
        n = max(bank.N, x.shape[0], u.shape[0])
        if bank.N not in (1, n):
            raise ValueError("batch of %d rows does not match a bank of %d models" % (n, bank.N))
        if bank.N == 1 and n > 1:                       # scalar-parameter model applied to a batch of states
            key = ("rep", n)
            if getattr(self, "_rep", (None,))[0] != key:
                self._rep = (key, ModelBank({k: np.full(n, float(getattr(self, k))) for k in PARAM_NAMES}))
            bank = self._rep[1]
        xs, us = int(x.shape[0] == 1 and n > 1), int(u.shape[0] == 1 and n > 1)
        xd = torch.from_numpy(x[:, :6].copy()).to(bank.device)
        ud = torch.from_numpy(u[:, :2].copy()).to(bank.device)
        out = torch.empty((n, cols), dtype=torch.float64, device=bank.device)
        L = _lib.lib()
        st = _lib.stream_ptr(torch)
        with torch.cuda.device(bank.device):
            if fn_name == "rk4":
                rc = L.llampc_rk4_batch_f32(bank.packed.data_ptr(), n, bank.Npad, xd.data_ptr(), xs, ud.data_ptr(), us,
                                            float(Ts), out.data_ptr(), cols, st)
            elif fn_name == "rhs":
                rc = L.llampc_rhs_batch_f32(bank.packed.data_ptr(), n, bank.Npad, xd.data_ptr(), xs, ud.data_ptr(), us,
                                            out.data_ptr(), st)
            else:
                rc = L.llampc_forces_batch_f32(bank.packed.data_ptr(), n, bank.Npad, xd.data_ptr(), xs, ud.data_ptr(),
                                               us, out.data_ptr(), st)
        _lib.check(rc, "llampc_%s_batch_f32" % fn_name)
        return out.cpu().numpy()

    # ------------------------------------------------------------------ batched API (model.py:32-40, dynamic.py:98-154)
    def _integrate_batch(self, x_t_batch, u_t_batch, t_start, t_end):
        return self._batch_call("rk4", x_t_batch, u_t_batch, 6, Ts=t_end - t_start)

    def _diffequation_batch(self, t, x_batch, u_batch):
        return self._batch_call("rhs", x_batch, u_batch, 6)

    def calc_forces_batch(self, x_batch, u_batch, return_slip=False):
        o = self._batch_call("forces", x_batch, u_batch, 5)
        if return_slip:
            return o[:, 0], o[:, 1], o[:, 2], o[:, 3], o[:, 4]
        return o[:, 0], o[:, 1], o[:, 2]

    # ------------------------------------------------------------------ scalar API (dynamic.py:59-96,156-193; model.py:18-30)
    def _diffequation(self, t, x, u):
        return self._batch_call("rhs", np.asarray(x, dtype=np.float64)[None], np.asarray(u, dtype=np.float64)[None], 6)[0]

    def calc_forces(self, x, u, return_slip=False):
        o = self._batch_call("forces", np.asarray(x, dtype=np.float64)[None], np.asarray(u, dtype=np.float64)[None], 5)[0]
        return tuple(o) if return_slip else tuple(o[:3])

    def _integrate(self, x_t, u_t, t_start, t_end):
        """Plant step (RK6, fp64 on the device)."""
        return self.plant_step(np.asarray(x_t, dtype=np.float64)[None], np.asarray(u_t, dtype=np.float64)[None],
                               t_end - t_start)[0]

    def plant_step(self, x, u, Ts):
        """V independent vehicles, one RK6 step each: x (V,6), u (V,2) -> (V,6) float64."""
        torch = _lib.require_cuda()
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        V = x.shape[0]
        p = np.stack([np.broadcast_to(np.asarray(getattr(self, k), dtype=np.float64), (V,)) for k in PARAM_NAMES], axis=1)
        dev = torch.device("cuda", torch.cuda.current_device())
        pd, xd, ud = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (p, x, u))
        out = torch.empty((V, 6), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().llampc_plant_rk6_f64(pd.data_ptr(), V, xd.data_ptr(), ud.data_ptr(), float(Ts),
                                                   out.data_ptr(), _lib.stream_ptr(torch)), "llampc_plant_rk6_f64")
        return out.cpu().numpy()

    def sim_continuous(self, x0, u, t):
        """dynamic.py:59-74: x0 (6,), u (2,n), t (n+1,) -> x (6,n+1), dxdt (6,n+1)."""
        n_steps = u.shape[1]
        x = np.zeros([6, n_steps + 1])
        dxdt = np.zeros([6, n_steps + 1])
        dxdt[:, 0] = self._diffequation(None, x0, [0, 0])
        x[:, 0] = x0
        for ids in range(1, n_steps + 1):
            x[:, ids] = self._integrate(x[:, ids - 1], u[:, ids - 1], t[ids - 1], t[ids])
            dxdt[:, ids] = self._diffequation(None, x[:, ids], u[:, ids - 1])
        return x, dxdt

    def casadi(self, *a, **k):
        raise NotImplementedError("symbolic CasADi model: out of scope of the B200 hot path (DESIGN.md)")

    casadi_parametric = sim_discrete = linearize = casadi
