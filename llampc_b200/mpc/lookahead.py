"""Look-ahead horizon rollout on the GPU: M adapted models x K sampled control sequences x H RK4 steps,
scored with the NMPC objective.

The reference has no batched rollout (its look-ahead is the IPOPT NLP); the semantics here are assembled
from reference pieces: integrator = Model._integrate_batch chained H times
(llampc/models/model.py:32-40), cost = the NLP objective (llampc/mpc/nmpc.py:48,66-71,111) with
Q = diag(1,1), P = 0, R = diag(5e-3, 1) (run_nmpc_orca_llampc_rt.py:60-62), xref from ConstantSpeed
(llampc/mpc/planner.py:12-67).
"""
import numpy as np

from .. import _lib
from ..bank import ModelBank


def _pad16(a_bytes):
    return (a_bytes + 15) // 16 * 16


class LookAhead:
    def __init__(self, model_params, Ts=0.02, Q=(1.0, 1.0), R=(5e-3, 1.0), P=(0.0, 0.0), device=None):
        self.torch = _lib.require_cuda()
        self.bank = model_params if isinstance(model_params, ModelBank) else ModelBank(model_params, device)
        self.Ts = float(Ts)
        self.qrp = np.array([Q[0], Q[1], R[0], R[1], P[0], P[1]], dtype=np.float32)

    def _padded(self, arr):
        """float32 device copy whose allocation extends to a multiple of 16 bytes (bulk-copy granularity)."""
        torch = self.torch
        flat = np.ascontiguousarray(arr, dtype=np.float32).ravel()
        n = _pad16(flat.size * 4) // 4
        buf = torch.zeros(n, dtype=torch.float32, device=self.bank.device)
        buf[:flat.size].copy_(torch.from_numpy(flat))
        return buf

    def rollout(self, x0, U, xref, uprev, model_idx=None, return_final=False):
        """x0 (6,) or (M,6); U (K,H,2) shared or (M,K,H,2); xref (2,H+1) [reference layout] or (M,2,H+1);
        uprev (2,) or (M,2); model_idx: optional (M,) rows of the bank to use (default: every model once).
        Returns J (M,K) float32->float64, best_k (M,), and x_final (M,K,6) if return_final."""
        plan = self.plan(x0, U, xref, uprev, model_idx=model_idx, return_final=return_final)
        plan.run()
        return plan.fetch()

    def warm_start(self, x0, U, xref, uprev, model_idx=None):
        """Best sampled sequence of every model in the layout of the reference NLP's decision vector
        (llampc/mpc/nmpc.py:113-117,196-197: xvars = [x(:,0), ..., x(:,H), u(:,0), ..., u(:,H-1)]), to be passed as
        the solver's initial guess ``arg['x0']`` (``setupNLP.solve`` currently starts IPOPT from zeros).
        Returns (J_best (M,), umpc (M,2,H), xmpc (M,6,H+1), guess (M, 6(H+1)+2H))."""
        J, best_k = self.rollout(x0, U, xref, uprev, model_idx=model_idx)
        U = np.asarray(U, dtype=np.float64)
        M = len(best_k)
        Ub = (U[np.arange(M), best_k] if U.ndim == 4 else U[best_k])[:, None]            # (M,1,H,2)
        plan = self.plan(x0, Ub, xref, uprev, model_idx=model_idx, return_traj=True)
        plan.run()
        traj = plan.traj.cpu().numpy()[:, 0]                                               # (M,H+1,6)
        umpc = np.swapaxes(Ub[:, 0], 1, 2)                                                 # (M,2,H)
        xmpc = np.swapaxes(traj, 1, 2)                                                     # (M,6,H+1)
        guess = np.concatenate([traj.reshape(M, -1), Ub[:, 0].reshape(M, -1)], axis=1)
        return J[np.arange(M), best_k], umpc, xmpc, guess

    def plan(self, x0, U, xref, uprev, model_idx=None, return_final=False, _force_general=False, return_traj=False):
        """Upload the inputs once and return a `RolloutPlan`: `.run()` enqueues the kernel on the current
        stream (device-resident inputs), `.fetch()` copies J / best_k (/ x_final) back."""
        torch = self.torch
        dev = self.bank.device
        U = np.asarray(U)
        M = self.bank.N if model_idx is None else len(model_idx)
        flags = 8 if _force_general else 0
        if U.ndim == 4:
            flags |= 1
            if U.shape[0] != M:
                raise ValueError("per-model U must have M rows")
        K, H = U.shape[-3], U.shape[-2]
        xref = np.asarray(xref, dtype=np.float64)
        if xref.ndim == 3:
            flags |= 2
            xr = np.ascontiguousarray(np.swapaxes(xref, 1, 2))          # (M, H+1, 2)
        else:
            xr = np.ascontiguousarray(xref.T)                            # (H+1, 2)
        if xr.shape[-2] != H + 1:
            raise ValueError("xref must have H+1 columns")
        uprev = np.asarray(uprev, dtype=np.float64)
        if uprev.ndim == 2:
            flags |= 4
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        n_x0 = 1 if x0.ndim == 1 else x0.shape[0]
        plan = RolloutPlan()
        plan.owner, plan.M, plan.K, plan.H, plan.flags, plan.n_x0 = self, M, K, H, flags, n_x0
        plan.x0 = torch.from_numpy(x0.reshape(n_x0, 6)).to(dev)
        plan.U, plan.xref, plan.uprev = self._padded(U), self._padded(xr), self._padded(uprev)
        plan.midx = None if model_idx is None else torch.as_tensor(np.asarray(model_idx, dtype=np.int32)).to(dev)
        plan.J = torch.empty((M, K), dtype=torch.float32, device=dev)
        plan.best = torch.empty(M, dtype=torch.int32, device=dev)
        plan.xf = torch.empty((M, K, 6), dtype=torch.float64, device=dev) if return_final else None
        plan.traj = torch.empty((M, K, H + 1, 6), dtype=torch.float64, device=dev) if return_traj else None
        return plan


class RolloutPlan:
    def run(self, bank_ptr=None, pdl=False):
        """Enqueue the rollout.  bank_ptr: another packed bank of the same shape (a sweep over banks); pdl: programmatic
        dependent launch for back-to-back rollouts that do not consume each other's results."""
        o = self.owner
        torch = o.torch
        with torch.cuda.device(o.bank.device):
            rc = _lib.lib().llampc_lookahead_rollout_f32(
                o.bank.packed.data_ptr() if bank_ptr is None else bank_ptr, o.bank.Npad, None if self.midx is None else self.midx.data_ptr(), self.M,
                self.x0.data_ptr(), self.n_x0, self.U.data_ptr(), self.K, self.H, self.xref.data_ptr(),
                self.uprev.data_ptr(), self.flags | (32 if pdl else 0), o.qrp.ctypes.data, o.Ts, self.J.data_ptr(), self.best.data_ptr(),
                None if self.xf is None else self.xf.data_ptr(), None if self.traj is None else self.traj.data_ptr(),
                _lib.stream_ptr(torch))
        _lib.check(rc, "llampc_lookahead_rollout_f32")

    def fetch(self):
        out = (self.J.cpu().numpy().astype(np.float64), self.best.cpu().numpy().astype(np.int64))
        if self.xf is not None:
            out += (self.xf.cpu().numpy(),)
        return out
