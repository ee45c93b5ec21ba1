"""Friction estimate from the K best candidates (host side, a few scalars per tick).

Mirrors the inline block run_nmpc_orca_llampc_rt.py:326-344 (+ ExponentialSmoother :103-113) and the planner call
:278-282 that consumes it.  Per tick ``idt`` of the closed loop the reference

* while ``idt <= LookBack_W`` appends the PRIOR ``mu_init * m * 9.8 * lr / (lf + lr)`` / ``... * lf / (lf + lr)`` to
  ``Drs_preds`` / ``Dfs_preds`` (:326-330, g = 9.8 there): W + 1 seeds that stay inside the 20-tick moving average
  for the first estimates;
* afterwards appends the mean Dr, Df of the top-K models of the previous tick's look-back (``ind_best_KM``) and forms
  ``MU_pred`` = (mean of the last ``smoothing_mu`` entries of each list) / (9.81 m)  (:341) -- the RAW moving average.
  ``MU_pred`` is what ``ConstantSpeed`` receives as ``curr_mu`` (:278-280), and only from tick ``LookBack_W + 2`` on;
  before that the planner runs with its defaults ``curr_mu = 1, scale = 1`` (:282, planner.py:12);
* the exponentially smoothed value x 0.95 is appended to ``MU_preds`` (:344), which is only saved and plotted (:465).

    est = MuEstimator(mass=params["mass"], lf=params["lf"], lr=params["lr"], W=LookBack_W)
    for idt in ...:
        xref, projidx, v = ConstantSpeed(..., **est.planner_args(idt))          # rt.py:278-282
        ...
        est.tick(idt, best_Dr, best_Df)                                         # rt.py:326-344
"""
import numpy as np


class MuEstimator:
    def __init__(self, mass, lf, lr, W, smoothing_mu=20, alpha=0.08, gain=0.95, mu_init=1.0, v_factor=0.9):
        self.mass, self.lf, self.lr, self.W = float(mass), float(lf), float(lr), int(W)
        self.smoothing_mu, self.alpha, self.gain = int(smoothing_mu), float(alpha), float(gain)
        self.mu_init, self.v_factor = float(mu_init), float(v_factor)
        self.Drs_preds, self.Dfs_preds, self.MU_preds = [], [], []
        self.MU_pred = None                     # raw moving average (rt.py:341); None until tick W + 1 produced it
        self.smooth_value = None

    def tick(self, idt, best_Dr=None, best_Df=None):
        """The friction block of tick ``idt``.  ``best_Dr`` / ``best_Df``: Dr, Df of the K best models of the previous
        tick's look-back (ignored while ``idt <= W``).  Returns ``MU_pred`` (None during warm-up)."""
        if idt <= self.W:                                                       # rt.py:326-330
            self.Drs_preds.append(self.mu_init * self.mass * 9.8 * self.lr / (self.lf + self.lr))
            self.Dfs_preds.append(self.mu_init * self.mass * 9.8 * self.lf / (self.lf + self.lr))
            self.MU_preds.append(self.mu_init)
            return None
        if best_Dr is None or best_Df is None:
            raise ValueError("tick %d > W needs the Dr, Df of the previous tick's top-K models" % idt)
        self.Drs_preds.append(np.mean(best_Dr))                                 # rt.py:339-340
        self.Dfs_preds.append(np.mean(best_Df))
        self.MU_pred = (np.mean(np.array(self.Drs_preds)[-self.smoothing_mu:])
                        + np.mean(np.array(self.Dfs_preds)[-self.smoothing_mu:])) / (9.81 * self.mass)      # :341
        if self.smooth_value is None:                                           # ExponentialSmoother.update, :108-113
            self.smooth_value = self.MU_pred
        else:
            self.smooth_value = self.alpha * self.MU_pred + (1 - self.alpha) * self.smooth_value
        self.MU_preds.append(self.smooth_value * self.gain)                     # :344
        return self.MU_pred

    @property
    def mu_display(self):
        """Last entry of ``MU_preds`` (smoothed x 0.95; ``mu_init`` during warm-up): the logged / plotted value."""
        return self.MU_preds[-1] if self.MU_preds else self.mu_init

    def planner_args(self, idt):
        """Keyword arguments of the ``ConstantSpeed`` call of tick ``idt`` (rt.py:278-282): the raw ``MU_pred`` and
        ``scale = v_factor`` once ``idt > W + 1``, the planner's own defaults (curr_mu = 1, scale = 1) before."""
        if idt > self.W + 1:
            return {"curr_mu": self.MU_pred, "scale": self.v_factor}
        return {}
