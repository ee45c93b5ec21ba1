"""Friction estimate from the K best candidates (host side, a few scalars per tick).

Restates the inline block run_nmpc_orca_llampc_rt.py:326-344 (+ ExponentialSmoother :100-110): mean Dr, Df of
the top-K models, `smoothing_mu`-tick moving average, / (9.81 m), exponential smoothing, x 0.95.
"""
import numpy as np


class MuEstimator:
    def __init__(self, mass, smoothing_mu=20, alpha=0.08, gain=0.95, g=9.81):
        self.mass, self.smoothing_mu, self.alpha, self.gain, self.g = mass, smoothing_mu, alpha, gain, g
        self.Drs_preds, self.Dfs_preds, self.smooth_value = [], [], None

    def update(self, best_Dr, best_Df):
        self.Drs_preds.append(np.mean(best_Dr))
        self.Dfs_preds.append(np.mean(best_Df))
        mu = (np.mean(np.array(self.Drs_preds)[-self.smoothing_mu:])
              + np.mean(np.array(self.Dfs_preds)[-self.smoothing_mu:])) / (self.g * self.mass)
        self.smooth_value = mu if self.smooth_value is None else self.alpha * mu + (1 - self.alpha) * self.smooth_value
        return self.smooth_value * self.gain
