"""CPU oracle for the look-ahead reference generator (planner) and its raceline tables.  TEST INFRASTRUCTURE ONLY.

Restates, in float64 NumPy / plain Python, the reference's
  * natural cubic spline     llampc/utils/pycubicspline.py:17-132  (Spline),  :135-162 (Spline2D)
  * point-to-segment projection   llampc/utils/projection.py:11-38
  * Track.project_fast       llampc/tracks/track.py:147-160
  * Track._load_raceline     llampc/tracks/track.py:52-83 (spline over the raceline + one speed spline per mu)
  * ConstantSpeed            llampc/mpc/planner.py:12-67
Pinned by tests/golden/planner_kat.npz (outputs of the reference's own ConstantSpeed, made by
tests/golden/make_golden_planner.py) and, when /root/reference is present, bit-for-bit against the imported
reference (tests/test_planner_oracle.py).
"""
import bisect
import math

import numpy as np


class Spline:
    """Natural cubic spline through (x_i, y_i) -- the arithmetic of pycubicspline.py:17-132, written with array
    operations: same tridiagonal system (first/last row = identity, i.e. c_0 = c_{n-1} = 0), same dense solve, same
    per-segment formulas, so the coefficients are bit-identical to the reference's."""

    def __init__(self, x, y):
        self.x = x
        self.nx = n = len(x)
        h = np.diff(x)
        self.a = [v for v in y]
        a = np.asarray(self.a, dtype=np.float64)
        # system matrix (:107-122): interior rows h_{i-1}, 2(h_{i-1}+h_i), h_i; boundary rows pin c to zero
        M = np.zeros((n, n))
        rows = np.arange(1, n - 1)
        M[rows, rows - 1] = h[:-1]
        M[rows, rows] = 2.0 * (h[:-1] + h[1:])
        M[rows, rows + 1] = h[1:]
        M[0, 0] = 1.0
        M[n - 1, n - 1] = 1.0
        # right-hand side (:124-132)
        rhs = np.zeros(n)
        rhs[1:n - 1] = 3.0 * (a[2:] - a[1:-1]) / h[1:] - 3.0 * (a[1:-1] - a[:-2]) / h[:-1]
        self.c = np.linalg.solve(M, rhs)
        c = self.c
        # remaining coefficients per segment (:40-45)
        self.d = list((c[1:] - c[:-1]) / (3.0 * h))
        self.b = list((a[1:] - a[:-1]) / h - h * (c[1:] + 2.0 * c[:-1]) / 3.0)

    def calc(self, t):
        """Value at t, None outside the knot range (:47-65); segment = bisect(x, t) - 1 (:104)."""
        if t < self.x[0] or t > self.x[-1]:
            return None
        k = bisect.bisect(self.x, t) - 1
        u = t - self.x[k]
        return self.a[k] + self.b[k] * u + self.c[k] * u ** 2.0 + self.d[k] * u ** 3.0


class Spline2D:
    """pycubicspline.py:135-162"""

    def __init__(self, x, y):
        dx, dy = np.diff(x), np.diff(y)
        self.ds = [math.sqrt(idx ** 2 + idy ** 2) for (idx, idy) in zip(dx, dy)]
        s = [0]
        s.extend(np.cumsum(self.ds))
        self.s = s
        self.sx, self.sy = Spline(s, x), Spline(s, y)

    def calc_position(self, s):
        return self.sx.calc(s), self.sy.calc(s)


def projection(point, line):
    """projection.py:11-38"""
    x, x1, x2 = np.array(point[0]), np.array(line[0]), np.array(line[len(line) - 1])
    dir1 = x2 - x1
    dir1 /= np.linalg.norm(dir1, 2)
    proj = x1 + dir1 * np.dot(x - x1, dir1)
    dir2, dir3 = (proj - x1), (proj - x2)
    if np.linalg.norm(dir2, 2) > 0 and np.linalg.norm(dir3, 2) > 0:
        dir2 /= np.linalg.norm(dir2)
        dir3 /= np.linalg.norm(dir3)
        is_on_line = np.linalg.norm(dir2 - dir3, 2) > 1e-10
        if not is_on_line:
            proj = x1 if np.linalg.norm(x1 - proj, 2) < np.linalg.norm(x2 - proj, 2) else x2
    return proj, np.linalg.norm(x - proj, 2)


def project_fast(x, y, raceline):
    """track.py:147-160"""
    n = raceline.shape[1]
    proj, dist = np.empty([2, n - 1]), np.empty([n - 1])
    for idl in range(n - 1):
        proj[:, idl], dist[idl] = projection([(x, y)], [raceline[:, idl], raceline[:, idl + 1]])
    optidx = np.argmin(dist)
    return proj[:, optidx], optidx


class RacelineOracle:
    """What Track._load_raceline builds (track.py:52-83): raceline (2, n), Spline2D over it, one speed spline per mu."""

    def __init__(self, x, y, speeds, mus):
        self.raceline = np.array([x, y])
        self.spline = Spline2D(x, y)
        self.mus = mus
        self.spline_v = [Spline(self.spline.s, vi) for vi in speeds]


def constant_speed(x0, v0, track, N, Ts, projidx, scale=1., curr_mu=1.):
    """planner.py:12-67"""
    raceline = track.raceline
    xy, idx = project_fast(x=x0[0], y=x0[1], raceline=raceline[:, projidx:projidx + 10])
    projidx = idx + projidx
    start = track.raceline[:, :projidx + 2]
    xref = np.zeros([2, N + 1])
    xref[:2, 0] = x0
    dist = np.sum(np.linalg.norm(np.diff(start), 2, axis=0))
    v = max(v0, .01)
    vr = 0.
    for idh in range(1, N + 1):
        dist += scale * v * Ts
        dist = dist % track.spline.s[-1]
        xref[:2, idh] = track.spline.calc_position(dist)
        if curr_mu < track.mus[0]:
            v = track.spline_v[0].calc(dist)
        elif curr_mu > track.mus[-1]:
            v = track.spline_v[-1].calc(dist)
        else:
            i = 0
            for i in range(len(track.mus)):
                if track.mus[i] >= curr_mu:
                    break
            vb = track.spline_v[i - 1].calc(dist)
            va = track.spline_v[i].calc(dist)
            v = vb * (track.mus[i] - curr_mu) / (track.mus[i] - track.mus[i - 1]) + va * (curr_mu - track.mus[i - 1]) / (
                track.mus[i] - track.mus[i - 1])
        if idh == 1:
            vr = v * scale
    return xref, projidx, vr
