"""Raceline tables for the device planner.

Host-side, one-off: the same data the reference builds in Track._load_raceline (llampc/tracks/track.py:52-83) --
a natural cubic spline x(s), y(s) over the raceline points (Spline2D, llampc/utils/pycubicspline.py:135-162) and
one speed spline v_j(s) per friction level mu_j -- flattened into the coefficient tables the planner kernel stages in shared memory.
The spline system is solved with the O(n) Thomas algorithm instead of the reference's dense np.linalg.solve
(identical up to rounding, ~1e-15).
"""
import numpy as np

from . import _lib


def natural_cubic_coeffs(s, y):
    """a, b, c, d (each n-1 long, c also returned full) of the natural cubic spline through (s_i, y_i):
    same system as pycubicspline.py:107-132 (c_0 = c_{n-1} = 0)."""
    s = np.asarray(s, dtype=np.float64)
    a = np.asarray(y, dtype=np.float64)
    n = len(s)
    h = np.diff(s)
    lower = np.zeros(n)
    diag = np.ones(n)
    upper = np.zeros(n)
    rhs = np.zeros(n)
    diag[1:n - 1] = 2.0 * (h[:-1] + h[1:])
    lower[1:n - 1] = h[:-1]
    upper[1:n - 1] = h[1:]
    rhs[1:n - 1] = 3.0 * (a[2:] - a[1:-1]) / h[1:] - 3.0 * (a[1:-1] - a[:-2]) / h[:-1]
    cp = np.zeros(n)
    dp = np.zeros(n)
    cp[0] = upper[0] / diag[0]
    dp[0] = rhs[0] / diag[0]
    for i in range(1, n):                                    # forward sweep
        m = diag[i] - lower[i] * cp[i - 1]
        cp[i] = upper[i] / m
        dp[i] = (rhs[i] - lower[i] * dp[i - 1]) / m
    c = np.zeros(n)
    c[n - 1] = dp[n - 1]
    for i in range(n - 2, -1, -1):                           # back substitution
        c[i] = dp[i] - cp[i] * c[i + 1]
    d = (c[1:] - c[:-1]) / (3.0 * h)
    b = (a[1:] - a[:-1]) / h - h * (c[1:] + 2.0 * c[:-1]) / 3.0
    return a[:-1].copy(), b, c[:-1].copy(), d


class RacelineTable:
    """x, y: raceline points (n,); speeds: (n_mu, n) speed profiles; mus: (n_mu,) ascending friction levels
    (the reference's raceline .npz files: keys 'x', 'y', 'speeds', 'mus')."""

    def __init__(self, x, y, speeds, mus, device=None):
        x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        speeds = np.atleast_2d(np.asarray(speeds, dtype=np.float64))
        mus = np.atleast_1d(np.asarray(mus, dtype=np.float64))
        if speeds.shape != (len(mus), len(x)) or len(x) != len(y) or len(x) < 3:
            raise ValueError("raceline table shapes: x, y (n,), speeds (n_mu, n), mus (n_mu,)")
        ds = np.sqrt(np.diff(x) ** 2 + np.diff(y) ** 2)
        self.s = np.concatenate([[0.0], np.cumsum(ds)])
        self.n, self.n_mu = len(x), len(mus)
        self.raceline = np.array([x, y])
        self.mus = mus
        cols = [np.stack(natural_cubic_coeffs(self.s, x), axis=1), np.stack(natural_cubic_coeffs(self.s, y), axis=1)]
        cols += [np.stack(natural_cubic_coeffs(self.s, v), axis=1) for v in speeds]
        self.coef = np.ascontiguousarray(np.concatenate(cols, axis=1))          # (n-1, 4*(2+n_mu))
        self._dev = None
        self._device = device

    @classmethod
    def from_track(cls, track, device=None):
        """Build from a reference track object (llampc.tracks.ETHZ / ETHZMobil with reference='optimal')."""
        cached = getattr(track, "_llampc_b200_table", None)
        if cached is None:
            cached = cls(track.x_raceline, track.y_raceline, np.asarray(track.v_raceline), np.asarray(track.mus), device)
            try:
                track._llampc_b200_table = cached
            except Exception:
                pass
        return cached

    def planner_tables(self):
        """Host arrays in the layout llampc_planner_constant_speed_f64 wants (include/llampc_b200.h): s, coef_xy
        [n-1][8] and coef_vp [n_mu][n-1][8] (pair j = profiles (j-1) mod n_mu and j), each padded by the window the
        kernel's bulk copies may read past the end (zeros, never used)."""
        n, n_mu, pad = self.n, self.n_mu, _lib.PLAN_WSEG
        s_pad = np.zeros(n + _lib.PLAN_SPAD)
        s_pad[:n] = self.s
        cxy = np.zeros((n - 1 + pad, 8))
        cxy[:n - 1] = self.coef[:, :8]
        cvp = np.zeros((n_mu, n - 1 + pad, 8))
        for j in range(n_mu):
            lo = (j - 1) % n_mu
            cvp[j, :n - 1, :4] = self.coef[:, 8 + 4 * lo:12 + 4 * lo]
            cvp[j, :n - 1, 4:] = self.coef[:, 8 + 4 * j:12 + 4 * j]
        return s_pad, cxy, cvp

    def device_tables(self):
        if self._dev is None:
            torch = _lib.require_cuda()
            dev = torch.device("cuda", torch.cuda.current_device()) if self._device is None else torch.device(self._device)
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
            s_pad, cxy, cvp = self.planner_tables()
            self._dev = (dev, t(s_pad), t(self.raceline.T), t(cxy), t(cvp), t(self.mus))
        return self._dev

    def plan(self, states, projidx, curr_mu, N, Ts, scale=1.0, want_f64=True):
        """ConstantSpeed for V vehicles: states (V,6), projidx (V,), curr_mu (V,) or scalar.
        Returns xref (V,2,N+1) float64, projidx_out (V,), vr (V,)."""
        torch = _lib.require_cuda()
        dev, s, xy, cxy, cvp, mus = self.device_tables()
        states = np.ascontiguousarray(np.atleast_2d(states), dtype=np.float64)
        V = states.shape[0]
        pid = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(projidx, (V,)), dtype=np.int32)).to(dev)
        mu = np.atleast_1d(np.asarray(curr_mu, dtype=np.float64))
        mu_shared = int(mu.shape[0] == 1 and V > 1) or int(V == 1)
        mud = torch.from_numpy(np.ascontiguousarray(mu)).to(dev)
        st = torch.from_numpy(states).to(dev)
        xref = torch.empty((V, N + 1, 2), dtype=torch.float64, device=dev)
        pout = torch.empty(V, dtype=torch.int32, device=dev)
        vr = torch.empty(V, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().llampc_planner_constant_speed_f64(
                s.data_ptr(), xy.data_ptr(), cxy.data_ptr(), cvp.data_ptr(), mus.data_ptr(), self.n, self.n_mu, st.data_ptr(), V,
                pid.data_ptr(), mud.data_ptr(), mu_shared, int(N), float(Ts), float(scale), None, xref.data_ptr(),
                pout.data_ptr(), vr.data_ptr(), _lib.stream_ptr(torch))
        _lib.check(rc, "llampc_planner_constant_speed_f64")
        return np.swapaxes(xref.cpu().numpy(), 1, 2), pout.cpu().numpy().astype(np.int64), vr.cpu().numpy()
