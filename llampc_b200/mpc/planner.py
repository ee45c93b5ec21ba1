"""Drop-in for llampc/mpc/planner.py:12-67: ``ConstantSpeed(x0, v0, track, N, Ts, projidx, scale=1., curr_mu=1.)``.

`track` is a ``llampc_b200.tracks.RacelineTable`` or a reference track object (``llampc.tracks.ETHZ`` /
``ETHZMobil`` built with reference='optimal'); the raceline tables are uploaded once and cached.  The planner runs
on the device in fp64 (one thread per vehicle; use ``RacelineTable.plan`` for many vehicles at once).
"""
import numpy as np

from ..tracks import RacelineTable


def ConstantSpeed(x0, v0, track, N, Ts, projidx, scale=1., curr_mu=1.):
    table = track if isinstance(track, RacelineTable) else RacelineTable.from_track(track)
    state = np.zeros(6)
    state[0], state[1], state[3] = x0[0], x0[1], v0
    xref, pout, vr = table.plan(state[None], np.array([projidx]), np.array([curr_mu]), N, Ts, scale)
    return xref[0], int(pout[0]), float(vr[0])
