"""Import the real reference (``/root/reference``) as a second, stronger oracle.  TEST INFRASTRUCTURE ONLY.

The reference's hot path is pure NumPy, but ``llampc/models/dynamic.py:17`` imports ``casadi`` and
``llampc/models/model.py:10`` / ``llampc/tracks/track.py:9`` import ``matplotlib``; neither is installed
here and neither is touched on this path, so empty stub modules are registered before the import.

``/root/reference`` exists only in the build container.  On the GPU box ``available()`` is False and
the tests fall back to the committed golden vectors under ``tests/golden/`` (made with this adapter
by ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))


def _resolve_root():
    """LLAMPC_REFERENCE_ROOT, else the reference checkout of the build container, else the offline install of the
    UNMODIFIED reference package under baseline/_ref (pip install --no-deps --target baseline/_ref; git-ignored, travels to
    the GPU box; it holds the Python package only, not the recorded dataset / raceline files)."""
    cands = [os.environ.get("LLAMPC_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "llampc", "models")):
            return c
    return cands[0] or cands[1]


REFERENCE_ROOT = _resolve_root()

_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "llampc", "models"))


def has_data() -> bool:
    """True when the tree also carries the reference's data files (recorded dataset, raceline tables)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "llampc", "data", "DYN-GPMPC-NOCONS-with_var_speedsETHZ.npz"))


def load():
    """Returns a namespace with the reference's ORCA, Dynamic, evaluate_models_vectorized, tracks, planner."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("casadi", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # `control is 'pwm'` SyntaxWarning, orca.py:37
        from llampc.params import ORCA
        from llampc.models import Dynamic
        from llampc.mpc.evaluate_models_vectorized import evaluate_models_vectorized
        try:                                         # planner / tracks need the raceline data files next to the package
            from llampc.mpc.planner import ConstantSpeed
            from llampc.tracks import ETHZ, ETHZMobil
        except Exception:
            ConstantSpeed = ETHZ = ETHZMobil = None
    ns = types.SimpleNamespace(ORCA=ORCA, Dynamic=Dynamic, ETHZ=ETHZ, ETHZMobil=ETHZMobil,
                               evaluate_models_vectorized=evaluate_models_vectorized,
                               ConstantSpeed=ConstantSpeed, root=REFERENCE_ROOT)
    _loaded = ns
    return ns
