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

REFERENCE_ROOT = os.environ.get("LLAMPC_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "llampc", "models"))


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
        from llampc.mpc.planner import ConstantSpeed
        from llampc.tracks import ETHZ, ETHZMobil
    ns = types.SimpleNamespace(ORCA=ORCA, Dynamic=Dynamic, ETHZ=ETHZ, ETHZMobil=ETHZMobil,
                               evaluate_models_vectorized=evaluate_models_vectorized,
                               ConstantSpeed=ConstantSpeed, root=REFERENCE_ROOT)
    _loaded = ns
    return ns
