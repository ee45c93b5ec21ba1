"""B200-native look-back / look-ahead hot path of LLA-MPC behind the reference's Python API.

    from llampc_b200.params import ORCA
    from llampc_b200.models import Dynamic
    from llampc_b200.mpc.evaluate_models_vectorized import evaluate_models_vectorized
    from llampc_b200.mpc import LookBack, LookAhead

All arithmetic of the hot path runs in hand-written sm_100a kernels inside ``libllampc_b200.so`` (C ABI in
``include/llampc_b200.h``).  There is no CPU fallback: compute calls raise if the library or a CUDA device
is missing.
"""
__version__ = "0.1.0"
