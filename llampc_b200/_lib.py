"""ctypes binding of libllampc_b200.so (C ABI: include/llampc_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C llampc_b200/csrc``.  There is no
fallback of any kind: if the shared object is missing, importing a compute module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LLAMPC_LIB") or os.path.join(_HERE, "libllampc_b200.so")   # env override: kernel experiments

NPARAM = 14
HIST_ROW = 20
HIST64_ROW = 12
MAX_W = 1024
MAX_K = 64
LIST_LEN = 16
MAX_H = 256
PLAN_WSEG = 48


def ring_rows(W):
    """LLAMPC_RING_ROWS: W error columns + ceil(W / 4) partial-sum rows of a rolling-mode error ring."""
    return W + (W + 3) // 4


PLAN_SPAD = 66
PARAM_NAMES = ("lf", "lr", "mass", "Iz", "Bf", "Br", "Cf", "Cr", "Df", "Dr", "Cm1", "Cm2", "Cr0", "Cr2")

_vp, _i, _f, _d = C.c_void_p, C.c_int, C.c_float, C.c_double


class Tick(C.Structure):
    """llampc_tick_t"""
    _fields_ = [("bank", _vp), ("N", _i), ("Npad", _i),
                ("hist", _vp), ("row32_h", _vp), ("slot", _i), ("W", _i), ("Ts", _d),
                ("geom_shared", _i), ("sin_arg_max", _f), ("sine", _i), ("kernel", _i), ("split", _i), ("idx_offset", _i),
                ("avg_err", _vp), ("K", _i), ("n_refine", _i),
                ("bank64", _vp), ("hist64", _vp), ("row64_h", _vp),
                ("result", _vp), ("result_h", _vp), ("sync", _i), ("zero_copy", _i),
                ("peer_bufs", _vp), ("peer_world", _i), ("peer_rank", _i), ("peer_seq", C.c_uint),
                ("pending_seq", C.c_ulonglong), ("pending_words", _i),
                ("err_ring", _vp), ("rolling", _i),
                ("workspace", _vp), ("workspace_bytes", C.c_ulonglong),
                ("mapped_dev", _vp), ("mapped_for", _vp), ("graph_state", _vp),
                ("hard_h", _vp), ("n_hard", _i), ("replay_state", _vp)]


class LookbackDesc(C.Structure):
    """llampc_lookback_desc_t"""
    _fields_ = [("bank", _vp), ("N", _i), ("Npad", _i), ("idx_offset", _i),
                ("geom_shared", _i), ("sin_arg_max", _f), ("sine", _i),
                ("hist", _vp), ("W", _i), ("n_vehicles", _i), ("hist_stride_rows", _i), ("Ts", _d),
                ("mode", _i), ("slot", _i), ("emit", _i), ("row32_h", _vp), ("err_ring", _vp),
                ("K", _i), ("avg_err", _vp), ("out", _vp),
                ("workspace", _vp), ("workspace_bytes", C.c_ulonglong),
                ("peer_bufs", _vp), ("world", _i), ("rank", _i), ("seq", C.c_uint),
                ("kernel", _i), ("split", _i), ("flags", _i)]


class LookbackPlan(C.Structure):
    """llampc_lookback_plan_t"""
    _fields_ = [("kernel", _i), ("split", _i), ("sine", _i), ("grid_x", _i), ("grid_y", _i), ("block", _i),
                ("launches", _i), ("workspace_bytes", C.c_ulonglong)]


LB_RECOMPUTE, LB_ROLLING = 0, 1
SIN_AUTO, SIN_SFU, SIN_STRICT = 0, 1, 2
KERNEL_AUTO, KERNEL_K1, KERNEL_K1P, KERNEL_K1B, KERNEL_K1R, KERNEL_K1V, KERNEL_K1E = range(7)
KERNEL_NAMES = ("auto", "K1", "K1p", "K1b", "K1r", "K1v", "K1e")
SIN_NAMES = ("auto", "MUFU.SIN", "strict polynomial")
E_PEER = -4
LB_FLAG_PDL = 1
LB_FLAG_WIDE = 2

# name -> (restype, argtypes); every symbol declared in include/llampc_b200.h
PROTOTYPES = {
    "llampc_abi_version": (_i, []),
    "llampc_error_string": (C.c_char_p, [_i]),
    "llampc_bank_pack_h": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "llampc_hist_row_pack_h": (_i, [_vp, _vp, _vp, _d, _d, _d, _vp, _vp]),
    "llampc_lookback_plan": (_i, [C.POINTER(LookbackDesc), C.POINTER(LookbackPlan)]),
    "llampc_lookback_launch": (_i, [C.POINTER(LookbackDesc), _vp]),
    "llampc_topk_scratch_ctas": (_i, [_i]),
    "llampc_topk_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "llampc_refine_f64": (_i, [_vp, _i, _vp, _i, _d, _vp, _i, _i, _vp, _vp]),
    "llampc_lookback_tick_workspace_bytes": (C.c_longlong, [C.POINTER(Tick)]),
    "llampc_lookback_tick": (_i, [C.POINTER(Tick), _vp]),
    "llampc_lookback_tick_release": (_i, [C.POINTER(Tick)]),
    "llampc_lookback_finish": (_i, [C.POINTER(Tick), _vp]),
    "llampc_lookback_decode": (_i, [C.POINTER(Tick), _vp, _vp, _vp]),
    "llampc_tick_sizeof": (_i, []),
    "llampc_tick_offsetof": (_i, [_i]),
    "llampc_lookback_desc_sizeof": (_i, []),
    "llampc_lookback_push": (_i, [C.POINTER(Tick), _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp]),
    "llampc_lookback_replay": (_i, [C.POINTER(Tick), _vp, _vp, _vp, _i, _i, _i, _d, _d, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp,
                                    _vp]),
    "llampc_rk4_batch_f32": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _d, _vp, _i, _vp]),
    "llampc_rhs_batch_f32": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "llampc_forces_batch_f32": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "llampc_lookahead_rollout_f32": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _d,
                                          _vp, _vp, _vp, _vp, _vp]),
    "llampc_planner_constant_speed_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _i, _i, _d, _d,
                                               _vp, _vp, _vp, _vp, _vp]),
    "llampc_pack_rows_f64": (_i, [_vp, _vp, _vp, _i, _d, _d, _d, _i, _i, _vp, _vp, _vp]),
    "llampc_mu_estimate_f64": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "llampc_mu_seed_f64": (_i, [_vp, _i, _i, _i, _d, _d, _vp]),
    "llampc_sample_controls_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "llampc_apply_best_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "llampc_mc_friction_schedule_f64": (_i, [_vp, _i, _i, _i, _vp, _d, _d, _vp, _vp]),
    "llampc_mc_advance_tick_f64": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _d, _vp]),
    "llampc_bank_generate_f32": (_i, [_vp, _vp, _i, _i, C.c_ulonglong, _vp, _vp, _vp, _vp]),
    "llampc_clock_probe": (_i, [_i, _vp, _vp, _vp]),
    "llampc_plant_rk6_f64": (_i, [_vp, _i, _vp, _vp, _d, _vp, _vp]),
}

_lib = None


class LlampcError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LlampcError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C llampc_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)           # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if handle.llampc_abi_version() != 6:
            raise LlampcError("libllampc_b200.so ABI version mismatch")
        if handle.llampc_tick_sizeof() != C.sizeof(Tick) or handle.llampc_lookback_desc_sizeof() != C.sizeof(LookbackDesc):
            raise LlampcError("llampc_tick_t / llampc_lookback_desc_t layout mismatch between _lib.py and "
                              "libllampc_b200.so (rebuild the library)")
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().llampc_error_string(rc).decode()
        if rc > 0:
            try:
                import torch
                torch.cuda.synchronize()
            except Exception as e:                # surface the asynchronous CUDA error text too
                msg += " [%s]" % e
        raise LlampcError("%s failed: rc=%d %s" % (what or "llampc call", rc, msg))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise LlampcError("llampc_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    return torch


def stream_ptr(torch):
    return torch.cuda.current_stream().cuda_stream
