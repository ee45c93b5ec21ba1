"""GPU experiment (run under gpurun): per-CTA timeline of one K1b launch (global timer at entry, after the history
staging, after the last RK4 row, at exit)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LLAMPC_BAL_TRACE"] = "1"
os.environ.setdefault("LLAMPC_LIB", os.path.join(ROOT, "llampc_b200", "libllampc_b200_trace.so"))
from llampc_b200 import _lib                       # noqa: E402
from llampc_b200.mpc import LookBack                # noqa: E402
from oracle import llampc_oracle as orc             # noqa: E402
from tools.gpu_balanced_check import S, U, Ts, L   # noqa: E402


def trace(N, W, ctas_x100=600):
    os.environ["LLAMPC_BAL_CTAS"] = str(ctas_x100)
    var = orc.RT_VARIATION + (("mass", 0.15),)
    lb = LookBack(orc.make_bank(N, 1, variation=var), W=W, Ts=Ts, K=10, refine=0, balanced=True)
    ts = np.arange(600 - W + 1, 601)
    lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(_lib.LIST_LEN + 1, dtype=torch.int64, device=lb.bank.device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    grid = min(148 * ctas_x100 // 100, 8192)
    fn = L.llampc_debug_balanced_trace
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int]
    for it in range(4):
        flush.fill_(1)
        L.llampc_lookback_window_balanced_f32(lb.bank.packed.data_ptr(), N, lb.bank.Npad, lb.hist.data_ptr(), W, Ts,
                                              lb.avg_err.data_ptr(), 0, int(lb.bank.geom_shared), 1, 10,
                                              lb.workspace.data_ptr(), lb.workspace.numel(), out.data_ptr(), None, 0, 0, 0, st)
        torch.cuda.synchronize()
    nw = min(grid * 4, 8192)
    buf = np.zeros((nw, 4), dtype=np.uint64)
    fn(buf.ctypes.data, nw)
    ok = buf[:, 3] > 0
    exit_abs = buf[:, 2].astype(np.int64)
    ref = int(exit_abs[ok].min())
    exit_ = (exit_abs - ref) / 1e3
    last_start = (buf[:, 0].astype(np.int64) - ref) / 1e3
    rows_us = buf[:, 1].astype(np.int64) / 1e3
    ntask = buf[:, 3].astype(np.int64)
    q = lambda a: "min %.1f p10 %.1f med %.1f p90 %.1f p99 %.1f max %.1f" % (a.min(), np.percentile(a, 10), np.median(a), np.percentile(a, 90), np.percentile(a, 99), a.max())
    print("N=%d W=%d warps=%d (us, relative to the FIRST loop exit = pool empty)" % (N, W, nw))
    print("  loop exit        ", q(exit_[ok]))
    print("  last task start  ", q(last_start[ok]))
    print("  last task rows   ", q(rows_us[ok]), " protocol+finalize", q((exit_ - last_start - rows_us)[ok]))
    print("  tasks per warp   ", q(ntask[ok].astype(np.float64)))
    late = np.argsort(exit_)[-12:]
    print("  last exits (cta, warp, tasks, last start, rows us, exit):", [(int(i) // 4, int(i) % 4, int(ntask[i]), round(float(last_start[i]), 1), round(float(rows_us[i]), 1), round(float(exit_[i]), 1)) for i in late], flush=True)


if __name__ == "__main__":
    os.environ["LLAMPC_TREE_KERNEL"] = "k1b"      # the timeline is compiled into K1b only
    trace(65536, 50)
    trace(1048576, 50)
