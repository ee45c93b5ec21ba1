"""GPU experiment (run under gpurun): K1b device time against the scheduling knobs the library reads from the
environment at every launch: LLAMPC_BAL_CTAS (CTAs per SM x 100), LLAMPC_BAL_TPW (target tasks per resident warp),
LLAMPC_BAL_RMIN (fewest rows per task)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200 import _lib                       # noqa: E402
from llampc_b200.mpc import LookBack                # noqa: E402
from oracle import llampc_oracle as orc             # noqa: E402
from tools.gpu_balanced_check import timed, S, U, Ts, L   # noqa: E402


def sweep(N, W, combos):
    var = orc.RT_VARIATION + (("mass", 0.15),)
    lb = LookBack(orc.make_bank(N, 1, variation=var), W=W, Ts=Ts, K=10, refine=0, balanced=True)
    ts = np.arange(600 - W + 1, 601)
    lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros(_lib.LIST_LEN + 1, dtype=torch.int64, device=lb.bank.device)
    ws = torch.zeros(256 << 20, dtype=torch.uint8, device=lb.bank.device)
    ref = None

    def bal():
        rc = L.llampc_lookback_window_balanced_f32(lb.bank.packed.data_ptr(), N, lb.bank.Npad, lb.hist.data_ptr(), W, Ts,
                                                   lb.avg_err.data_ptr(), 0, int(lb.bank.geom_shared), 1, 10,
                                                   ws.data_ptr(), ws.numel(), out.data_ptr(), None, 0, 0, 0, st)
        _lib.check(rc, "balanced")
    res = []
    for (c, f, m) in combos:
        os.environ["LLAMPC_BAL_CTAS"], os.environ["LLAMPC_BAL_TPW"], os.environ["LLAMPC_BAL_RMAX"] = str(c), str(f), str(m)
        out.zero_()
        ws.zero_()
        t = timed(bal, 30)
        k = out.cpu().numpy().view(np.uint64)[:11] & np.uint64(0xFFFFFFFF)
        if ref is None:
            ref = k
        res.append("(%d,%d,%d):%.1f%s" % (c, f, m, t, "" if np.array_equal(k, ref) else "!"))
    print("N=%d W=%d us by (CTAs/SM x100, TPW, RMAX): %s" % (N, W, "  ".join(res)), flush=True)


if __name__ == "__main__":
    combos = [(c, t, r) for c in (400, 500, 600) for (t, r) in ((6, 5), (12, 3), (12, 2), (4, 10))]
    sweep(65536, 50, combos)
    sweep(1048576, 50, [(600, 6, 5), (600, 6, 10), (600, 6, 25), (600, 6, 50)])
