"""GPU experiment (run under gpurun): device time of the one-launch tick for a few bank sizes (library chosen with
LLAMPC_LIB, kernel knobs with LLAMPC_K1_PACKED / LLAMPC_TREE_KERNEL)."""
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

if __name__ == "__main__":
    res = []
    for (N, W) in ((65536, 50), (1048576, 50), (131072, 50), (1024, 20)):
        for split in ((0, 1, 2, 4, 8) if N == 65536 else (0,)):
            var = orc.RT_VARIATION + (("mass", 0.15),)
            lb = LookBack(orc.make_bank(N, 1, variation=var), W=W, Ts=Ts, K=10, refine=0, split=split)
            ts = np.arange(600 - W + 1, 601)
            lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
            st = torch.cuda.current_stream().cuda_stream

            def tick():
                lb._tick.row32_h, lb._tick.row64_h, lb._tick.slot, lb._tick.sync = None, None, 0, 0
                L.llampc_lookback_tick(lb._tick_ref, st)
            t = timed(tick, 200)
            res.append("N=%d W=%d split=%d: %.1f us (%.3e steps/s)" % (N, W, split, t, N * W / t * 1e6))
            del lb
    print(os.environ.get("LLAMPC_LIB", "default"), os.environ.get("LLAMPC_K1_PACKED", "1"), " | ".join(res), flush=True)
