"""GPU diagnostic (run under gpurun): accuracy of the fp32 look-back scores against the fp64 oracle for several
banks / windows, polynomial vs MUFU tyre sine, and raw kernel timings.  Prints one line per case."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200 import _lib                       # noqa: E402
from llampc_b200.mpc import LookBack                # noqa: E402
from oracle import llampc_oracle as orc             # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
L = _lib.lib()


def run_case(name, bank, W, t_end, split=0):
    lb = LookBack(bank, W=W, Ts=Ts, K=10, refine=16, split=split)
    ts = np.arange(t_end - W + 1, t_end + 1)
    lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    best, topk, berr = lb.evaluate()
    avg = lb.avg_errors()
    ref = np.mean(orc.window_errors(bank, S, U, t_end, W, Ts), axis=1)
    rbest, rtopk = orc.select(ref, 10)
    rel = np.abs(avg - ref) / ref
    order = np.argsort(ref)
    print("%-28s N=%7d W=%3d t=%4d split=%d | rel err max %.2e p99 %.2e med %.2e | top100 max %.2e | best %s topk %s | "
          "min ref %.3e best_err64 rel %.1e gap12 %.2e" % (
              name, lb.bank.N, W, t_end, split, rel.max(), np.percentile(rel, 99), np.median(rel), rel[order[:100]].max(),
              best == rbest, list(topk) == list(rtopk), ref.min(), abs(berr - ref[rbest]) / ref[rbest],
              (ref[order[1]] - ref[order[0]]) / ref[order[0]]), flush=True)
    return lb


def time_kernel(lb, split, reps=50):
    """scores-only launch (no selection) of the kernel the LookBack object would run"""
    from llampc_b200.mpc.lookback import LookbackLaunch
    ll = LookbackLaunch(lb.bank, lb.hist, lb.W, lb.Ts, K=0, avg_err=lb.avg_err, split=split)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(5):
        ll.launch()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        ll.launch()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    return ms


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for t_end in (100, 600, 1100, 1600):
        run_case("C1 rt-bank", orc.make_bank(1024, 0), 20, t_end)
    c2var = orc.RT_VARIATION + (("mass", 0.15),)
    for t_end in (600, 1600):
        lb = run_case("C2 +mass", orc.make_bank(65536, 1, variation=c2var), 50, t_end)
    for t_end in (600, 1600):
        run_case("C2 +mass  MUFU.SIN tyre", orc.make_bank(65536, 1, variation=c2var), 50, t_end, split=32 + 2)
    for split in (1, 2, 4, 32 + 1, 32 + 2):
        ms = time_kernel(lb, split)
        print("K1 N=65536 W=50 split=%2d: %.1f us  -> %.3e steps/s" % (split, ms * 1e3, 65536 * 50 / ms * 1e3), flush=True)
    wide = (("Br", 2.0), ("Cr", 2.0), ("Dr", 2.0), ("Bf", 2.0), ("Cf", 2.0), ("Df", 2.0))
    run_case("sigma=2.0 (plot_comp_time)", orc.make_bank(8192, 3, variation=wide), 10, 900)
    run_case("sigma=2.0 MUFU.SIN", orc.make_bank(8192, 3, variation=wide), 10, 900, split=32 + 1)
    run_case("C1 MUFU.SIN", orc.make_bank(1024, 0), 20, 1600, split=32 + 4)
    rng = np.random.RandomState(9)
    p = orc.orca_params()
    all14 = {k: p[k] * (1 + 0.1 * rng.randn(8192)) for k in orc.PARAM_NAMES}
    run_case("all 14 varied", all14, 50, 1200)
    big = orc.make_bank(1 << 20, 5, variation=c2var)
    lbb = LookBack(big, W=50, Ts=Ts, K=10, refine=16)
    ts = np.arange(551, 601)
    lbb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    print("1M bank evaluate:", lbb.evaluate()[0])
    for split in (1, 2, 33):
        ms = time_kernel(lbb, split, reps=10)
        print("K1 N=1M W=50 split=%2d: %.1f us  -> %.3e steps/s" % (split, ms * 1e3, (1 << 20) * 50 / ms * 1e3), flush=True)
    # tick latency through the public API
    t0 = time.perf_counter()
    n = 200
    for i in range(n):
        t = 700 + i
        lb.push(S[:, t], U[:, t], S[:, t + 1])
    dt = (time.perf_counter() - t0) / n
    print("LookBack.push N=65536 W=50 K=10 refine=16: %.1f us per tick" % (dt * 1e6))
