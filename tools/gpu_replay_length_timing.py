import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from llampc_b200.mpc import LookBack
from oracle import llampc_oracle as orc
W = bench.W_C2
S, U = bench.synthetic_history(W + 700, lambda p, x, u: orc.rk6_step(p, x, u, 0, bench.TS))
lb = LookBack(bench.make_bank(bench.N_C2, seed=1), W=W, Ts=bench.TS, K=10, refine=16)
for t in range(W + 5):
    lb.push(S[:, t], U[:, t], S[:, t + 1])
base = W + 5
for T in (8, 16, 32, 64, 128, 256, 8, 256):
    ts = np.arange(base, base + T)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lb.replay(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T, depth=4)
    print("T=%3d: %.1f us per tick" % (T, (time.perf_counter() - t0) / T * 1e6), flush=True)
    time.sleep(0.2)
# same with a pause inside: chunks of 8
t0 = time.perf_counter()
for c in range(32):
    ts = np.arange(base + 8 * c, base + 8 * c + 8)
    lb.replay(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T, depth=4)
print("32 chunks of 8 back to back: %.1f us per tick" % ((time.perf_counter() - t0) / 256 * 1e6))
