"""GPU diagnostic: 30 ticks of the 4,096-vehicle Monte-Carlo loop (run under ncu --metrics gpu__time_duration.sum
for the per-kernel device times of one tick)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.mpc.montecarlo import MonteCarlo
from llampc_b200.tracks import RacelineTable
from bench import make_bank_rt, NOMINAL, TS
rl = np.load(os.path.join(ROOT, "tests", "golden", "raceline_ethzmobil.npz"))
tab = RacelineTable(rl["x"], rl["y"], rl["speeds"], rl["mus"])
V = 4096
r4 = np.random.RandomState(4)
start = r4.randint(0, 400, V)
x_init = np.zeros((V, 6))
x_init[:, 0] = 0.6 * rl["x"][start + 1] + 0.4 * rl["x"][start + 2]
x_init[:, 1] = 0.6 * rl["y"][start + 1] + 0.4 * rl["y"][start + 2]
x_init[:, 2] = np.arctan2(rl["y"][start + 2] - rl["y"][start + 1], rl["x"][start + 2] - rl["x"][start + 1])
x_init[:, 3] = 1.0
mc = MonteCarlo(make_bank_rt(1024, seed=0), tab, x_init, start, NOMINAL, r4.uniform(3.0, 15.0, V), W=20, K_models=10, K_seq=32,
                H=20, Ts=TS, seed=4, lookback_mode=os.environ.get("MC_MODE", "rolling"))
mc.run(30)
torch.cuda.synchronize()
print("done", mc.tick_count)
