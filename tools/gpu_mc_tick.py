"""GPU diagnostic: device time of the Monte-Carlo tick (4,096 vehicles x 1,024 candidates x W = 20, K = 32 x H = 20; CUDA-graph
replay) for the overlap orders of look-back and planner, rolling and recompute look-back."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.mpc.montecarlo import MonteCarlo            # noqa: E402
from llampc_b200.tracks import RacelineTable                  # noqa: E402
from bench import make_bank_rt, NOMINAL, TS                   # noqa: E402

rl = np.load(os.path.join(ROOT, "tests", "golden", "raceline_ethzmobil.npz"))
tab = RacelineTable(rl["x"], rl["y"], rl["speeds"], rl["mus"])
V = 4096
modes = sys.argv[1:] or ["rolling"]
for mode in modes:
    for overlap in ("lookback_first", "plan_first", "none"):
        r4 = np.random.RandomState(4)
        start = r4.randint(0, 400, V)
        x_init = np.zeros((V, 6))
        x_init[:, 0] = 0.6 * rl["x"][start + 1] + 0.4 * rl["x"][start + 2]
        x_init[:, 1] = 0.6 * rl["y"][start + 1] + 0.4 * rl["y"][start + 2]
        x_init[:, 2] = np.arctan2(rl["y"][start + 2] - rl["y"][start + 1], rl["x"][start + 2] - rl["x"][start + 1])
        x_init[:, 3] = 1.0
        mc = MonteCarlo(make_bank_rt(1024, seed=0), tab, x_init, start, NOMINAL, r4.uniform(3.0, 15.0, V), W=20, K_models=10,
                        K_seq=32, H=20, Ts=TS, seed=4, lookback_mode=mode, use_graphs=True, overlap=overlap)
        mc.run(48)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mc.run(40)
        b.record()
        torch.cuda.synchronize()
        h = mc.host()
        print("%-10s overlap=%-22s tick %.1f us   (checksum x %.9f, model %d)" % (mode, overlap, a.elapsed_time(b) * 1e3 / 40,
                                                                                 float(h["x"].sum()), int(h["model_idx"].sum())))
        del mc
