"""GPU diagnostic: device timeline of one Monte-Carlo tick (4,096 vehicles) without a profiler -- a CUDA event is recorded
after every C-ABI launch on the stream it was issued to (main / side), times are relative to the start of the tick."""
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
                H=20, Ts=TS, seed=4)
mc.run(30)
torch.cuda.synchronize()
main = torch.cuda.current_stream()
streams = {main.cuda_stream: ("main", main), mc._side.cuda_stream: ("side", mc._side)}
log = []


class Timed:
    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)

        def call(*a):
            rc = fn(*a)
            tag, s = streams.get(a[-1], ("?", None)) if isinstance(a[-1], int) else ("?", None)
            if s is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(s)
                log.append((name.replace("llampc_", ""), tag, ev))
            return rc
        return call


mc.L = Timed(mc.L)
mc.lb._L = Timed(mc.lb._L)                           # the look-back launch goes through LookbackLaunch
for rep in range(3):
    log.clear()
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record(main)
    mc.tick()
    t1 = torch.cuda.Event(enable_timing=True)
    t1.record(main)
    torch.cuda.synchronize()
    print("tick %d: %.1f us (eager launches, two streams)" % (rep, t0.elapsed_time(t1) * 1e3))
    for name, tag, ev in log:
        print("   %-32s %-4s done at %7.1f us" % (name, tag, t0.elapsed_time(ev) * 1e3))
