"""GPU diagnostic: worst relative score error on the sigma = 2 bank (plot_comp_time.py:178-192) for short windows,
default (automatic) sine mode, SFU tyre sine and strict polynomial sine (LookBack recompute tick = the same step arithmetic
as K1r / K1v).  The default must show 0 scores above 1e-4: a bank whose tyre-sine argument |C| pi/2 can leave [-pi, pi] is
routed to the polynomial (include/llampc_b200.h, LLAMPC_SIN_AUTO)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.mpc import LookBack
from oracle import llampc_oracle as orc

g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
wide = tuple((k, 2.0) for k in ("Br", "Cr", "Dr", "Bf", "Cf", "Df"))
bank = orc.make_bank(1024, seed=3, variation=wide)
W = 8
for fast in (None, True, False):                   # None = the default: the library picks the sine mode from the bank
    lb = LookBack(bank, W=W, Ts=Ts, K=10, refine=0, fast_sin=fast)
    worst, n_bad, n = 0.0, 0, 0
    for t_end in list(range(70, 1700, 37)) + [1312]:
        ts = np.arange(t_end - W + 1, t_end + 1)
        lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
        lb.evaluate()
        got = lb.avg_errors()
        ref = np.mean(orc.window_errors(bank, S, U, t_end, W, Ts), axis=1)
        rel = np.abs(got - ref) / ref
        worst = max(worst, rel.max()); n_bad += int((rel > 1e-4).sum()); n += rel.size
        if t_end == 1312:
            i = int(np.argmax(rel))
            print("  t_end 1312: worst %.3e at candidate %d  (Bf %.3g Cf %.3g Df %.3g Br %.3g Cr %.3g Dr %.3g) score %.3e"
                  % (rel.max(), i, bank["Bf"][i], bank["Cf"][i], bank["Df"][i], bank["Br"][i], bank["Cr"][i], bank["Dr"][i], ref[i]))
    print("fast_sin=%s (%s, sin_arg_max %.2f rad): worst relative score error %.3e over %d scores, %d above 1e-4"
          % (fast, lb.sine_name, lb.bank.sin_arg_max, worst, n, n_bad))
