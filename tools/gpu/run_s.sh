cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/san_target.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from llampc_b200.mpc import LookBack, LookAhead
from llampc_b200.mpc.lookback import LookbackLaunch
from llampc_b200.bank import ModelBank
from llampc_b200 import _lib
from oracle import llampc_oracle as orc
g = np.load("tests/golden/ethz_history.npz"); S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
# tree merge (K1p + fence/atomic message passing), rolling K1r, multi-vehicle last-CTA merge, K1v, K2p, planner: small sizes
bank = orc.make_bank(9000, seed=0)
for mode in ("recompute", "rolling"):
    lb = LookBack(bank, W=6, Ts=Ts, K=10, refine=16, mode=mode)
    for t in range(600, 612):
        out = lb.push(S[:, t], U[:, t], S[:, t + 1])
    print(mode, out[0])
mb = ModelBank(orc.make_bank(600, seed=1))
L = _lib.lib()
V, W = 5, 6
rows = np.zeros((V, W, 20), dtype=np.float32)
for v in range(V):
    for j in range(W):
        t = 300 + 50 * v + j
        xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, t], U[:, t], S[:, t + 1]))
        L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, Ts, mb.lf_shared, mb.lr_shared, rows[v, j].ctypes.data, None)
hist = torch.from_numpy(rows).cuda()
for kern in ("k1", "k1p"):
    ll = LookbackLaunch(mb, hist, W, Ts, K=10, n_vehicles=V, kernel=kern)
    ll.launch(); ll.launch(); print(kern, ll.keys()[:, 0] & np.uint64(0xffffffff))
ring = torch.zeros((V, W, mb.Npad), dtype=torch.float32, device="cuda")
for kern in (None, "k1r"):
    ll = LookbackLaunch(mb, hist, W, Ts, K=10, n_vehicles=V, mode="rolling", err_ring=ring, kernel=kern)
    for i in range(W + 2):
        ll.launch(slot=i % W, emit=int(i + 1 >= W))
    print(ll.kernel_name, ll.keys()[:, 0] & np.uint64(0xffffffff))
la = LookAhead(orc.make_bank(33, seed=2), Ts=Ts)
rng = np.random.RandomState(3)
Useq = U[:, 600:620].T[None] + np.stack([0.1 * rng.randn(8, 20), 0.05 * rng.randn(8, 20)], axis=-1)
J, bk = la.rollout(S[:, 600], Useq, S[:2, 600:621], U[:, 599]); print("K2p", bk[:5])
from llampc_b200.tracks import RacelineTable
rl = np.load("tests/golden/raceline_ethzmobil.npz")
tab = RacelineTable(rl["x"], rl["y"], rl["speeds"], rl["mus"])
st = np.zeros((37, 6)); idx = np.arange(37) * 11; st[:, 0] = rl["x"][idx + 1]; st[:, 1] = rl["y"][idx + 1] + 0.003; st[:, 3] = 2.0
xr, po, vr = tab.plan(st, idx, np.full(37, 0.83), 20, Ts, 0.9); print("planner", po[:5])
torch.cuda.synchronize(); print("sanitizer target done")
PY
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_target.py > gpurun_out/r2s_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -12 gpurun_out/r2s_memcheck.log
