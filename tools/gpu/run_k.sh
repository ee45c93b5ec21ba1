cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "planner or montecarlo or c4" > gpurun_out/r2k_pytest.log 2>&1; tail -6 gpurun_out/r2k_pytest.log
python - <<'PY' > gpurun_out/r2k_planner.txt 2>&1
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from llampc_b200 import _lib
from llampc_b200.tracks import RacelineTable
L = _lib.lib()
for name in ("ethzmobil", "ethz"):
    rl = np.load("tests/golden/raceline_%s.npz" % name)
    tab = RacelineTable(rl["x"], rl["y"], rl["speeds"], rl["mus"])
    V = 4096
    r4 = np.random.RandomState(4)
    start = r4.randint(0, 400, V)
    st = np.zeros((V, 6))
    st[:, 0] = 0.6 * rl["x"][start + 1] + 0.4 * rl["x"][start + 2]
    st[:, 1] = 0.6 * rl["y"][start + 1] + 0.4 * rl["y"][start + 2]
    for v0, mu in ((1.0, 0.8), (2.5, 0.95), (4.5, 1.3)):
        st[:, 3] = v0
        dev, s, xy, cxy, cvp, mus = tab.device_tables()
        sd = torch.from_numpy(st).cuda(); pid = torch.from_numpy(start.astype(np.int32)).cuda()
        mud = torch.full((V,), mu, dtype=torch.float64, device="cuda")
        xref = torch.empty((V, 21, 2), dtype=torch.float32, device="cuda"); pout = torch.empty(V, dtype=torch.int32, device="cuda")
        stm = torch.cuda.current_stream().cuda_stream
        f = lambda: L.llampc_planner_constant_speed_f64(s.data_ptr(), xy.data_ptr(), cxy.data_ptr(), cvp.data_ptr(), mus.data_ptr(), tab.n, tab.n_mu,
                                                        sd.data_ptr(), V, pid.data_ptr(), mud.data_ptr(), 0, 20, 0.02, 0.9, xref.data_ptr(), None, pout.data_ptr(), None, stm)
        for _ in range(3): assert f() == 0
        torch.cuda.synchronize()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in evs:
            flush.fill_(1); a.record(); f(); b.record()
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in evs])
        print("%s V=%d v0=%.1f mu=%.2f: planner %.1f us mean %.1f us min (L2 flushed)" % (name, V, v0, mu, ms.mean() * 1e3, ms.min() * 1e3))
PY
cat gpurun_out/r2k_planner.txt
