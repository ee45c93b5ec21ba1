"""GPU diagnostic: the look-ahead launch of the Monte-Carlo layout (V vehicles, one model / start state / control table /
reference path / previous input per vehicle, K = 32, H = 20): packed per-model kernel K2q against the scalar kernel K2."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200 import _lib                                     # noqa: E402
from llampc_b200.bank import ModelBank                            # noqa: E402
from oracle import llampc_oracle as orc                           # noqa: E402

L = _lib.lib()
V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K, H, Ts = 32, 20, 0.02
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U0 = g["states"], g["inputs"]
rng = np.random.RandomState(0)
bank = ModelBank(orc.make_bank(1024, seed=0))
midx = torch.from_numpy(rng.randint(0, 1024, V).astype(np.int32)).cuda()
t0s = rng.randint(100, 1700, V)
x0 = torch.from_numpy(np.ascontiguousarray(S[:, t0s].T)).cuda()
Um = np.stack([U0[:, t:t + H].T for t in t0s])[:, None] + np.stack([0.1 * rng.randn(V, K, H), 0.05 * rng.randn(V, K, H)], axis=-1)
Um[..., 0] = np.clip(Um[..., 0], -0.1, 1.0); Um[..., 1] = np.clip(Um[..., 1], -0.35, 0.35)
Ud = torch.from_numpy(Um.astype(np.float32)).cuda()
xr = torch.from_numpy(np.stack([S[:2, t:t + H + 1].T for t in t0s]).astype(np.float32)).cuda()
up = torch.from_numpy(np.ascontiguousarray(U0[:, t0s - 1].T).astype(np.float32)).cuda()
qrp = np.array([1.0, 1.0, 5e-3, 1.0, 0.0, 0.0], dtype=np.float32)
J = torch.empty((V, K), dtype=torch.float32, device="cuda")
bk = torch.empty(V, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for name, extra in (("K2q packed", 0), ("K2 scalar", 16)):
    f = lambda: L.llampc_lookahead_rollout_f32(bank.packed.data_ptr(), bank.Npad, midx.data_ptr(), V, x0.data_ptr(), V, Ud.data_ptr(), K, H,
                                               xr.data_ptr(), up.data_ptr(), 1 | 2 | 4 | extra, qrp.ctypes.data, Ts, J.data_ptr(), bk.data_ptr(),
                                               None, None, st)
    for _ in range(3):
        assert f() == 0
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in evs:
        flush.fill_(1); a.record(); f(); b.record()
    torch.cuda.synchronize()
    ms = np.mean([a.elapsed_time(b) for a, b in evs])
    res[name] = J.cpu().numpy().copy()
    print("%s: V=%d  %.1f us  %.3e steps/s" % (name, V, ms * 1e3, V * K * H / (ms * 1e-3)))
d = np.abs(res["K2q packed"] - res["K2 scalar"]) / res["K2 scalar"]
print("K2q vs K2: median rel diff %.2e, p99 %.2e" % (np.median(d), np.percentile(d, 99)))
