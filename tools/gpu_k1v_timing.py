"""GPU diagnostic: device time of the Monte-Carlo rolling look-back launch alone (llampc_lookback_rolling_multi_f32,
4,096 vehicles x 1,024 candidates x 20-slot ring = 335 MB), K1v (one CTA per vehicle) against K1r
(LLAMPC_K1R_CTA=0).  LLAMPC_LIB selects an experimental build."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200 import _lib
from llampc_b200.bank import ModelBank
from bench import make_bank_rt, TS

L = _lib.lib()
V, N, W, K = int(os.environ.get("V", 4096)), int(os.environ.get("N", 1024)), 20, 10
bank = ModelBank(make_bank_rt(N, seed=0))
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U = g["states"], g["inputs"]
rows = np.zeros((W, 20), dtype=np.float32)
for j in range(W):
    t = 600 + j
    xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, t], U[:, t], S[:, t + 1]))
    L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, TS, bank.lf_shared, bank.lr_shared,
                             rows[j].ctypes.data, None)
hist = torch.from_numpy(np.tile(rows[None], (V, 1, 1))).cuda().contiguous()
ring = torch.zeros((V, W, bank.Npad), dtype=torch.float32, device="cuda")
keys = torch.full((V,), -1, dtype=torch.int64, device="cuda")
lists = torch.empty((V, (N + 127) // 128, 16), dtype=torch.int64, device="cuda")
ticket = torch.zeros(V, dtype=torch.int32, device="cuda")
out = torch.zeros((V, 17), dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def tick(i, emit=1):
    _lib.check(L.llampc_lookback_rolling_multi_f32(bank.packed.data_ptr(), N, bank.Npad, hist.data_ptr(), V, i % W, W, TS,
                                                   ring.data_ptr(), None, keys.data_ptr(), lists.data_ptr(), 0,
                                                   int(bank.geom_shared), emit, K, ticket.data_ptr(), out.data_ptr(), st))


for mode in ("1", "0"):
    os.environ["LLAMPC_K1R_CTA"] = mode
    for i in range(W + 5):
        tick(i, int(i + 1 >= W))
    torch.cuda.synchronize()
    ts = []
    for i in range(60):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tick(i)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = np.sort(ts)
    print("LLAMPC_K1R_CTA=%s  V=%d N=%d W=%d: p50 %.1f us  min %.1f us  (%.2e candidate-ticks/s, ring %.0f MB -> %.0f GB/s)"
          % (mode, V, N, W, np.median(ts), ts[0], V * N / (np.median(ts) * 1e-6), ring.numel() * 4 / 1e6,
             ring.numel() * 4 / (np.median(ts) * 1e-6) / 1e9))
    res = out.cpu().numpy().copy()
    if mode == "1":
        first = res
    else:
        print("top-%d keys identical to K1r: %s" % (K, bool(np.array_equal(first[:, :K + 1], res[:, :K + 1]))))
