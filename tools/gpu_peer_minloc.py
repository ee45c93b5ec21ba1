"""Multi-GPU parity of the path `bench.py --gpus N` times (torchrun, >= 2 GPUs): llampc_lookback_launch on a sharded bank
with the NVLink min-loc fused into the merge-tree root (peer_bufs set), for >= 120 consecutive ticks.

Every tick, on every rank, out[0] (the GLOBAL arg-min key) must equal
  (a) every other rank's out[0]                       (all-gather),
  (b) the NCCL variant: MIN all-reduce of the rank-local arg-min keys,
  (c) at a few ticks, the arg-min of the float64 oracle over the WHOLE bank (index; fp32 score within tolerance).
The tick sequence covers the parity wrap of the double-buffered exchange (seq odd / even) and deliberate skew: one rank
is delayed by a long dummy kernel on alternating ticks, so a fast rank runs up to one tick ahead of a slow one.
Prints one line per rank: OK/FAIL and the per-tick device time of both variants; exit code 0 only if every check passed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.bank import ModelBank                        # noqa: E402
from llampc_b200.dist import PeerExchange, shard_range        # noqa: E402
from llampc_b200.mpc.lookback import LookbackLaunch, decode_keys   # noqa: E402
from llampc_b200 import _lib                                  # noqa: E402
from oracle import llampc_oracle as orc                       # noqa: E402  (checker only)
from bench import make_bank, W_C2, TS                         # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
L = _lib.lib()
N = int(os.environ.get("PEER_N", str(32768 * world)))
N_TICKS = int(os.environ.get("PEER_TICKS", "120"))
W = int(os.environ.get("PEER_W", "20"))
lo, hi = shard_range(N, rank, world)
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U = g["states"], g["inputs"]
full_bank = make_bank(N, seed=5)
bank = ModelBank({k: (v[lo:hi] if np.ndim(v) else v) for k, v in full_bank.items()})
hist = torch.zeros((W, _lib.HIST_ROW), dtype=torch.float32, device=dev)
px = PeerExchange(device=dev)
lb = LookbackLaunch(bank, hist, W, TS, K=10, idx_offset=lo, peer=px)
assert lb.plan.launches == 1
rows = np.zeros((W, _lib.HIST_ROW), dtype=np.float32)


def load(t_end):
    for j, t in enumerate(range(t_end - W + 1, t_end + 1)):
        xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, t], U[:, t], S[:, t + 1]))
        L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, TS, bank.lf_shared, bank.lr_shared,
                                 rows[j].ctypes.data, None)
    hist.copy_(torch.from_numpy(rows))


ok = True
spin = torch.zeros(1 << 24, dtype=torch.float32, device=dev)
gathered = torch.zeros(world, dtype=torch.int64, device=dev)
for i in range(N_TICKS):
    t_end = 300 + 11 * i
    load(t_end)
    if (i % 4 == 1 and rank == 0) or (i % 4 == 3 and rank == world - 1):
        for _ in range(8):
            spin.mul_(1.0001)                                  # skew: this rank enters the tick a few hundred us late
    lb.out.zero_()
    lb.launch()                                                # K1 / K1p + tree merge + NVLink min-loc: one launch
    k_peer = lb.out[0, 0:1].clone()
    local_min = lb.out[0, 1:2].clone()                         # rank-local arg-min key = head of the local top-K
    td.all_gather_into_tensor(gathered, k_peer)
    td.all_reduce(local_min, op=td.ReduceOp.MIN)
    kp, kn, allk = int(k_peer.item()), int(local_min.item()), gathered.cpu().numpy()
    if not (allk == kp).all() or kp != kn or kp == -1:
        ok = False
        print(rank, "MISMATCH tick", i, hex(kp & (2**64 - 1)), hex(kn & (2**64 - 1)), [hex(int(x) & (2**64 - 1)) for x in allk], flush=True)
    if i % 30 == 7:                                            # float64 oracle over the whole bank
        ref = np.mean(orc.window_errors(full_bank, S, U, t_end, W, TS), axis=1)
        e, idx = decode_keys(np.array([kp], dtype=np.int64).view(np.uint64))
        b = int(np.argmin(ref))
        if int(idx[0]) != b or abs(float(e[0]) - ref[b]) > 1e-4 * ref[b]:
            ok = False
            print(rank, "ORACLE MISMATCH tick", i, int(idx[0]), b, float(e[0]), ref[b], flush=True)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def f_nccl():
    lb.launch(use_peer=False)
    td.all_reduce(lb.out[0, :1], op=td.ReduceOp.MIN)


t_peer, t_nccl = timed(lb.launch), timed(f_nccl)
flag = torch.tensor([1 if ok else 0], device=dev)
td.all_reduce(flag, op=td.ReduceOp.MIN)
print("rank %d/%d N=%d W=%d ticks=%d kernel=%s: %s  tick with in-kernel NVLink min-loc %.1f us, with NCCL all-reduce %.1f us" % (
    rank, world, N, W, N_TICKS, lb.kernel_name, "OK" if ok else "FAIL", t_peer, t_nccl), flush=True)
td.barrier()
td.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
