"""Multi-GPU check + timing (torchrun, >= 2 GPUs): the in-kernel NVLink min-loc (llampc_lookback_window_topk_peer_f32)
must give the same global key as K1 + NCCL MIN all-reduce, on every rank, for many consecutive ticks.
Prints one line per rank: OK/FAIL and the per-tick device time of both variants."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200 import _lib                               # noqa: E402
from llampc_b200.dist import PeerExchange, shard_range     # noqa: E402
from llampc_b200.mpc import LookBack                        # noqa: E402
from bench import make_bank, W_C2, TS                       # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
L = _lib.lib()
N = int(os.environ.get("PEER_N", str(131072 * world)))
lo, hi = shard_range(N, rank, world)
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U = g["states"], g["inputs"]
lb = LookBack(make_bank(N, seed=5, lo=lo, hi=hi), W=W_C2, Ts=TS, K=10, refine=0, idx_offset=lo)
px = PeerExchange(device=dev)
st = torch.cuda.current_stream().cuda_stream
ticket = torch.zeros(1, dtype=torch.int32, device=dev)
out_peer = torch.zeros(17, dtype=torch.int64, device=dev)
n = hi - lo
ok = True
for i, t_end in enumerate(range(700, 760, 3)):
    ts = np.arange(t_end - W_C2 + 1, t_end + 1)
    lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    rc = L.llampc_lookback_window_topk_peer_f32(lb.bank.packed.data_ptr(), n, lb.bank.Npad, lb.hist.data_ptr(), W_C2, W_C2, TS,
                                                lb.avg_err.data_ptr(), lb.best_key.data_ptr(), lb.cta_lists.data_ptr(), lo,
                                                int(lb.bank.geom_shared), lb.split, 10, ticket.data_ptr(), out_peer.data_ptr(),
                                                px.peer_ptrs.data_ptr(), world, rank, px.next_seq(), st)
    _lib.check(rc, "peer")
    k_peer = int(out_peer[0].item())
    ref = out_peer[1:2].clone()                              # local arg-min key = first of the local top-K
    td.all_reduce(ref, op=td.ReduceOp.MIN)
    if k_peer != int(ref.item()) or k_peer == 0:
        ok = False
        print(rank, "MISMATCH tick", i, hex(k_peer), hex(int(ref.item())), flush=True)


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


def f_peer():
    L.llampc_lookback_window_topk_peer_f32(lb.bank.packed.data_ptr(), n, lb.bank.Npad, lb.hist.data_ptr(), W_C2, W_C2, TS,
                                           lb.avg_err.data_ptr(), lb.best_key.data_ptr(), lb.cta_lists.data_ptr(), lo,
                                           int(lb.bank.geom_shared), lb.split, 10, ticket.data_ptr(), out_peer.data_ptr(),
                                           px.peer_ptrs.data_ptr(), world, rank, px.next_seq(), st)


def f_nccl():
    L.llampc_lookback_window_topk_f32(lb.bank.packed.data_ptr(), n, lb.bank.Npad, lb.hist.data_ptr(), W_C2, 1, W_C2, TS,
                                      lb.avg_err.data_ptr(), lb.best_key.data_ptr(), lb.cta_lists.data_ptr(), lo,
                                      int(lb.bank.geom_shared), lb.split, 10, ticket.data_ptr(), lb.result.data_ptr(), st)
    td.all_reduce(lb.result[:1], op=td.ReduceOp.MIN)


t_peer, t_nccl = timed(f_peer), timed(f_nccl)
print("rank %d/%d N=%d: %s  tick with in-kernel NVLink min-loc %.1f us, with NCCL all-reduce %.1f us" % (
    rank, world, N, "OK" if ok else "FAIL", t_peer, t_nccl), flush=True)
td.barrier()
td.destroy_process_group()
sys.exit(0 if ok else 1)
