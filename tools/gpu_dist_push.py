"""Multi-GPU check (torchrun, >= 2 GPUs): LookBack.push with a sharded bank must return, on every rank, the same
arg-min / top-10 / best error as the float64 oracle on the whole bank -- with the NVLink peer-memory finalist gather
(default) and with the NCCL all-gather (LLAMPC_PEER_GATHER=0).  Prints OK/FAIL and the push latency per rank."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llampc_b200.dist import shard_range                   # noqa: E402
from llampc_b200.mpc import LookBack                        # noqa: E402
from oracle import llampc_oracle as orc                     # noqa: E402  (checker)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "ethz_history.npz"))
S, U, Ts = g["states"], g["inputs"], float(g["Ts"])
N, W, K = 6000, 10, 10
bank = orc.make_bank(N, seed=17)
lo, hi = shard_range(N, rank, world)
shard = {k: (bank[k][lo:hi] if np.ndim(bank[k]) else bank[k]) for k in orc.PARAM_NAMES}
ok = True
for mode in ("recompute", "rolling"):
    lb = LookBack(shard, W=W, Ts=Ts, K=K, refine=16, idx_offset=lo, group=td.group.WORLD, mode=mode)
    ref = orc.LookBackOracle(bank, W, Ts, K)
    lat = []
    for t in range(400, 400 + 4 * W):
        a = time.perf_counter()
        got = lb.push(S[:, t], U[:, t], S[:, t + 1])
        lat.append(time.perf_counter() - a)
        rbest, rtopk, ravg = ref.push(S[:, t], U[:, t], S[:, t + 1])
        if rbest is None:
            ok &= got == (None, None, None)
            continue
        good = got[0] == rbest and list(got[1]) == list(rtopk) and abs(got[2] - ravg[rbest]) <= 1e-9 * ravg[rbest]
        if not good:
            print(rank, mode, "MISMATCH tick", t, got[0], rbest, list(got[1]), list(rtopk), flush=True)
        ok &= good
    print("rank %d/%d %s peer_gather=%s: %s  push p50 %.1f us" % (rank, world, mode, lb._peer is not None, "OK" if ok else "FAIL",
                                                                  np.median(lat[W:]) * 1e6), flush=True)
    del lb
td.barrier()
td.destroy_process_group()
sys.exit(0 if ok else 1)
