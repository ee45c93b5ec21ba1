"""World-size-2 gloo tests (CPU) of the multi-GPU exchange step: packed-key min-loc all-reduce and the
finalist all-gather must give the same answer as a single-process argmin / argsort."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from llampc_b200 import dist as lldist
from llampc_b200.mpc.lookback import decode_keys


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(err32, idx):
    return (err32.astype(np.float32).view(np.uint32).astype(np.uint64) << np.uint64(32)) | idx.astype(np.uint64)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(123)
    n = 1000
    err = rng.rand(n).astype(np.float32)
    err[[10, 700]] = err.min() / 2                      # a tie across the two shards -> lowest index wins
    lo, hi = lldist.shard_range(n, rank, world)
    keys = _pack(err[lo:hi], np.arange(lo, hi))
    local_best = torch.tensor([int(keys.min())], dtype=torch.int64)
    g = lldist.minloc_allreduce(local_best, None)
    e, i = decode_keys(np.array([g], dtype=np.uint64))
    order = np.argsort(keys)[:10]
    sc, ix = lldist.gather_finalists(err[lo:hi][order].astype(np.float64), np.arange(lo, hi)[order], None, torch.device("cpu"))
    merged = ix[np.lexsort((ix, sc))][:10]
    q.put((rank, int(i[0]), float(e[0]), merged.tolist(), int(np.argmin(err)), np.argsort(err, kind="stable")[:10].tolist()))
    td.destroy_process_group()


def test_minloc_and_finalists_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, best, e, merged, ref_best, ref_top in res:
        assert best == ref_best == 10
        assert merged == ref_top


def test_shard_range_covers_everything():
    for n, w in ((10, 3), (1 << 20, 8), (5, 8)):
        spans = [lldist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
