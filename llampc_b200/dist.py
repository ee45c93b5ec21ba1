"""Multi-GPU exchange step of the look-back: candidates are sharded by contiguous index range, every rank
reduces locally, and ONE small collective merges the results.

NCCL has no MINLOC: the packed key (float_bits(err) << 32 | global_index) is reduced with MIN as a signed
64-bit integer, which is exact because scores are non-negative (sign bit clear) and the low word gives
np.argmin's first-index tie-break.  The same functions run on the gloo backend with CPU tensors (tests).
"""
import numpy as np


def _backend_device(group, cuda_device):
    import torch
    import torch.distributed as td
    return cuda_device if td.get_backend(group) == "nccl" else torch.device("cpu")


def minloc_allreduce(key, group=None):
    """key: 1-element int64 tensor holding the local packed key.  In-place MIN all-reduce; returns the
    global key as a Python int.  One collective of 8 bytes."""
    import torch.distributed as td
    dev = _backend_device(group, key.device)
    buf = key if key.device == dev else key.to(dev)
    td.all_reduce(buf, op=td.ReduceOp.MIN, group=group)
    if buf is not key:
        key.copy_(buf)
    return int(buf.item()) & 0xFFFFFFFFFFFFFFFF


def gather_finalists(scores, idx, group, cuda_device):
    """All-gather every rank's finalists (fp64 score, global index) in one collective of 16*Kt bytes/rank."""
    import torch
    import torch.distributed as td
    world = td.get_world_size(group)
    dev = _backend_device(group, cuda_device)
    local = np.empty((len(idx), 2), dtype=np.int64)
    local[:, 0] = np.asarray(scores, dtype=np.float64).view(np.int64)
    local[:, 1] = idx
    send = torch.from_numpy(local).to(dev)
    recv = torch.empty((world,) + tuple(send.shape), dtype=torch.int64, device=dev)
    if td.get_backend(group) == "nccl":
        td.all_gather_into_tensor(recv, send, group=group)
    else:
        td.all_gather(list(recv.unbind(0)), send, group=group)
    out = recv.cpu().numpy().reshape(-1, 2)
    return out[:, 0].copy().view(np.float64), out[:, 1].copy()


def shard_range(n, rank, world):
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


class PeerExchange:
    """Symmetric NVLink exchange buffers for the in-kernel min-loc (llampc_lookback_window_topk_peer_f32): one small
    buffer per rank, mapped into every process of the group with torch symmetric memory (CUDA IPC underneath).
    `peer_ptrs` is the device array of the `world` buffer addresses; `next_seq()` is the per-tick sequence number
    (every rank must call it the same number of times)."""

    def __init__(self, group=None, device=None, words=None):
        import torch
        import torch.distributed as td
        import torch.distributed._symmetric_memory as symm_mem
        self.group = td.group.WORLD if group is None else group
        self.world, self.rank = td.get_world_size(self.group), td.get_rank(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # default size: the min-loc layout [2 parities][world][key, sequence]
        self.buf = symm_mem.empty(4 * self.world if words is None else int(words), dtype=torch.int64, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group=self.group)
        self.peer_ptrs = torch.tensor(list(self.handle.buffer_ptrs), dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        td.barrier(self.group)
        self.seq = 0

    def next_seq(self):
        self.seq = self.seq % 0xFFFFFFFF + 1                       # 32-bit on the device, never 0 (the buffers start zeroed)
        return self.seq
