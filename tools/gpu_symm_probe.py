"""Probe (2+ GPUs, torchrun): does torch symmetric memory give peer-mapped device pointers in this environment?"""
import os
import torch
import torch.distributed as td
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm_mem.empty(64, dtype=torch.int64, device=torch.device("cuda", local))
t.fill_(rank + 100)
hdl = symm_mem.rendezvous(t, group=td.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
td.barrier()
peer = hdl.get_buffer((rank + 1) % world, (64,), torch.int64)
print(rank, "peer value", int(peer[0].item()), flush=True)
peer[1] = rank + 1000                       # remote store
torch.cuda.synchronize()
td.barrier()
print(rank, "my slot 1 written by peer:", int(t[1].item()), flush=True)
td.destroy_process_group()
