#!/usr/bin/env python3
"""Probe: can two ranks map each other's device buffers (torch symmetric memory) and store into them?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=f"cuda:{local}")
t.zero_()
h = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "backend", symm_mem.get_backend(torch.device(f"cuda:{local}")) if hasattr(symm_mem, "get_backend") else "?",
      "ptrs", [hex(p) for p in h.buffer_ptrs], "rank/world", h.rank, h.world_size, flush=True)
peer = h.get_buffer((rank + 1) % world, (16,), torch.uint8)
peer.fill_(rank + 1)
torch.cuda.synchronize()
dist.barrier()
print(rank, "my first bytes now", t[:4].tolist(), flush=True)
dist.destroy_process_group()
