"""Probe (run under torchrun on >= 2 GPUs): is torch's symmetric memory usable on this box, and how fast is a plain
peer write over NVLink compared with NCCL all_to_all_single?"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 64 * 1024 * 1024                      # 256 MB of fp32 per rank
try:
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(f"rank {rank}: rendezvous ok, world {hdl.world_size}, ptrs {len(hdl.buffer_ptrs)}", flush=True)
    src = torch.full((n // world,), float(rank + 1), device=dev)
    hdl.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        torch.cuda.synchronize()
        e0.record()
        for peer in range(world):
            buf = hdl.get_buffer(peer, (n,), torch.float32)
            buf[rank * (n // world):(rank + 1) * (n // world)].copy_(src)          # peer write
        e1.record()
        hdl.barrier()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok = all(abs(float(t[r * (n // world)]) - (r + 1)) < 1e-6 for r in range(world))
    print(f"rank {rank}: peer writes of {n * 4 / 1e6:.0f} MB total in {ms:.3f} ms = {n * 4 / ms / 1e6:.0f} GB/s, data ok {ok}",
          flush=True)
    # NCCL all_to_all_single of the same volume
    a = torch.empty(n, device=dev)
    b = torch.empty(n, device=dev)
    for it in range(3):
        torch.cuda.synchronize()
        e0.record()
        dist.all_to_all_single(b, a)
        e1.record()
        torch.cuda.synchronize()
    print(f"rank {rank}: NCCL all_to_all_single {n * 4 / 1e6:.0f} MB in {e0.elapsed_time(e1):.3f} ms = "
          f"{n * 4 / e0.elapsed_time(e1) / 1e6:.0f} GB/s", flush=True)
except Exception as exc:                                                     # noqa: BLE001
    print(f"rank {rank}: symmetric memory FAILED: {exc!r}", flush=True)
dist.barrier()
dist.destroy_process_group()
