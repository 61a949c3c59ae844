"""BASELINE config 5 at (up to) FULL size -- 4000 x 1024 x 1024 fp32 = 16.8 GB, randomized SVD ncomp=50 -- as ONE cube
sharded over the GPUs of one box (the workload of the north star's multi-GPU target).  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_c5_full.py [frames]

No host ever holds the whole cube: every rank synthesises ITS pixel shard on its own GPU (same halo / speckle modes
/ temporal coefficients on every rank from a shared seed, rank-seeded read noise), copies it to a pinned host shard
for the end-to-end arm and hands `pca_sharded` that shard (`host_shard=`, `shape=`).  Generation takes about a
second per rank instead of minutes of numpy on the box's host cores (8-GPU box time is charged 8x).
Prints one line: device-resident and end-to-end milliseconds (max over ranks) and frames/s.
Written in round 1 after the GPU minutes were spent: not yet run on hardware."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import pa_track                          # noqa: E402
from vip_b200.parallel import pca_sharded, shard_bounds   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
size, k, K = 1024, 50, 20
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
p = size * size
pb = shard_bounds(p, world)
p0, p1 = int(pb[rank]), int(pb[rank + 1])

# ---- synthetic cube, this rank's pixel columns only (SURVEY 8d recipe: halo + K speckle modes with AR(1) temporal
# coefficients + read noise; the companion is left out, it does not change the cost)
rng = np.random.default_rng(20260105)
angs = pa_track(rng, n, 90.0)
ar = np.empty((n, K))
ar[0] = rng.standard_normal(K)
for t in range(1, n):
    ar[t] = 0.9 * ar[t - 1] + np.sqrt(1 - 0.81) * rng.standard_normal(K)
coef = torch.from_numpy((0.05 * (0.88 ** np.arange(K))[None, :] * (0.5 + ar)).astype(np.float32)).to(dev)
g_all = torch.Generator(device=dev).manual_seed(1234)              # identical on every rank
yy, xx = torch.meshgrid(torch.arange(size, device=dev), torch.arange(size, device=dev), indexing="ij")
halo = 1e4 / (1.0 + ((yy - size // 2) ** 2 + (xx - size // 2) ** 2).float() / 16.0)
fy = torch.fft.fftfreq(size, device=dev)[:, None]
fx = torch.fft.rfftfreq(size, device=dev)[None, :]
filt = torch.exp(-2.0 * (np.pi * 2.0) ** 2 * (fy ** 2 + fx ** 2))
modes = torch.fft.irfft2(torch.fft.rfft2(torch.randn((K, size, size), device=dev, generator=g_all)) * filt,
                         s=(size, size))
modes = modes / modes.std(dim=(1, 2), keepdim=True) * halo
modes_g = modes.reshape(K, p)[:, p0:p1].contiguous()
halo_g = halo.reshape(p)[p0:p1].contiguous()
del modes, halo, yy, xx
g_rank = torch.Generator(device=dev).manual_seed(99 + rank)
shard = torch.empty((n, p1 - p0), dtype=torch.float32, device=dev)
for s0 in range(0, n, 256):                                        # chunked: bounded temporaries
    s1 = min(n, s0 + 256)
    blk = coef[s0:s1] @ modes_g
    blk += halo_g[None]
    blk += 3.0 * torch.randn(blk.shape, device=dev, generator=g_rank)
    shard[s0:s1] = blk
del blk
host = torch.empty((n, p1 - p0), dtype=torch.float32).pin_memory()
host.copy_(shard)
host_np = host.numpy()
torch.cuda.synchronize()
shape = (n, size, size)


def timed(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def resident():
    return pca_sharded(None, angs, k, resident_shard=shard, shape=shape, svd_mode="randsvd", random_state=7)


def e2e():
    return pca_sharded(None, angs, k, host_shard=host_np, shape=shape, svd_mode="randsvd", random_state=7)


resident()
ms_res = timed(resident, 2)
frame = e2e()
ms_e2e = timed(e2e, 2)
if rank == 0:
    assert frame.shape == (size, size) and np.isfinite(frame).all()
    print(f"C5 {n}x{size}x{size} randsvd ncomp={k}, world={world}: resident {ms_res:.1f} ms "
          f"({n / ms_res * 1e3:.0f} frames/s), e2e {ms_e2e:.1f} ms ({n / ms_e2e * 1e3:.0f} frames/s), "
          f"{n * p * 4 / 1e9:.1f} GB cube", flush=True)
dist.destroy_process_group()
