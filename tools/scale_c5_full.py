"""BASELINE config 5 at (up to) FULL size -- 4000 x 1024 x 1024 fp32 = 16.8 GB, randomized SVD ncomp=50 -- as ONE cube
sharded over the GPUs of one box (the workload of the north star's multi-GPU target).  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_c5_full.py [frames]

No host ever holds the whole cube: every rank synthesises ITS pixel shard on its own GPU (same halo / speckle modes
/ temporal coefficients on every rank from a shared seed; the read noise is drawn per 1/8-of-the-frame pixel block
from a block-seeded generator, so the cube is THE SAME for world sizes 1, 2, 4 and 8), copies it to a pinned host
shard for the end-to-end arm and hands `pca_sharded` that shard (`host_shard=`, `shape=`).
Prints one JSON line (rank 0): device-resident and end-to-end milliseconds (max over ranks), frames/s, the
per-stage split (CUDA events, max over ranks) and a fingerprint of the final frame; the frame itself is written to
gpurun_out/c5_frame_world{N}.npy so that the runs at different N can be compared with each other
(tools/compare_c5_frames.py -> parity of the sharded result against the single-GPU one at full size)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.synth import pa_track                          # noqa: E402
from vip_b200.parallel import pca_sharded, shard_bounds, StageTimer   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
size, k, K, NBLK = 1024, 50, 50, 8        # K = ncomp speckle modes: spectrum gapped at ncomp
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
assert NBLK % world == 0, "world size must divide 8"
p = size * size
pb = shard_bounds(p, world)
p0, p1 = int(pb[rank]), int(pb[rank + 1])
bw = p // NBLK

# ---- synthetic cube, this rank's pixel columns only (SURVEY 8d recipe: halo + K speckle modes with AR(1) temporal
# coefficients + read noise; the companion is left out, it does not change the cost)
rng = np.random.default_rng(20260105)
angs = pa_track(rng, n, 90.0)
ar = np.empty((n, K))
ar[0] = rng.standard_normal(K)
for t in range(1, n):
    ar[t] = 0.9 * ar[t - 1] + np.sqrt(1 - 0.81) * rng.standard_normal(K)
coef = torch.from_numpy((0.05 * (0.97 ** np.arange(K))[None, :] * (0.5 + ar)).astype(np.float32)).to(dev)
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
blocks = list(range(p0 // bw, p1 // bw))
gens = {b: torch.Generator(device=dev).manual_seed(99 + b) for b in blocks}
shard = torch.empty((n, p1 - p0), dtype=torch.float32, device=dev)
for s0 in range(0, n, 256):                                        # chunked: bounded temporaries
    s1 = min(n, s0 + 256)
    blk = coef[s0:s1] @ modes_g
    blk += halo_g[None]
    for b in blocks:
        c0 = b * bw - p0
        blk[:, c0:c0 + bw] += 3.0 * torch.randn((s1 - s0, bw), device=dev, generator=gens[b])
    shard[s0:s1] = blk
del blk
from vip_b200._device import gpu_local_cpus                # noqa: E402
with gpu_local_cpus(local):                                # pinned pages on the NUMA node of this rank's GPU
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


def resident(timer=None):
    return pca_sharded(None, angs, k, resident_shard=shard, shape=shape, svd_mode="randsvd", random_state=7,
                       timer=timer)


def e2e():
    return pca_sharded(None, angs, k, host_shard=host_np, shape=shape, svd_mode="randsvd", random_state=7)


resident()
ms_res = timed(resident, 2)
frame = e2e()
ms_e2e = timed(e2e, 2)
dist.barrier()
tm = StageTimer(dev)
resident(tm)
st = tm.summary()
names = sorted(st)
t = torch.tensor([st[nm] for nm in names], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
stage = dict(zip(names, [round(float(v), 3) for v in t.tolist()]))
if rank == 0:
    assert frame.shape == (size, size) and np.isfinite(frame).all()
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    np.save(os.path.join(out, f"c5_frame_world{world}_n{n}.npy"), frame.astype(np.float32))
    line = {"config": f"C5 {n}x{size}x{size} randsvd ncomp={k}", "world": world, "cube_GB": n * p * 4 / 1e9,
            "resident_ms": ms_res, "resident_frames_per_s": n / ms_res * 1e3, "e2e_ms": ms_e2e,
            "e2e_frames_per_s": n / ms_e2e * 1e3, "stage_ms": stage,
            "frame_fingerprint": [float(np.abs(frame).max()), float(frame.astype(np.float64).sum()),
                                  float(frame[300, 700])]}
    print("C5FULL " + json.dumps(line), flush=True)
dist.destroy_process_group()
