"""BASELINE config 4 at FULL size -- IFS cube 39 x 300 x 256 x 256 fp32 (3.07 GB), ADI+mSDI double PCA ncomp=(3, 10) --
sharded BY ADI FRAME over the GPUs of one box (`vip_b200.parallel.pca_adimsdi_double_sharded`), or on one GPU through
the public call `vip_b200.pca(..., scale_list=, adimsdi='double')` when run without torchrun.

    python tools/scale_c4.py [frames]                                                  # 1 GPU, public API
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_c4.py [frames]

Every rank synthesises only ITS ADI frames of the 4-d cube (the pages of the other frames are never touched), so the
host work per rank shrinks with N like the GPU work.  Prints one JSON line (rank 0): milliseconds per cube (max over
ranks, host cube -> final frame, i.e. end to end: the stage-1 upload is part of the path), ADI frames/s, and a
fingerprint of the frame; the frame is written to gpurun_out/c4_frame_world{N}.npy for the cross-N comparison."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.synth import adi_cube                          # noqa: E402
import vip_b200                                            # noqa: E402
from vip_b200.parallel import pca_adimsdi_double_sharded, shard_bounds   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
z, S, ncomp = 39, 256, (3, 10)
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lam = np.linspace(0.95, 1.65, z)
scale_list = lam.max() / lam
base, angs = adi_cube(n, S, 10, 60.0, seed=20260104)
fb = shard_bounds(n, world)
f0, f1 = int(fb[rank]), int(fb[rank + 1])
cube = np.empty((z, n, S, S), dtype=np.float32)              # only the own frames are ever touched
for f in range(f0, f1):
    rng = np.random.default_rng(1000 + f)                    # per-frame seed: the cube does not depend on N
    noise = rng.normal(scale=1.0, size=(z, S, S)).astype(np.float32)
    cube[:, f] = base[f][None] * (1.0 + 0.01 * np.arange(z, dtype=np.float32))[:, None, None] + noise


def run():
    if world == 1:
        return vip_b200.pca(cube, angs, scale_list=scale_list, adimsdi="double", ncomp=ncomp, verbose=False)
    return pca_adimsdi_double_sharded(cube, angs, scale_list, ncomp)


def timed(steps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, out


run()
ms, frame = timed(2)
if rank == 0:
    frame = np.asarray(frame, dtype=np.float32)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    np.save(os.path.join(out, f"c4_frame_world{world}_n{n}.npy"), frame)
    print("C4FULL " + json.dumps({"config": f"C4 {z}x{n}x{S}x{S} ADI+mSDI double PCA ncomp={ncomp}", "world": world,
                                  "cube_GB": z * n * S * S * 4 / 1e9, "e2e_ms": ms, "adi_frames_per_s": n / ms * 1e3,
                                  "timing": "host wall clock around synchronised calls (the call ends with a D2H of "
                                            "the frame), max over ranks, 2 steps after 1 warm-up",
                                  "frame_fingerprint": [float(np.nanmax(np.abs(frame))),
                                                        float(np.nansum(frame.astype(np.float64)))]}), flush=True)
if world > 1:
    dist.destroy_process_group()
