"""Strong scaling of ONE BASELINE-config-5-shaped cube (1024x1024 frames -> 4096-point FFT planes, randomized SVD,
ncomp=50) over the GPUs of one box: the workload the north star's multi-GPU target is stated on, with a reduced
number of frames so that it also fits one GPU.   Run under torchrun (any world size, 1 included):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_c5.py [frames]

Prints one line: device-resident and end-to-end (pinned host cube -> frame on rank 0) milliseconds, max over ranks.
The cube is generated once (rank 0) and cached in /dev/shm for the following runs on the same box."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube, pa_track             # noqa: E402
from vip_b200 import kernels                            # noqa: E402
from vip_b200.parallel import pca_sharded, shard_bounds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
size, k = 1024, 50
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
path = f"/dev/shm/vipb200_c5_{n}.npy"
if rank == 0 and not os.path.exists(path):
    t0 = time.perf_counter()
    cube, _ = adi_cube(n, size, 20, 90.0, seed=20260105)
    np.save(path, cube)
    print(f"generated {n}x{size}x{size} in {time.perf_counter() - t0:.0f} s", flush=True)
    del cube
dist.barrier()
angs = pa_track(np.random.default_rng(20260105), n, 90.0)      # the PA track adi_cube draws first from this seed
pinned = torch.from_numpy(np.load(path, mmap_mode="r")[:]).pin_memory()
cube = pinned.numpy()
pb = shard_bounds(size * size, world)
shard = kernels.upload_columns(cube.reshape(n, -1), int(pb[rank]), int(pb[rank + 1]), torch.device("cuda", local))


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
    return pca_sharded(cube, angs, k, resident_shard=shard, svd_mode="randsvd", random_state=7)


def e2e():
    return pca_sharded(cube, angs, k, svd_mode="randsvd", random_state=7)


for _ in range(2):
    resident()
ms_res = timed(resident, 3)
e2e()
ms_e2e = timed(e2e, 3)
if rank == 0:
    print(f"C5 slice {n}x{size}x{size} randsvd ncomp={k}, world={world}: resident {ms_res:.1f} ms "
          f"({n / ms_res * 1e3:.0f} frames/s), e2e {ms_e2e:.1f} ms ({n / ms_e2e * 1e3:.0f} frames/s)", flush=True)
dist.destroy_process_group()
