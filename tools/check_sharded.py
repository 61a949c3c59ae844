"""Multi-GPU check (run under torchrun on an N-GPU box): sharded PCA == single-GPU pca()."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube                     # noqa: E402
import vip_b200                                       # noqa: E402
from vip_b200.parallel import pca_sharded             # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
cube, angs = adi_cube(203, 128, 10, 80.0, seed=77)        # 203 frames: uneven frame shards
for collapse in ("median", "mean"):
    fr = pca_sharded(cube, angs, 10, collapse=collapse)
    if rank == 0:
        ref = vip_b200.pca(cube, angs, ncomp=10, collapse=collapse, verbose=False)
        err = float(np.max(np.abs(fr - ref)) / np.max(np.abs(ref)))
        print(f"sharded world={world} collapse={collapse}: rel err vs single-GPU pca = {err:.2e}")
        assert err < 1e-5, err
# BASELINE config 5's mode: randomized SVD on pixel shards (sketch all-reduces) == single-GPU randsvd, same omega
np.random.seed(5)
fr = pca_sharded(cube, angs, 10, svd_mode="randsvd")
if rank == 0:
    np.random.seed(5)
    ref = vip_b200.pca(cube, angs, ncomp=10, svd_mode="randsvd", verbose=False)
    err = float(np.max(np.abs(fr - ref)) / np.max(np.abs(ref)))
    print(f"sharded world={world} randsvd: rel err vs single-GPU pca(randsvd) = {err:.2e}")
    assert err < 1e-4, err
# BASELINE config 4's mode: ADI+mSDI double PCA sharded by ADI frame == single-GPU pca(adimsdi='double')
from tools.make_golden import ifs_cube                                   # noqa: E402
from vip_b200.parallel import pca_adimsdi_double_sharded                  # noqa: E402
cube4, angs4, sl = ifs_cube(z=6, n=21, size=64, seed=3)
for collapse in ("median", "mean"):
    fr = pca_adimsdi_double_sharded(cube4, angs4, sl, (2, 3), collapse=collapse)
    if rank == 0:
        ref = vip_b200.pca(cube4, angs4, scale_list=sl, adimsdi="double", ncomp=(2, 3), collapse=collapse,
                           verbose=False)
        err = float(np.max(np.abs(fr - ref)) / np.max(np.abs(ref)))
        print(f"sharded world={world} ADI+mSDI double collapse={collapse}: rel err vs single-GPU pca = {err:.2e}")
        assert err < 1e-4, err
dist.barrier()
dist.destroy_process_group()
