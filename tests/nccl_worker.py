"""Worker of tests/test_gpu_multirank.py (run under ``python -m torch.distributed.run``, one rank per GPU):
the sharded paths of vip_b200/parallel.py over NCCL against the single-GPU result computed on rank 0.
Writes the measured relative errors as JSON to the path given as argv[1] (rank 0)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.synth import adi_cube                                            # noqa: E402
from tools.make_golden import ifs_cube                                      # noqa: E402
import vip_b200                                                              # noqa: E402
import vip_b200.parallel                                                     # noqa: E402
from vip_b200.parallel import pca_sharded, pca_adimsdi_double_sharded        # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main(out_path):
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    res = {"world": world}
    cube, angs = adi_cube(203, 128, 10, 80.0, seed=77)           # 203 frames: uneven frame shards
    for collapse in ("median", "mean", "sum", "max"):
        fr = pca_sharded(cube, angs, 10, collapse=collapse)
        if rank == 0:
            res[f"exact_{collapse}"] = rel(fr, vip_b200.pca(cube, angs, ncomp=10, collapse=collapse, verbose=False))
    # the three exchange variants: NCCL all-to-alls (subtract first / raw cube overlapped with the eigensolver) and the
    # exchanges fused into the kernels over peer memory (default when symmetric memory is available)
    os.environ["VIP_B200_SHARD_FUSED"] = "0"
    for overlap in (False, True):
        fr = pca_sharded(cube, angs, 10, overlap_exchange=overlap)
        if rank == 0:
            res[f"exact_overlap_{int(overlap)}"] = rel(fr, vip_b200.pca(cube, angs, ncomp=10, verbose=False))
    cube2, angs2 = adi_cube(130, 512, 8, 70.0, seed=78)
    fr = pca_sharded(cube2, angs2, 8)
    if rank == 0:
        ref512 = vip_b200.pca(cube2, angs2, ncomp=8, verbose=False)
        res["exact_512_nccl"] = rel(fr, ref512)
    os.environ["VIP_B200_SHARD_FUSED"] = "1"
    from vip_b200.parallel import PeerExchange
    res["fused_available"] = bool(PeerExchange.eligible(vip_b200.parallel.CudaOps(), world, 130, 512, 512, "median", False)
                                  and PeerExchange.get(130, 512, 512, world, rank, vip_b200.parallel.shard_bounds(130, world),
                                                       vip_b200.parallel.shard_bounds(512 * 512, world),
                                                       torch.device("cuda", local), None) is not None)
    for it in range(2):                                   # twice: the second step reuses the peer buffers
        fr = pca_sharded(cube2, angs2, 8)
    if rank == 0:
        res["exact_512"] = rel(fr, ref512)
    fr = pca_sharded(cube2, angs2, 8, collapse="mean")
    if rank == 0:
        res["exact_512_mean"] = rel(fr, vip_b200.pca(cube2, angs2, ncomp=8, collapse="mean", verbose=False))
    # BASELINE config 5's mode: randomized SVD on pixel shards (sketch all-reduces), same omega
    fr = pca_sharded(cube, angs, 10, svd_mode="randsvd", random_state=5)
    if rank == 0:
        ref = vip_b200.psfsub.pca_fullfr._adi_rdi_pca_device(cube, None, angs, 10, None, None, "randsvd", "median",
                                                             False, False, random_state=5)
        res["randsvd"] = rel(fr, ref.cpu().numpy())
    # BASELINE config 4's mode: ADI+mSDI double PCA sharded by ADI frame
    cube4, angs4, sl = ifs_cube(z=6, n=21, size=64, seed=3)
    for collapse in ("median", "mean"):
        fr = pca_adimsdi_double_sharded(cube4, angs4, sl, (2, 3), collapse=collapse)
        if rank == 0:
            ref = vip_b200.pca(cube4, angs4, scale_list=sl, adimsdi="double", ncomp=(2, 3), collapse=collapse,
                               verbose=False)
            res[f"sdi_double_{collapse}"] = rel(fr, ref)
    dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
        print("nccl_worker:", json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
