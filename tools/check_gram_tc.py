"""Bring-up / accuracy / timing check of the tcgen05 Gramian against an fp64 matmul on the GPU.
   python tools/check_gram_tc.py            (run under gpurun)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vip_b200 import kernels  # noqa: E402
from tools.synth import adi_cube  # noqa: E402


def check(M, tag):
    G64 = M.double() @ M.double().T
    scale = torch.sqrt(torch.diag(G64))
    out = {}
    for bk in (64, 32):
        for pieces in (3, 2):
            os.environ["VIP_B200_GRAM_BK"] = str(bk)
            os.environ["VIP_B200_GRAM_PIECES"] = str(pieces)
            G = kernels.gram(M)
            torch.cuda.synchronize()
            err = ((G - G64).abs() / (scale[:, None] * scale[None, :])).max().item()
            errmax = (G - G64).abs().max().item() / G64.abs().max().item()
            # time
            for _ in range(2):
                kernels.gram(M)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                kernels.gram(M)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{tag} BK={bk} pieces={pieces}: cos-normalised err {err:.2e}, max-normalised {errmax:.2e}, {ms:.3f} ms",
                  flush=True)
            out[(bk, pieces)] = (err, ms)
    os.environ["VIP_B200_GRAM_TC"] = "0"
    G = kernels.gram(M)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kernels.gram(M)
    e1.record()
    torch.cuda.synchronize()
    err = ((G - G64).abs() / (scale[:, None] * scale[None, :])).max().item()
    print(f"{tag} fp64 CUDA-core path: err {err:.2e}, {e0.elapsed_time(e1):.3f} ms", flush=True)
    os.environ["VIP_B200_GRAM_TC"] = "1"
    return out


if __name__ == "__main__":
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    which = sys.argv[1:] or ["small", "ragged", "c2"]
    if "small" in which:
        M = torch.randn(130, 130048, device=dev) * 3 + 100.0
        check(M, "130x130048 randn+100")
    if "ragged" in which:
        M = torch.randn(300, 70001, device=dev)
        M[:, ::7] *= 1e3
        check(M, "300x70001 ragged")
    if "c2" in which:
        cube, _ = adi_cube(500, 512, 20, 90.0, seed=20260102)
        M = torch.as_tensor(cube.reshape(500, -1)).to(dev)
        check(M, "C2 500x262144 synthetic ADI")
    if "c3" in which:
        M = torch.randn(1000, 48028, device=dev) + 5
        check(M, "1000x48028")
