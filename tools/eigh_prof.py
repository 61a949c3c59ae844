"""Phase timestamps of the second-generation fused eigensolver (VIP_B200_TOPK_PROF=1 prints them from the kernel)
on the BASELINE config-2 Gramian (500 frames of 512 x 512, k = 20)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube          # noqa: E402
from vip_b200 import kernels              # noqa: E402

cube, _ = adi_cube(500, 512, 20, 90.0, seed=20260102)
M = torch.from_numpy(cube.reshape(500, -1)).cuda()
G = kernels.gram(M)
for mode in ("1", "2"):
    os.environ["VIP_B200_TOPK_FUSED"] = mode
    os.environ["VIP_B200_TOPK_PROF"] = "0"
    for _ in range(3):
        kernels.eigh_topk_async(G, 20)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        kernels.eigh_topk_async(G, 20)
    t1.record()
    torch.cuda.synchronize()
    ev, E, info = kernels.eigh_topk(G, 20)
    print(f"config-2 Gramian, fused={mode}: {t0.elapsed_time(t1) / 20:.4f} ms, {info}", flush=True)
os.environ["VIP_B200_TOPK_PROF"] = "1"
kernels.eigh_topk(G, 20)
torch.cuda.synchronize()
