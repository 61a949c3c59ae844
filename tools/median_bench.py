"""Median collapse of a derotated-cube-like stack (500 x 512 x 512, NaN corners) timed with CUDA events, for the
algorithms selectable with VIP_B200_MEDIAN_ALGO (default: warp-per-pixel bracket search; radix: 4-bit radix kernel)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vip_b200 import kernels              # noqa: E402

n, S = int(sys.argv[1]) if len(sys.argv) > 1 else 500, 512
g = torch.Generator(device="cuda").manual_seed(1)
cube = torch.randn((n, S * S), device="cuda", generator=g) * 5.0
yy, xx = np.mgrid[:S, :S]
corner = torch.from_numpy((np.hypot(yy - S / 2, xx - S / 2) > S / 2 * 1.3).reshape(-1)).cuda()
cube[:n // 2, corner] = float("nan")                      # rotated corners: NaN in half of the frames
reps = int(os.environ.get("REPS", "20"))
for algo in (os.environ.get("ALGOS", "warp,radix")).split(","):
    if algo == "warp":
        os.environ.pop("VIP_B200_MEDIAN_ALGO", None)
    else:
        os.environ["VIP_B200_MEDIAN_ALGO"] = algo
    out = kernels.collapse(cube, "median")
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        out = kernels.collapse(cube, "median")
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    print(f"median {n}x{S}x{S} algo={algo}: {ms:.4f} ms = {4.0 * n * S * S / ms / 1e6:.0f} GB/s", flush=True)
