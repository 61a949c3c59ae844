"""One BASELINE config-4 slice (39 x N x 256 x 256, ADI+mSDI double PCA ncomp=(3, 10)) for an ncu launch list / wall time."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube           # noqa: E402
import vip_b200                            # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
z, S = 39, 256
rng = np.random.default_rng(0)
lam = np.linspace(0.95, 1.65, z)
sl = lam.max() / lam
base, angs = adi_cube(n, S, 10, 60.0, seed=20260104)
cube = np.empty((z, n, S, S), np.float32)
for c in range(z):
    cube[c] = base * (1.0 + 0.01 * c) + rng.normal(scale=1.0, size=base.shape).astype(np.float32)
for rep in range(reps):
    torch.cuda.synchronize()
    t = time.perf_counter()
    fr = vip_b200.pca(cube, angs, scale_list=sl, adimsdi="double", ncomp=(3, 10), verbose=False)
    torch.cuda.synchronize()
    print(f"C4 slice 39x{n}x256x256 call {rep}: {time.perf_counter() - t:.3f} s", flush=True)
