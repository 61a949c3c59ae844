"""BASELINE config 3 (1000 x 512 x 512, pca_annular ncomp=10, asize=32): wall time of the whole call (pageable host cube
in, frame out) and the per-section split of VIP_B200_TIMING=1."""
import os
import sys
import time

os.environ["VIP_B200_TIMING"] = "1"
import numpy as np                         # noqa: E402
import torch                               # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube           # noqa: E402
import vip_b200                            # noqa: E402

cube, angs = adi_cube(1000, 512, 10, 90.0, seed=20260103)
for rep in range(3):
    torch.cuda.synchronize()
    t = time.perf_counter()
    fr = vip_b200.pca_annular(cube, angs, ncomp=10, asize=32, verbose=False)
    torch.cuda.synchronize()
    print(f"C3 call {rep}: {time.perf_counter() - t:.3f} s", flush=True)
