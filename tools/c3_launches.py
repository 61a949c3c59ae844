"""One BASELINE config-3 call (1000 x 512 x 512, pca_annular ncomp=10, asize=32) for an ncu launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube           # noqa: E402
import vip_b200                            # noqa: E402

cube, angs = adi_cube(1000, 512, 10, 90.0, seed=20260103)
fr = vip_b200.pca_annular(cube, angs, ncomp=10, asize=32, verbose=False)
torch.cuda.synchronize()
