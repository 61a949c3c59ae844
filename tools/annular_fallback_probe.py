"""Measured accuracy of the frame-by-frame fallback of pca_annular (ncomp > 24 or libraries > 256 frames) against the
oracle on the float64-cast cube AND of the oracle's own fp32 run against it (one-off probe, third session)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube           # noqa: E402
from oracle import vip_oracle as O         # noqa: E402
import vip_b200 as vb                      # noqa: E402


def probe(tag, cube, angs, **kw):
    co, cd, fr = vb.pca_annular(cube, angs, verbose=False, full_output=True, **kw)
    oo, od, of = O.pca_annular(cube.astype(np.float64), angs, full_output=True, **kw)
    o32 = O.pca_annular(cube, angs, full_output=True, **kw)[0]
    s = np.max(np.abs(oo))
    m = ~np.isnan(of)
    print(f"{tag}: residual cube vs fp64 oracle {np.max(np.abs(co - oo)) / s:.2e} (the oracle's fp32 run: "
          f"{np.max(np.abs(o32 - oo)) / s:.2e}); frame {np.max(np.abs(fr[m] - of[m])) / np.max(np.abs(of[m])):.2e}",
          flush=True)


cube, angs = adi_cube(40, 36, 3, 80.0, seed=9)
probe("ncomp=30, 36-frame libraries", cube, angs, ncomp=30, asize=6, delta_rot=0.05, radius_int=2)
probe("ncomp=26, 36-frame libraries", cube, angs, ncomp=26, asize=6, delta_rot=0.05, radius_int=2)
cube, angs = adi_cube(300, 20, 3, 170.0, seed=10)
probe("ncomp=4, 270-frame libraries", cube, angs, ncomp=4, asize=5, delta_rot=0.1, max_frames_lib=270, radius_int=2)
