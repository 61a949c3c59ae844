"""Time the BASELINE configs that are not the bench workload and check parity on a bounded subset.
   python tools/run_configs.py c1|c3|c4 [...]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube          # noqa: E402
from oracle import vip_oracle as O         # noqa: E402
import vip_b200                            # noqa: E402
from vip_b200 import _cabi                 # noqa: E402


def timed(fn, reps=2):
    best = None
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return out, best


def c1():
    cube, angs = adi_cube(50, 101, 5, 60.0, seed=20260101)
    fr, dt = timed(lambda: vip_b200.pca(cube, angs, ncomp=5, verbose=False), reps=3)
    t = time.perf_counter()
    ref = O.pca_fullframe(cube, angs, ncomp=5)
    cpu = time.perf_counter() - t
    err = np.max(np.abs(fr - ref)) / np.max(np.abs(ref))
    print(f"C1 50x101x101 ncomp=5: GPU {dt*1e3:.2f} ms ({50/dt:.0f} frames/s) | CPU oracle {cpu:.2f} s "
          f"({50/cpu:.1f} frames/s) | rel err {err:.2e}")


def c3(n=1000, size=512, ncomp=10, asize=32, check_frames=(0, 487, 999)):
    cube, angs = adi_cube(n, size, ncomp, 90.0, seed=20260103)
    n0 = _cabi.launch_count()
    fr, dt = timed(lambda: vip_b200.pca_annular(cube, angs, ncomp=ncomp, asize=asize, verbose=False), reps=3)
    nl = (_cabi.launch_count() - n0) // 3
    print(f"C3 {n}x{size}x{size} pca_annular ncomp={ncomp} asize={asize}: GPU {dt:.3f} s ({n/dt:.0f} frames/s), "
          f"{nl} launches (pageable host cube in, frame out)")
    co, cd, fr = vip_b200.pca_annular(cube, angs, ncomp=ncomp, asize=asize, verbose=False, full_output=True)
    t = time.perf_counter()
    ref = O.pca_annular(cube, angs, ncomp=ncomp, asize=asize, frames=list(check_frames), derotate=False)
    cpu = time.perf_counter() - t
    scale = max(np.max(np.abs(ref[f])) for f in check_frames)
    err = max(np.max(np.abs(co[f] - ref[f])) for f in check_frames) / scale
    print(f"   parity on frames {check_frames} (PCA stage, all annuli): rel err {err:.2e}; CPU oracle "
          f"{cpu/len(check_frames):.1f} s per frame -> {n*cpu/len(check_frames)/3600:.2f} h for the PCA stage alone")
    # the same comparison against the oracle run in float64 (which side is the fp32 error on?)
    ref64 = O.pca_annular(cube.astype(np.float64), angs, ncomp=ncomp, asize=asize, frames=list(check_frames),
                          derotate=False)
    e_ours = max(np.max(np.abs(co[f] - ref64[f])) for f in check_frames) / scale
    e_ref = max(np.max(np.abs(ref[f] - ref64[f])) for f in check_frames) / scale
    print(f"   vs float64 oracle: ours {e_ours:.2e}, fp32 reference algorithm {e_ref:.2e}")


def c4(nframes=32):
    rng = np.random.default_rng(0)
    z, S = 39, 256
    lam = np.linspace(0.95, 1.65, z)
    sl = lam.max() / lam
    base, angs = adi_cube(nframes, S, 10, 60.0, seed=20260104)
    cube = np.empty((z, nframes, S, S), np.float32)
    for c in range(z):
        cube[c] = base * (1.0 + 0.01 * c) + rng.normal(scale=1.0, size=base.shape).astype(np.float32)
    fr, dt = timed(lambda: vip_b200.pca(cube, angs, scale_list=sl, adimsdi="double", ncomp=(3, 10), verbose=False))
    print(f"C4 slice 39x{nframes}x256x256 double PCA (3,10): GPU {dt:.3f} s ({nframes/dt:.1f} ADI frames/s)")
    t = time.perf_counter()
    ref = O.pca_adimsdi_double(cube, angs, sl, (3, 10), frames=[0])
    cpu = time.perf_counter() - t
    print(f"   CPU oracle stage 1: {cpu:.1f} s per ADI frame")


def c5(n=400, size=1024, ncomp=50):
    """BASELINE config 5 geometry (1024x1024 frames -> 4096-point FFT planes, randomized SVD, ncomp=50) on ONE GPU
    with a reduced number of frames (the full config is 4000 frames over 8 GPUs)."""
    cube, angs = adi_cube(n, size, 20, 90.0, seed=20260105)
    np.random.seed(7)
    fr, dt = timed(lambda: vip_b200.pca(cube, angs, ncomp=ncomp, svd_mode="randsvd", verbose=False), reps=2)
    print(f"C5 slice {n}x{size}x{size} randsvd ncomp={ncomp}: GPU {dt:.3f} s ({n/dt:.0f} frames/s)")
    fr2, dt2 = timed(lambda: vip_b200.pca(cube, angs, ncomp=20, verbose=False), reps=2)
    print(f"   same cube, exact PCA ncomp=20: GPU {dt2:.3f} s ({n/dt2:.0f} frames/s)")
    # derotation parity on two frames at this size (oracle: ~9 s per frame)
    sub = np.ascontiguousarray(cube[:2] - cube[:2].mean(0))
    t = time.perf_counter()
    ref = O.cube_derotate(sub, angs[:2])
    cpu = time.perf_counter() - t
    out = vip_b200.cube_derotate(sub, angs[:2])
    err = np.max(np.abs(out - ref)) / np.max(np.abs(ref))
    print(f"   derotation 1024x1024 vs oracle: rel err {err:.2e} (CPU {cpu/2:.1f} s per frame)")


if __name__ == "__main__":
    for name in sys.argv[1:]:
        {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[name]()
