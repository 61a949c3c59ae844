"""Micro-benchmark of single stages on the GPU (development tool).
   python tools/bench_stage.py derotate [n] [size]     (env: VIP_B200_DEROT_SCRATCH, VIP_B200_FFT_NT)
   python tools/bench_stage.py eigh [n] [k]
   python tools/bench_stage.py gram|median|proj [n] [size]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vip_b200 import kernels, _cabi                                   # noqa: E402
from vip_b200.preproc.derotation import derotate_device                # noqa: E402
from vip_b200.preproc.subsampling import collapse_device               # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    what = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    g = torch.Generator(device="cuda").manual_seed(1)
    lib = _cabi.lib()
    if what == "derotate":
        cube = torch.randn((n, size, size), device="cuda", generator=g)
        amax = float(os.environ.get("BENCH_ANGLE_MAX", "93.0"))
        angs = np.linspace(3.0, amax, n)
        lib.vb_profile_enable(0)
        ms = timeit(lambda: derotate_device(cube, -angs))
        lib.vb_profile_enable(1)
        derotate_device(cube, -angs)
        prof = (C.c_float * 4)()
        lib.vb_profile_read(prof)
        lib.vb_profile_enable(0)
        print(f"derotate n={n} size={size} scratch={os.environ.get('VIP_B200_DEROT_SCRATCH')} "
              f"nt={os.environ.get('VIP_B200_FFT_NT')}: {ms:.3f} ms  passes {prof[0]:.2f}/{prof[1]:.2f}/{prof[2]:.2f} "
              f"chunks {int(prof[3])}")
    elif what == "eigh":
        k = size if len(sys.argv) > 3 else 20
        from tools.synth import adi_cube
        cube, _ = adi_cube(n, 64, k, 60.0, seed=3)
        M = torch.from_numpy(cube.reshape(n, -1)).cuda()
        G = kernels.gram(M)
        t0 = timeit(lambda: kernels.eigh_topk(G, k), reps=3, warm=1)
        _, _, info = kernels.eigh_topk(G, k)
        t1 = timeit(lambda: kernels.eigh(G), reps=2, warm=1)
        _, _, info2 = kernels.eigh(G)
        print(f"eigh n={n} k={k}: topk {t0:.3f} ms {info}; jacobi {t1:.3f} ms {info2}")
    elif what == "gram":
        M = torch.randn((n, size * size), device="cuda", generator=g)
        print(f"gram n={n} p={size*size}: {timeit(lambda: kernels.gram(M)):.3f} ms")
    elif what == "median":
        cube = torch.randn((n, size, size), device="cuda", generator=g)
        print(f"median n={n} size={size}: {timeit(lambda: collapse_device(cube, 'median')):.3f} ms")
    elif what == "proj":
        k = 20
        M = torch.randn((n, size * size), device="cuda", generator=g)
        Wt = torch.randn((k, n), device="cuda", generator=g)
        Cm = torch.randn((n, k), device="cuda", generator=g)
        V = kernels.pcs(Wt, M)
        print(f"pcs {timeit(lambda: kernels.pcs(Wt, M)):.3f} ms; subtract {timeit(lambda: kernels.project_subtract(M, Cm, V)):.3f} ms")


if __name__ == "__main__":
    main()
