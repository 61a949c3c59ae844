"""In-process A/B of the eigensolver and median variants on the BASELINE config-2 cube (1 GPU).
The switches are read at every call, so one process generates the cube once and times every setting."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube                                        # noqa: E402
from vip_b200 import kernels                                             # noqa: E402
from vip_b200.preproc.derotation import derotate_device                  # noqa: E402
from vip_b200.preproc.subsampling import collapse_device                 # noqa: E402


def timeit(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


n, size, k = 500, 512, 20
cube, angs = adi_cube(n, size, k, 90.0, seed=20260102)
dev = torch.from_numpy(cube).cuda()
M = dev.reshape(n, -1)
G = kernels.gram(M)
w_ref = np.linalg.eigvalsh(G.cpu().numpy())[::-1][:k]
for env in ({}, {"VIP_B200_TOPK_CHOL": "1"}, {"VIP_B200_TOPK_RR": "0"},
            {"VIP_B200_TOPK_CHOL": "1", "VIP_B200_TOPK_RR": "0"},
            {"VIP_B200_TOPK_CHOL": "1", "VIP_B200_TOPK_RR": "0", "VIP_B200_TOPK_RR0": "6"},
            {"VIP_B200_TOPK_CHOL": "1", "VIP_B200_TOPK_RR": "0", "VIP_B200_TOPK_RR0": "10"}):
    def run():
        ms = timeit(lambda: kernels.eigh_topk(G, k))
        ev, _, info = kernels.eigh_topk(G, k)
        err = float(np.max(np.abs(ev.cpu().numpy() - w_ref) / w_ref))
        return ms, info, err
    ms, info, err = with_env(env, run)
    print(f"eigh_topk n={n} k={k} {env or 'default'}: {ms:.3f} ms {info} eval rel err {err:.1e}", flush=True)

# median on the real derotated residual cube
evals, evecs, _ = kernels.eigh_topk(G, k)
S = torch.sqrt(evals)
V = kernels.pcs((evecs / S[:, None]).contiguous(), M)
R = kernels.project_subtract(M, (evecs * S[:, None]).t().float().contiguous(), V).reshape(n, size, size)
D = derotate_device(R, -angs)
ref = None
for env in ({"VIP_B200_MEDIAN_ALGO": "radix"}, {}, {"VIP_B200_MEDIAN_CFG": "8,24"}, {"VIP_B200_MEDIAN_CFG": "8,48"},
            {"VIP_B200_MEDIAN_CFG": "4,40"}, {"VIP_B200_MEDIAN_CFG": "4,80"}, {"VIP_B200_MEDIAN_CFG": "16,24"},
            {"VIP_B200_MEDIAN_CFG": "16,40"}):
    def run():
        ms = timeit(lambda: collapse_device(D, "median"))
        return ms, collapse_device(D, "median")
    ms, out = with_env(env, run)
    if ref is None:
        ref = out
    same = bool(torch.equal(out, ref))
    print(f"median {n}x{size}x{size} {env or 'default (range)'}: {ms:.3f} ms  identical to radix: {same}", flush=True)
