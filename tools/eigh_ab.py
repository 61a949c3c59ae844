"""A/B timing of the top-k eigensolver generations on the config-2 Gramian (500 x 500, k = 20) and at n = 1000:
VIP_B200_TOPK_FUSED=1 (three barriers per iteration, serial Cholesky / solve) against =2 (default)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube          # noqa: E402
from vip_b200 import kernels              # noqa: E402

out = {}
for n, size, k in ((500, 96, 20), (1000, 64, 20), (500, 96, 8)):
    cube, _ = adi_cube(n, size, k, 90.0, seed=20260102)
    M = cube.reshape(n, -1).astype(np.float64)
    G = torch.from_numpy(M @ M.T).cuda()
    w = np.linalg.eigvalsh(G.cpu().numpy())[::-1]
    for mode in ("1", "2"):
        os.environ["VIP_B200_TOPK_FUSED"] = mode
        ev, E, info = kernels.eigh_topk(G, k)
        err = float(np.max(np.abs(ev.cpu().numpy() - w[:k]) / w[:k]))
        for _ in range(3):
            kernels.eigh_topk_async(G, k)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            kernels.eigh_topk_async(G, k)
        t1.record()
        torch.cuda.synchronize()
        out[f"n{n}_k{k}_fused{mode}"] = {"ms": t0.elapsed_time(t1) / 20, "iters": info["iters"],
                                         "converged": info["converged"], "eval_rel_err": err}
        print(f"n={n} k={k} fused={mode}: {out[f'n{n}_k{k}_fused{mode}']}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/eigh_ab.json", "w"), indent=1)
