import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vip_b200 import kernels
n, S = int(sys.argv[1]) if len(sys.argv) > 1 else 500, 512
g = torch.Generator(device="cuda").manual_seed(1)
cube = torch.randn((n, S * S), device="cuda", generator=g) * 5.0
nan_mode = sys.argv[2] if len(sys.argv) > 2 else "corner"
if nan_mode == "corner":
    yy, xx = np.mgrid[:S, :S]
    corner = torch.from_numpy((np.hypot(yy - S / 2, xx - S / 2) > S / 2 * 1.3).reshape(-1)).cuda()
    cube[:n // 2, corner] = float("nan")
for rep in range(3):
    ts = []
    for _ in range(6):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); out = kernels.collapse(cube, "median"); t1.record(); torch.cuda.synchronize()
        ts.append(round(t0.elapsed_time(t1), 4))
    print(n, nan_mode, ts, flush=True)
