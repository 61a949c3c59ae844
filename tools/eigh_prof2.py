import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import adi_cube
from vip_b200 import kernels
cube, _ = adi_cube(500, 128, 20, 90.0, seed=20260102)
M = torch.from_numpy(cube.reshape(500, -1)).cuda()
G = kernels.gram(M)
for prof in (1, 3, 5, 9, 15):
    os.environ["VIP_B200_TOPK_PROF"] = str(prof)
    print("prof bits", prof, flush=True)
    try:
        kernels.eigh_topk(G, 20, max_iter=8)
    except Exception as e:
        print("err", e)
    torch.cuda.synchronize()
