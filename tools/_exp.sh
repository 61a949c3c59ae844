python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for sc in 8589934592 100663296 67108864; do VIP_B200_DEROT_SCRATCH=$sc python tools/bench_stage.py derotate 500 512 2>&1 | tail -1; done
python - <<'PY'
# C4-like timing: 39 x 32 x 256 x 256 double PCA (32 ADI frames of the 300)
import time, numpy as np, torch, sys
sys.path.insert(0,'.')
import vip_b200
rng=np.random.default_rng(0)
z,n,S=39,32,256
lam=np.linspace(0.95,1.65,z); sl=lam.max()/lam
cube=(rng.normal(size=(z,n,S,S))*3+100).astype(np.float32)
angs=np.linspace(0,60,n)
for _ in range(2):
    torch.cuda.synchronize(); t=time.time()
    fr=vip_b200.pca(cube,angs,scale_list=sl,adimsdi='double',ncomp=(3,10),verbose=False)
    torch.cuda.synchronize(); print('C4 slice 39x32x256x256: %.3f s -> %.1f ADI frames/s'%(time.time()-t, n/(time.time()-t)))
PY
