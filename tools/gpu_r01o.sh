#!/bin/bash
# r01o: tests, eigensolver / median A/B on the config-2 cube, config-5 strong scaling with the r01n kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_r01o.log
python tools/ab_r01o.py 2>&1 | grep -E "eigh_topk|median|rror" | tee gpurun_out/ab_r01o.log
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4; do
  if [ $n -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n tools/scale_c5.py 600 2>&1 | grep -E "C5 slice|rror" | tail -2 | tee -a gpurun_out/c5scale_r01o.log
  fi
done
