#!/bin/bash
# round 2, call P (2 GPUs): multi-rank parity after the staged-upload change, config 4 at N = 2 and N = 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA 2>&1 | grep -n "passed\|failed\|\[parity\]\|FAILED\|Error" | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582"
timeout 600 $TR tools/scale_c4.py 300 2>&1 | grep "C4FULL\|Error\|error" | tail -3 | tee -a gpurun_out/c4_full_r02p.log
timeout 600 python tools/scale_c4.py 300 2>&1 | grep "C4FULL\|Error\|error" | tail -3 | tee -a gpurun_out/c4_full_r02p.log
