#!/bin/bash
# round 2, call O (8 GPUs): config 4 at full size on 8 / 4 / 2 / 1 GPUs, config 5 on 8 GPUs with NUMA-local pinned shards
mkdir -p gpurun_out
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N"
  timeout 600 $TR tools/scale_c4.py 300 2>&1 | grep "C4FULL\|Error\|error" | tail -3 | tee -a gpurun_out/c4_full_r02o.log
done
timeout 600 python tools/scale_c4.py 300 2>&1 | grep "C4FULL\|Error\|error" | tail -3 | tee -a gpurun_out/c4_full_r02o.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591"
timeout 600 $TR tools/scale_c5_full.py 4000 2>&1 | grep "C5FULL\|Error\|error" | tail -3 | tee -a gpurun_out/c5_full_r02o.log
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_r02o.txt
