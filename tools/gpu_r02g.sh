#!/bin/bash
# round 2, call G (2 GPUs): multi-rank NCCL parity from pytest, bench --gpus 2, config 5 at full size on 2 GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_r02g.log; grep -n "passed\|failed\|\[parity\]\|FAILED\|Error\|nccl_worker" gpurun_out/pytest_r02g.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02g_n2.json 2> gpurun_out/bench_r02g_n2.err; tail -c 1800 gpurun_out/bench_r02g_n2.json; tail -3 gpurun_out/bench_r02g_n2.err
VIP_B200_SHARD_OVERLAP=0 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02g_n2_nooverlap.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_r02g_n2_nooverlap.json')); print('no overlap: step %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']), d['stage_ms'])"
timeout 600 $TR tools/scale_c5_full.py 4000 2>&1 | grep "C5FULL\|Error\|error" | tail -3
