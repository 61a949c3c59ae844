#!/bin/bash
# round 2, call M (8 GPUs): config 5 at full size on 8 / 4 / 2 / 1 GPUs of ONE box, bench --gpus 8 (fused and NCCL exchanges), NCCL parity at world 8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 8 4 2 1; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N"
  timeout 600 $TR tools/scale_c5_full.py 4000 2>&1 | grep "C5FULL\|Error\|error" | tail -3 | tee -a gpurun_out/c5_full_r02m.log
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591"
for F in 1 0; do
VIP_B200_SHARD_FUSED=$F timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/bench_r02m_n8_f$F.err | grep '^{' > gpurun_out/bench_r02m_n8_f$F.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r02m_n8_f$F.json")); print("fused=$F step %.3f e2e %.3f parity %.2e" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_vs_single"]["rel_err"]), d.get("exchange"), d["stage_ms"])
except Exception as e: print("failed", e)
PY
done
timeout 600 $TR tests/nccl_worker.py gpurun_out/parity_nccl_world8.json 2>&1 | grep "nccl_worker\|Error" | tail -3
