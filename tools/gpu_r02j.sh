#!/bin/bash
# round 2, call J (2 GPUs): exchanges fused over peer memory -- multi-rank parity, bench --gpus 2, config 5 full size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_r02j.log; grep -n "passed\|failed\|\[parity\]\|FAILED\|Error\|error" gpurun_out/pytest_r02j.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
for F in 1 0; do
VIP_B200_SHARD_FUSED=$F timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench_r02j_n2_f$F.err | grep '^{' > gpurun_out/bench_r02j_n2_f$F.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r02j_n2_f$F.json")); print("fused=$F step %.3f e2e %.3f parity %.2e" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_vs_single"]["rel_err"]), d.get("exchange"), d["stage_ms"])
except Exception as e: print("failed", e)
PY
tail -3 gpurun_out/bench_r02j_n2_f$F.err | grep -v "^\*\|OMP"
done
timeout 600 $TR tools/scale_c5_full.py 4000 2>&1 | grep "C5FULL\|Error\|error" | tail -3 | tee -a gpurun_out/c5_full_r02j.log
