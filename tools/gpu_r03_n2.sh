#!/bin/bash
# round 2 final state on 2 GPUs: multi-rank parity (world 2) with the new eigensolver / median kernels, bench --gpus 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_r03_n2.log; grep -n "passed\|failed\|\[parity\]\|FAILED\|Error\|error" gpurun_out/pytest_r03_n2.log | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench_r03_n2.err | grep '^{' > gpurun_out/bench_r03_n2.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r03_n2.json")); print("N=2 step %.3f e2e %.3f parity %.2e" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_vs_single"]["rel_err"]), d.get("exchange"), d["stage_ms"])
except Exception as e: print("failed", e)
PY
tail -3 gpurun_out/bench_r03_n2.err | grep -v "^\*\|OMP"
