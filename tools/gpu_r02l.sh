#!/bin/bash
# round 2, call L: full GPU suite after the eigensolver / whitening / hp-kernel changes, bench line, config 5 on one GPU
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rA 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_r02l.log; grep -n "passed\|failed\|\[parity\]\|FAILED\|Error" gpurun_out/pytest_r02l.log | tail -40
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02l.json"))
print("step %.3f ms e2e %.3f pageable %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"]))
print(json.dumps(d["stage_ms"]))
print(json.dumps(d.get("parity_vs_reference_golden")))
PY
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 $TR tools/scale_c5_full.py 4000 2>&1 | grep "C5FULL\|Error" | tee gpurun_out/c5_full_r02l.log
