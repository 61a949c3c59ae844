#!/bin/bash
# round 2, call F: parity of the reworked randsvd / hp projection / padded FFTs on the whole GPU suite, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_r02f.log; grep -n "passed\|failed\|\[parity\]\|FAILED\|Error" gpurun_out/pytest_r02f.log | tail -40
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02f.json"))
s=d["stage_ms"]
print("step %.3f ms e2e %.3f pageable %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"]))
print(json.dumps(s))
print(json.dumps(d.get("parity_vs_reference_golden")))
PY
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 $TR tools/scale_c5_full.py 4000 2>&1 | grep C5FULL
