#!/bin/bash
# round 2, call C: full GPU suite (config-size parity), smoke, bench line with the new keys
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_r02c.log; grep -n "passed\|failed\|\[parity\]\|FAILED\|Error" gpurun_out/pytest_r02c.log | tail -40
python __graft_entry__.py smoke 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -c 2500 gpurun_out/bench_r02c.json; tail -3 gpurun_out/bench_r02c.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02c.json 2>&1; tail -c 800 gpurun_out/bench_ref_r02c.json
