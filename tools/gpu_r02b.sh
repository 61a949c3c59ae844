#!/bin/bash
# round 2, call B: packed-fp32 microbenchmarks, the full GPU suite with the config-size parity tests, bench line
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32_rate tools/microbench/fp32_rate.cu && /tmp/fp32_rate > gpurun_out/fp32_rate.txt 2>&1
./tools/microbench/fp32x2_rate > gpurun_out/fp32x2_rate.txt 2>&1; cat gpurun_out/fp32x2_rate.txt
timeout 1500 python -m pytest tests -m gpu -q -rA -x 2>&1 | grep -v "Warning\|warnings.warn" | tail -60 > gpurun_out/pytest_r02b.log; tail -45 gpurun_out/pytest_r02b.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -c 1500 gpurun_out/bench_r02b.json
