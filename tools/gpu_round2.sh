#!/bin/bash
# GPU-box session: parity tests, bench line, launch list, other configs.
TAG=${1:-r01c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_${TAG}.log
tail -5 gpurun_out/pytest_${TAG}.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launches_${TAG}.log
python tools/run_configs.py c1 c3 c4 2>&1 | tail -8 | tee gpurun_out/configs_${TAG}.log
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
nproc
