#!/bin/bash
# r01p: final check of the committed defaults: parity tests, bench line, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_r01p.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01p.json 2> gpurun_out/bench_r01p.err
tail -c 700 gpurun_out/bench_r01p.json; tail -2 gpurun_out/bench_r01p.err
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r01p.csv python bench.py --steps 1 --warmup 1 --no-cpu \
    > gpurun_out/ncu_launches_r01p.log 2>&1
tail -c 300 gpurun_out/ncu_launches_r01p.log
