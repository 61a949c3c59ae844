#!/bin/bash
# One GPU-box session: parity tests, bench line, ncu launch list + full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
# launch list of one bench step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3200 -c 1300 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launches_${TAG}.log
# full capture of the three shear kernels + gram + median (one launch each)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'shear_cols_fft|shear_rows_first_fft|shear_rows_last_fft|gram_tile_kernel|collapse_median_kernel|subtract_kernel|pcs_kernel' \
    -s 21 -c 7 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out
