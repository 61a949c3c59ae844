#!/bin/bash
# round 2, third session final state (r04): bench line, launch list of one bench step, ncu --set full of the hot kernels (config 2)
TAG=${1:-r04}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launches_${TAG}.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'gram_umma|split_planes|slab_mean|subtract_hp|pcs_kernel|collapse_median_warp|shear_rows_first_pk|shear_rows_last_pk|shear_cols_pk|topk_fused2' \
    -c 11 -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
