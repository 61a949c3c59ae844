#!/bin/bash
# round 2 evidence: launch list of one bench step + ncu --set full of the hot kernels (config 2), config-5 slice launch list
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launches_${TAG}.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'gram_umma|split_planes|slab_mean|subtract_hp|pcs_kernel|collapse_median_warp|shear_rows_first_pk|shear_rows_last_pk|shear_cols_pk|topk_fused' \
    -c 11 -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches_c5_${TAG}.csv $TR tools/scale_c5.py 300 > gpurun_out/ncu_c5_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_c5_${TAG}.log | cut -c1-200
