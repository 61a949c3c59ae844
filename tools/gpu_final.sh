#!/bin/bash
# End-of-round evidence run: parity tests, bench line, launch list, ncu --set full of the hot kernels, other configs.
TAG=${1:-r01g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_${TAG}.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launches_${TAG}.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'gram_umma|split_planes|slab_mean|subtract_kernel|pcs_kernel|collapse_median_smem|shear_rows_first_pk|shear_rows_last_pk|shear_cols_pk|topk_fused' \
    -c 11 -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
python tools/run_configs.py c1 c3 c4 c5 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/configs_${TAG}.log
# config-5-shaped slice: launch list and ncu --set full of the fp64 CUDA-core kernels of the randomized SVD
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_c5_${TAG}.csv $TR tools/scale_c5.py 300 > gpurun_out/ncu_c5_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_tile2|jacobi_small|pcs_kernel' \
    -c 4 -f -o gpurun_out/prof_c5_${TAG} $TR tools/scale_c5.py 300 > gpurun_out/ncu_full_c5_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_c5_${TAG}.log
