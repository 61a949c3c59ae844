#!/bin/bash
# A/B of the pass-2 variants + ncu of the packed kernels
mkdir -p gpurun_out
for cv in 0 1 2; do VIP_B200_FFT_COLS=$cv python tools/bench_stage.py derotate 500 512 2>&1 | tail -1; done
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01k.json 2> gpurun_out/bench_r01k.err
tail -c 1500 gpurun_out/bench_r01k.json | head -c 1500; tail -3 gpurun_out/bench_r01k.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'shear_rows_first_pk|shear_cols_pk|shear_rows_last_pk' \
    -c 3 -f -o gpurun_out/prof_r01k_pk python tools/bench_stage.py derotate 100 512 > gpurun_out/ncu_pk.log 2>&1
tail -2 gpurun_out/ncu_pk.log
