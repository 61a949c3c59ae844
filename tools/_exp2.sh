python tools/bench_stage.py median 500 512 2>&1 | tail -1
python tools/bench_stage.py median 1000 512 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x -k "collapse or single or grid" 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'shear_cols_fft|collapse_median_smem' -c 2 -o gpurun_out/prof_r01e -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full_r01e.log 2>&1
tail -3 gpurun_out/ncu_full_r01e.log
ls -la gpurun_out/*.ncu-rep
