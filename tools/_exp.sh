python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for nt in 2 3 4; do VIP_B200_FFT_NT=$nt python tools/bench_stage.py derotate 500 512 2>&1 | tail -1; done
python tools/bench_stage.py derotate 100 1024 2>&1 | tail -1
python tools/bench_stage.py derotate 300 256 2>&1 | tail -1
python tools/bench_stage.py proj 500 512 2>&1 | tail -1
