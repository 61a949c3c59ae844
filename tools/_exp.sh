for nt in 4 2; do for sc in 8589934592 268435456 100663296; do VIP_B200_FFT_NT=$nt VIP_B200_DEROT_SCRATCH=$sc python tools/bench_stage.py derotate 500 512 2>&1 | tail -1; done; done
python tools/bench_stage.py eigh 500 20 2>&1 | tail -1
python tools/bench_stage.py eigh 200 10 2>&1 | tail -1
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
