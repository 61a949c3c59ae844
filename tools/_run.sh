VIP_B200_FFT_SLAB=8 python tools/bench_stage.py derotate 500 512 2>&1 | tail -1
VIP_B200_FFT_SLAB=4 python tools/bench_stage.py derotate 500 512 2>&1 | tail -1
VIP_B200_FFT_SLAB=8 timeout 900 python -m pytest tests -m gpu -q -x -k "derot" 2>&1 | tail -2
