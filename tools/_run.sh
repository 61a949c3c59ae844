python tools/bench_stage.py derotate 100 1024 2>&1 | tail -1
VIP_B200_FFT_NT=2 python tools/bench_stage.py derotate 100 1024 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x -k "derot" 2>&1 | tail -2
python tools/run_configs.py c5 2>&1 | grep -v Warn | tail -3
