BENCH_ANGLE_MAX=40 python tools/bench_stage.py derotate 500 512 2>&1 | tail -1
python tools/bench_stage.py derotate 500 512 2>&1 | tail -1
