python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for nt in 2 3 5 1; do VIP_B200_FFT_NT=$nt python tools/bench_stage.py derotate 500 512 2>&1 | tail -1; done
python tools/bench_stage.py gram 500 512 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r01d.json 2>gpurun_out/bench_r01d.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_r01d.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches_per_step'])
PY
tail -3 gpurun_out/bench_r01d.err
