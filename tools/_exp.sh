for ch in 64 128 256; do echo "== chunk $ch"; VIP_B200_GRAM_CHUNK=$ch timeout 300 python tools/check_gram_tc.py c2 2>&1 | grep "BK=64"; done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python tools/bench_stage.py median 500 512 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r01d.json 2>gpurun_out/bench_r01d.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_r01d.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches_per_step']); print(d['stage_ms'])
PY
tail -3 gpurun_out/bench_r01d.err
