for rr in 1 2 3 4 6; do echo "fused rr=$rr"; VIP_B200_TOPK_RR=$rr timeout 120 python tools/bench_stage.py eigh 500 20 2>&1 | tail -1; done
timeout 600 python -m pytest tests -m gpu -q -x -k "eigh or topk or decomposition or pca_c1 or pca_medium or c2_full or annular" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r01h.json 2>gpurun_out/bench_r01h.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_r01h.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches_per_step']); print(d['stage_ms'])
PY
