#!/bin/bash
# r01n: A/B of the fp64 tile kernel v2 / single-launch Jacobi / two-pixel pcs+subtract against the previous kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_r01n.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r01n_new.json 2> gpurun_out/bench_r01n_new.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01n_new.json").read().strip().splitlines()[-1])
print("NEW", d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items()})
PY
VIP_B200_PCS_PX=1 VIP_B200_SUB_PX=1 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r01n_old.json 2> gpurun_out/bench_r01n_old.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01n_old.json").read().strip().splitlines()[-1])
print("OLD", d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items()})
PY
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1"
$TR --master-port 29531 tools/scale_c5.py 300 2>&1 | grep -E "C5 slice|rror" | tail -3
VIP_B200_GRAM_V1=1 VIP_B200_JACOBI_SMALL=0 VIP_B200_PCS_PX=1 VIP_B200_SUB_PX=1 $TR --master-port 29532 tools/scale_c5.py 300 2>&1 | grep -E "C5 slice|rror" | tail -3
VIP_B200_GRAM_V1=1 $TR --master-port 29533 tools/scale_c5.py 300 2>&1 | grep -E "C5 slice|rror" | tail -3
VIP_B200_JACOBI_SMALL=0 $TR --master-port 29534 tools/scale_c5.py 300 2>&1 | grep -E "C5 slice|rror" | tail -3
python tools/run_configs.py c1 c5 2>&1 | grep -v Warning | tail -6
