#!/bin/bash
# round 2, call E: stable randomized SVD parity, C2 goldens with the fp64 truth, FFT mode A/B (0 scalar / 1 packed / 2 packed+padded)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA -k "randsvd or c5 or c2_ or sharded_world1 or derotate_golden or derotate_vs_oracle or nyquist" 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_r02e.log; grep -n "passed\|failed\|\[parity\]\|FAILED\|Error" gpurun_out/pytest_r02e.log | tail -30
for v in 0 2; do
  VIP_B200_FFT_F32X2=$v python -m pytest tests -m gpu -q -k "derotate" 2>&1 | tail -1
  VIP_B200_FFT_F32X2=$v python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r02e_f32x2_$v.json 2> gpurun_out/bench_r02e_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02e_f32x2_$v.json"))
s=d["stage_ms"]
print("F32X2=$v step %.3f ms e2e %.3f pageable %.3f | derotate %.3f (rows1 %.3f cols %.3f rows3 %.3f) fft frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"], s["derotate_ms"], s["shear_rows_first_ms"], s["shear_cols_ms"], s["shear_rows_last_ms"], d["roofline_fft"]["frac"]))
print(json.dumps(d.get("parity_vs_reference_golden")))
PY
done
python tools/scale_c5.py 300 2>&1 | tail -2
