#!/bin/bash
# round 2, call D: packed-fp32 (FFMA2) shear transforms -- parity tests and A/B timing; reworked randsvd parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -k "derotate or randsvd or c5 or median_sub or cube_shift or c1_golden" 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_r02d.log; grep -n "passed\|failed\|\[parity\]\|FAILED\|Error" gpurun_out/pytest_r02d.log | tail -30
for v in 0 1; do
  VIP_B200_FFT_F32X2=$v python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r02d_f32x2_$v.json 2> gpurun_out/bench_r02d_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02d_f32x2_$v.json"))
s=d["stage_ms"]
print("F32X2=$v step %.3f ms e2e %.3f | derotate %.3f (rows1 %.3f cols %.3f rows3 %.3f) fft frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], s["derotate_ms"], s["shear_rows_first_ms"], s["shear_cols_ms"], s["shear_rows_last_ms"], d["roofline_fft"]["frac"]))
PY
done
