#!/bin/bash
# compute-sanitizer over the small-configuration parity tests: memcheck on every kernel family,
# racecheck on the shared-memory heavy ones (FFT shears, median, fused eigensolver, tcgen05 Gramian).
# Slow (10-50x): run on one GPU with a generous timeout, e.g.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
mkdir -p gpurun_out
SEL="c1_golden or collapse_bit_exact or collapse_median_edge or cross_gram_shapes or eigh_small or eigh_topk_matches or derotate_golden or annular_golden or pcs_and_project"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck.log
RSEL="c1_golden or collapse_median_edge or eigh_topk_matches or derotate_golden"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$RSEL" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer_racecheck.log
