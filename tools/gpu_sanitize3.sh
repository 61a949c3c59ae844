#!/bin/bash
# compute-sanitizer memcheck over the kernels of the third session of round 2: tcgen05 batched GEMM + bf16x3 split,
# packed shared-memory direct eigensolver (tiny and 200-frame libraries), batched ADI+mSDI first pass with the
# prefetching upload.
mkdir -p gpurun_out
SEL="gemm_tc or tiny_libraries or sdi_double_golden or annular_direct_solver_vs_numpy"
[ "$1" = "0" ] || timeout ${1:-170} compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer3_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer3_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitizer3_memcheck.log
if [ -n "$2" ]; then
RSEL="tiny_libraries or annular_direct_solver_vs_numpy or gemm_tc"
timeout $2 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$RSEL" > gpurun_out/sanitizer3_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer3_racecheck.log
echo "hazards: $(grep -c 'Race reported\|hazard' gpurun_out/sanitizer3_racecheck.log)"
fi
true
