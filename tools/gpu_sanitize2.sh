#!/bin/bash
# compute-sanitizer over the kernels added in the second half of round 2: second-generation fused eigensolver, warp-per-pixel
# median, ncomp='auto' direct annular solver, peak mask, FITS decode, annular ADI+mSDI path.
mkdir -p gpurun_out
SEL="eigh_topk_fused2_sizes and (128-8 or 257-20 or 500-11) or collapse_median_warp_kernel and (96 or 257 or 513) or annular_ncomp_auto or local_max_mask or fits_decode or annular_adimsdi"
timeout 1000 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer2_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer2_memcheck.log
RSEL="eigh_topk_fused2_sizes and (128-8 or 257-20) or collapse_median_warp_kernel and (96 or 513) or local_max_mask"
timeout 1000 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$RSEL" > gpurun_out/sanitizer2_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer2_racecheck.log
grep -c "Race reported\|hazard" gpurun_out/sanitizer2_racecheck.log
