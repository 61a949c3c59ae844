#!/bin/bash
# round 2, call A: compute-sanitizer pass (bounded) + BASELINE config 5 at FULL size on one GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
SEL="c1_golden or collapse_bit_exact or cross_gram_shapes or eigh_small_and_odd or derotate_golden or annular_golden or pcs_and_project"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck.log
RSEL="c1_golden or derotate_golden or collapse_bit_exact"
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$RSEL" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer_racecheck.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29561"
timeout 420 $TR tools/scale_c5_full.py 4000 > gpurun_out/c5_full_n1.log 2>&1
echo "c5 full rc=$?"; tail -3 gpurun_out/c5_full_n1.log
