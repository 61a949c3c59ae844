#!/bin/bash
# round 2, call H (4 GPUs): config 5 at full size on 1 / 2 / 4 GPUs (same box, same cube), bench --gpus 4 with and without the overlapped exchange
mkdir -p gpurun_out
for N in 1 2 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N"
  timeout 600 $TR tools/scale_c5_full.py 4000 2>&1 | grep "C5FULL\|Error\|error" | tail -3 | tee -a gpurun_out/c5_full_r02h.log
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591"
timeout 600 $TR bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/bench_r02h_n4.err | grep '^{' > gpurun_out/bench_r02h_n4.json
VIP_B200_SHARD_OVERLAP=0 timeout 600 $TR bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/bench_r02h_n4_nooverlap.err | grep '^{' > gpurun_out/bench_r02h_n4_nooverlap.json
tail -5 gpurun_out/bench_r02h_n4_nooverlap.err
python - <<PY
import json
for f in ("gpurun_out/bench_r02h_n4.json", "gpurun_out/bench_r02h_n4_nooverlap.json"):
    try:
        d=json.load(open(f)); print(f, "step %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), d["parity_vs_single"]["rel_err"], d["stage_ms"])
    except Exception as e: print(f, "failed", e)
PY
