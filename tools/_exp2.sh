python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP" | tail -2 | cut -c1-900
python -m pytest tests -m gpu -q -x -k "sharded or upload" 2>&1 | tail -3
