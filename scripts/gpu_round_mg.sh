#!/bin/bash
# multi-GPU round: N = $1 (2, 4 or 8)
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== sharded parity"; timeout 600 $TR tests/check_sharded_torchrun.py 2>&1 | grep -v -i "warn" | tail -8 | tee $OUT/sharded_parity_n$N.txt
echo "== bench N=$N"; timeout 600 $TR bench.py --gpus $N --steps 400 --warmup 20 2>&1 | grep '^{' | tail -1 | tee $OUT/bench_n$N.json
echo "== bench N=$N pipeline depth 1"; timeout 600 $TR bench.py --gpus $N --steps 400 --warmup 20 --pipeline 1 2>&1 | grep '^{' | tail -1 | tee $OUT/bench_n${N}_depth1.json
echo "== bench N=$N pipeline depth 4"; timeout 600 $TR bench.py --gpus $N --steps 400 --warmup 20 --pipeline 4 2>&1 | grep '^{' | tail -1 | tee $OUT/bench_n${N}_depth4.json
echo "== reference arm under torchrun"; timeout 300 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | grep '^{' | tail -1 | cut -c1-300
