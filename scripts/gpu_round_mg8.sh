#!/bin/bash
# 8-GPU round: parity, strong scaling at N=8 and N=4, and BASELINE configs[3] (100M x 384 over 8 GPUs)
set -u
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi -L | head -8
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@" 2>&1 | grep -v -i "warn\|OMP_NUM\|\*\*\*\*"; }
echo "== sharded parity N=8"; timeout 600 bash -c "$(declare -f run); run 8 tests/check_sharded_torchrun.py" | tail -6 | tee $OUT/sharded_parity_n8.txt
echo "== bench N=8 (10M strong)"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 1000 --warmup 50" | grep '^{' | tail -1 | tee $OUT/bench_n8.json | cut -c1-400
echo "== bench N=8 depth 4"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 1000 --warmup 50 --pipeline 4" | grep '^{' | tail -1 | tee $OUT/bench_n8_depth4.json | cut -c1-300
echo "== bench N=8 depth 1"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 1000 --warmup 50 --pipeline 1" | grep '^{' | tail -1 | tee $OUT/bench_n8_depth1.json | cut -c1-300
echo "== bench N=4"; timeout 600 bash -c "$(declare -f run); run 4 bench.py --gpus 4 --steps 800 --warmup 40" | grep '^{' | tail -1 | tee $OUT/bench_n4.json | cut -c1-300
echo "== C4: 100M x 384 over 8 GPUs"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 200 --warmup 10 --rows 100000000" | grep '^{' | tail -1 | tee $OUT/bench_c4_100m_n8.json | cut -c1-400
