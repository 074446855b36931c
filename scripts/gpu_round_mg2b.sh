#!/bin/bash
# 2-GPU round after the bf16 work: sharded parity incl. the batched form, the N=2 headline bench, sharded batch bench.
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== sharded single-rank batch test"; timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -k "sharded_search_batch" 2>&1 | tail -3
echo "== sharded parity"; timeout 600 $TR tests/check_sharded_torchrun.py 2>&1 | grep -v -i "warn" | tail -9 | tee $OUT/sharded_parity_r34_n$N.txt
echo "== bench N=$N"; timeout 600 $TR bench.py --gpus $N --steps 400 --warmup 20 2>&1 | grep '^{' | tail -1 | tee $OUT/bench_r34_n$N.json | cut -c1-400
echo "== sharded batch N=$N"; timeout 600 $TR scripts/bench_batch_sharded.py 10 2>&1 | grep '^{' | tail -1 | tee $OUT/bench_batch_sharded_r34_n$N.json
echo "== sharded batch N=1"; timeout 600 python scripts/bench_batch_sharded.py 10 2>&1 | grep '^{' | tail -1 | tee $OUT/bench_batch_sharded_r34_n1.json
