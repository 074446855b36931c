#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== sharded single-rank batch test"; timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -k "sharded_search_batch" 2>&1 | tail -3
echo "== sharded batch N=$N depth 2"; timeout 600 $TR scripts/bench_batch_sharded.py 20 2 2>&1 | grep '^{\|Error\|error' | tail -3 | tee $OUT/bench_batch_sharded_r39_n${N}_d2.json
echo "== sharded batch N=$N depth 1"; timeout 600 $TR scripts/bench_batch_sharded.py 20 1 2>&1 | grep '^{\|Error\|error' | tail -3 | tee $OUT/bench_batch_sharded_r39_n${N}_d1.json
echo "== sharded batch N=1 depth 2"; timeout 600 python scripts/bench_batch_sharded.py 20 2 2>&1 | grep '^{\|Error\|error' | tail -3 | tee $OUT/bench_batch_sharded_r39_n1_d2.json
