#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== gpu tests (rank-local)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu12.txt
echo "== sharded parity N=2 (incl. micro-batched exchange)"; timeout 600 $TR tests/check_sharded_torchrun.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" | tail -8
for m in 1 4 8; do
echo "== bench N=2 micro=$m"; timeout 600 $TR bench.py --gpus $N --steps 400 --warmup 20 --micro $m 2>&1 | grep '^{' | tail -1 | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print({k:b[k] for k in ('value','ms_per_step','queries_in_flight')}, 'e2e', b['e2e']['value'])"
done
timeout 300 python scripts/latency.py $OUT/latency12.json 2>&1 | head -3
