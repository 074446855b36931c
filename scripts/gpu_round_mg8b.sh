#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@" 2>&1 | grep '^{' | tail -1; }
for cfg in "8 4 2" "8 8 2" "8 1 4" "4 4 2"; do set -- $cfg
  echo "== bench N=$1 micro=$2 depth=$3"; timeout 600 bash -c "$(declare -f run); run $1 bench.py --gpus $1 --steps 1200 --warmup 60 --micro $2 --pipeline $3" | tee $OUT/bench_n$1_m$2_d$3.json | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print({k:b[k] for k in ('value','ms_per_step','queries_in_flight')}, 'e2e', round(b['e2e']['value'],1), b['clocks']['reasons'])"
done
echo "== C4 100M"; timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 300 --warmup 20 --rows 100000000" | tee $OUT/bench_c4_100m_n8b.json | cut -c1-200
