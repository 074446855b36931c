#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
echo "== two-device test"; timeout 300 python -m pytest tests/test_gpu_engine.py -q -k "two_devices" 2>&1 | tail -3
echo "== bench N=1 (NVML clocks)"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_r40_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks'], d['shadow_filtered']['value'])"
echo "== bench N=2"; timeout 600 $TR bench.py --gpus 2 --steps 400 --warmup 20 2>&1 | grep '^{' | tail -1 | tee $OUT/bench_r40_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks'])"
