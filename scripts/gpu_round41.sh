#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_r41.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['clocks']); print(d['shadow_filtered']); print(d['batched']); print(d['cpu_baseline'])"
