#!/bin/bash
# Last validation of the round on a fresh box: what the driver runs (GPU suite with -x, smoke, bench.py).
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== pytest -x -q -m gpu"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_gpu_final3.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_final3.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['shadow_filtered']['value'], d['batched']['value'], d['clocks'])"
