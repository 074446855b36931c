#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_gpu17.txt
echo "== dims sweep"; timeout 900 python scripts/dims_sweep.py $OUT/dims_sweep.json 2>&1 | tail -16
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 2>&1 | tail -1 | tee $OUT/bench17.json | cut -c1-200
