#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== batch tests first (new tensor path; short timeout guards against a hang)"
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -25 | tee $OUT/pytest_batch.txt
echo "== all gpu tests"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu3.txt
echo "== sweep3"; timeout 900 python scripts/sweep3.py $OUT/sweep3.json 2>&1 | tail -60
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -2 | tee $OUT/bench3.json
