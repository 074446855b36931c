#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== batch tests"; timeout 600 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -4
echo "== sweep4"; timeout 900 python scripts/sweep4.py $OUT/sweep4.json 2>&1 | tail -12
