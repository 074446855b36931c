#!/bin/bash
# GPU call 33: single_shadow mode test + bench.py with the shadow_filtered arm.
set -u
mkdir -p gpurun_out; OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_batch.py -q -k "single_shadow or retry_on_tf32" 2>&1 | tail -8
timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | tee $OUT/bench_r33.json | cut -c1-3000
