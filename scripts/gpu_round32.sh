#!/bin/bash
# GPU call 32: load(from:) mirror test, small-batch curve on the tensor path (incl. 1 query through the bf16 shadow).
set -u
mkdir -p gpurun_out; OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -q -k "load_replays" 2>&1 | tail -3
timeout 600 python scripts/small_batch.py 2>&1 | tee $OUT/small_batch_r32.jsonl
