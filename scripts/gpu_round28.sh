#!/bin/bash
# GPU call 28: bf16-shadow nominations for the batched path -- parity tests first, then the TF32 / bf16 sweep.
set -u
mkdir -p gpurun_out; OUT=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader | tee $OUT/box_r28.txt
echo "== bf16 batch tests"
timeout 600 python -m pytest tests/test_gpu_batch.py -q -k "bf16" -o timeout_method=thread --timeout 150 2>&1 | tail -25 | tee $OUT/pytest_bf16.txt
echo "== sweep"
timeout 500 python scripts/sweep_bf16.py 5 > $OUT/sweep_bf16.jsonl 2> $OUT/sweep_bf16.err
cat $OUT/sweep_bf16.jsonl; tail -5 $OUT/sweep_bf16.err
