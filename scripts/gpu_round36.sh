#!/bin/bash
# GPU call 36: filter level -- batch tests, clustered probe again.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_random.py -q -o timeout_method=thread --timeout 200 2>&1 | tail -15 | tee gpurun_out/pytest_r36.txt
timeout 900 python scripts/clustered_probe.py 2>&1 | tee gpurun_out/clustered_probe_r36.jsonl
