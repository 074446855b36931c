#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_batch.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_final4.txt
