#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -o timeout_method=thread --timeout 300 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r37.txt
