#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_final5.txt
