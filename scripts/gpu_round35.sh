#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/clustered_probe.py 2>&1 | tee gpurun_out/clustered_probe_r35.jsonl
