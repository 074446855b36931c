#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_gpu11.txt
echo "== k sweep + latency"; timeout 600 python - <<'PY' 2>&1 | tail -20
import sys, json; sys.path.insert(0, '.')
from wax_b200 import CUDAVectorEngine, VectorMetric
e = CUDAVectorEngine(VectorMetric.cosine, 384); e.fill_synthetic(2, 10_000_000)
for k in (1, 10, 32, 33, 72, 100, 128, 129, 1000):
    ms, l = e.time_search(k, 20, warmup=3, n_queries=4)
    print(json.dumps({"k": k, "ms": round(ms / 20, 4), "launches_per_query": l / 20}), flush=True)
PY
timeout 300 python scripts/latency.py $OUT/latency11.json 2>&1 | head -3
