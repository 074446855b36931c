#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== batch tests"; timeout 600 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -4
echo "== sweep4 (auto heap shape)"; timeout 900 python scripts/sweep4.py $OUT/sweep4b.json 2>&1 | tail -12
echo "== heap shape comparison + overlap"; timeout 600 python - <<'PY' 2>&1 | tail -20
import sys, json; sys.path.insert(0, '.')
from wax_b200 import CUDAVectorEngine, VectorMetric
e = CUDAVectorEngine(VectorMetric.cosine, 384); e.fill_synthetic(2, 10_000_000)
for heap in (16, 64):
    e.set_option("batch_heap", heap)
    for b in (256, 1024):
        ms, l, bad = e.time_search_batch(b, 10, 3, warmup=1)
        print(json.dumps({"heap": heap, "batch": b, "ms": round(ms / 3, 3), "qps": round(b / (ms / 3) * 1e3), "unproven": bad}), flush=True)
e.set_option("batch_heap", 0)
for rows in (10_000_000, 5_000_000, 1_250_000):
    e.fill_synthetic(2, rows)
    for ov in (0, 1):
        e.set_option("time_overlap", ov)
        ms, l = e.time_search(10, 60, warmup=4, n_queries=8)
        print(json.dumps({"rows": rows, "overlap": ov, "ms": round(ms / 60, 4), "gbs": round(rows * 1536 / (ms / 60) / 1e6, 1)}), flush=True)
PY
