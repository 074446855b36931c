#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== new tests first"; timeout 600 python -m pytest tests/test_gpu_filtered.py tests/test_gpu_engine.py -x -q 2>&1 | tail -8
echo "== all gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu15.txt
echo "== batch bench"; timeout 600 python scripts/bench_batch.py 5 2>&1 | tee $OUT/bench_batch.jsonl | cut -c1-420
echo "== filtered timing"; timeout 300 python - <<'PY' 2>&1 | tail -8
import sys, json, time; sys.path.insert(0, '.')
import numpy as np
from wax_b200 import CUDAVectorEngine, VectorMetric
e = CUDAVectorEngine(VectorMetric.cosine, 384); e.fill_synthetic(2, 10_000_000)
q = np.random.default_rng(0).standard_normal(384).astype(np.float32)
def t(fn, n=20):
    fn(); s = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - s) / n * 1e3
rng = np.random.default_rng(1)
small = rng.choice(10_000_000, 1000, replace=False).astype(np.uint64)
big = rng.choice(10_000_000, 5_000_000, replace=False).astype(np.uint64)
deny = rng.choice(10_000_000, 100_000, replace=False).astype(np.uint64)
print(json.dumps({"unfiltered_ms": round(t(lambda: e.search(q, 24)), 3), "overfetch_72_ms": round(t(lambda: e.search(q, 72)), 3),
                  "allow_1000_ids_gather_ms": round(t(lambda: e.search_filtered(q, 24, allow=small)), 3),
                  "allow_5M_ids_bitset_ms": round(t(lambda: e.search_filtered(q, 24, allow=big), 3), 3),
                  "deny_100K_ids_bitset_ms": round(t(lambda: e.search_filtered(q, 24, deny=deny), 5), 3)}))
PY
