#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== TS-shape tests (tight timeout)"
timeout 150 python -m pytest tests/test_gpu_batch.py -x -q -k "tmem_shape" 2>&1 | tail -15 | tee $OUT/pytest_ts.txt
echo "== TS timing"; timeout 200 python - <<'PY' 2>&1 | tail -12
import sys, json; sys.path.insert(0, '.')
from wax_b200 import CUDAVectorEngine, VectorMetric
e = CUDAVectorEngine(VectorMetric.cosine, 384); e.fill_synthetic(2, 10_000_000)
for ts in (0, 1):
    e.set_option("batch_ts", ts)
    for b in (256, 1024):
        for noins in (0, 1):
            e.set_option("batch_noinsert", noins)
            ms, l, bad = e.time_search_batch(b, 10, 3, warmup=1)
            print(json.dumps({"ts": ts, "batch": b, "noinsert": noins, "ms": round(ms / 3, 3), "qps": round(b / (ms / 3) * 1e3), "tflops": round(2 * b * 1e7 * 384 / (ms / 3 * 1e-3) / 1e12), "unproven": bad}), flush=True)
PY
