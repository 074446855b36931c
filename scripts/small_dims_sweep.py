#!/usr/bin/env python
"""Single-query scan across row lengths (round 1's dims sweep was below the HBM roofline at 64 ... 256, 1000, 1536, 2048+):
the shipped launch shape and the direct-load kernel (variant 2) per dims.  ~8 GB corpus each, CUDA-event time inside the library, one JSON line per (dims, options)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

only = [int(a) for a in sys.argv[1:]]
CASES = {d: [{}, {"variant": 2}] for d in (32, 64, 96, 100, 128, 160, 256, 300, 384, 400, 512, 640, 768, 1000, 1024, 1280, 1536, 2048, 2560,
                                           3072, 3584, 4096, 6144, 8192)}
for dims, cases in CASES.items():
    if only and dims not in only:
        continue
    rows = int(8e9 // (dims * 4))
    for opts in cases:
        eng = CUDAVectorEngine(VectorMetric.cosine, dims)
        eng.fill_synthetic(3, rows)
        for k_, v_ in opts.items():
            eng.set_option(k_, v_)
        try:
            ms, _ = eng.time_search(10, 10, warmup=3, n_queries=4)
            rec = {"dims": dims, "rows": rows, "options": opts, "ms": round(ms / 10, 4), "gbs": round(rows * dims * 4 / (ms / 10) / 1e6, 1)}
        except Exception as ex:  # noqa: BLE001
            rec = {"dims": dims, "options": opts, "error": repr(ex)[:200]}
        print(json.dumps(rec), flush=True)
        eng.close()
