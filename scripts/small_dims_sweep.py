#!/usr/bin/env python
"""Single-query scan at the row lengths where round 1's dims sweep was below the HBM roofline (64, 128, 256, 1000, 1536):
launch-shape options per dims.  ~8 GB corpus each, CUDA-event time inside the library, one JSON line per (dims, options)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

only = [int(a) for a in sys.argv[1:]]
CASES = {
    64: [{"variant": 2}, {"variant": 2, "ldg_ctas_per_sm": 8}],
    128: [{}, {"stages": 3}, {"stages": 4}, {"warps": 8}, {"warps": 8, "stages": 4}, {"rows_per_step": 4}, {"rows_per_step": 4, "stages": 4},
          {"chunk_steps": 32}, {"chunk_steps": 0}, {"variant": 2}, {"variant": 2, "ldg_ctas_per_sm": 8}],
    256: [{}, {"stages": 3}, {"warps": 16}, {"warps": 8}, {"rows_per_step": 4}, {"chunk_steps": 32}],
    1000: [{}, {"stages": 3}, {"rows_per_step": 1}, {"variant": 2}],
    1536: [{}, {"rows_per_step": 1}, {"stages": 3}, {"rows_per_step": 1, "stages": 3}, {"variant": 2}],
    768: [{}], 384: [{}], 1024: [{}], 3072: [{}],
}
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
