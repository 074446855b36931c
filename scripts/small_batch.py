#!/usr/bin/env python
"""Small batches on the tensor path: ms per batch and q/s for 1..256 queries (10 M x 384, top-10 cosine), bf16-shadow
vs TF32 nominations, next to the single-query fused scan.  Below one query group (128) the pass is HBM-bound on the
operand it streams: 7.68 GB (shadow) vs 15.36 GB (fp32 corpus)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
e = CUDAVectorEngine(VectorMetric.cosine, 384)
e.fill_synthetic(2, rows)
e.set_option("batch_min", 1)
ms1, _ = e.time_search(10, 20, warmup=3, n_queries=4)
print(json.dumps({"mode": "single-query fused scan (fp32)", "ms_per_query": ms1 / 20, "qps": 20 / ms1 * 1e3}), flush=True)
for bf16 in (1, 0):
    e.set_option("batch_bf16", bf16)
    for b in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        ms, launches, bad = e.time_search_batch(b, 10, 10, warmup=2)
        per = ms / 10
        print(json.dumps({"mode": "bf16 shadow" if bf16 else "tf32", "batch": b, "ms_per_batch": round(per, 4),
                          "qps": round(b / per * 1e3), "operand_gbs": round(rows * 384 * (2 if bf16 else 4) / per / 1e6),
                          "unproven": bad}), flush=True)
