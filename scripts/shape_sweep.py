#!/usr/bin/env python
"""Launch shape of the fused scan against corpus size: warps per CTA x tail (merge / selection) at the sizes that matter
-- a real Wax index (<= 174 K rows), the 8-GPU shard of the 10 M corpus (1.25 M rows), the headline (10 M rows)."""
import itertools
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [174_000, 500_000, 1_250_000, 2_500_000, 10_000_000]
for rows in sizes:
    eng = CUDAVectorEngine(VectorMetric.cosine, 384)
    eng.fill_synthetic(2, rows)
    n = 300 if rows <= 2_500_000 else 60
    for tail, warps, stages in itertools.product((1, 0), (0, 12, 16), (0, 3)):
        eng.set_option("tail_select", tail); eng.set_option("warps", warps); eng.set_option("stages", stages)
        ms, _ = eng.time_search(10, n, warmup=10, n_queries=8)
        ms72, _ = eng.time_search(72, n, warmup=10, n_queries=8)
        print(json.dumps({"rows": rows, "tail_select": tail, "warps": warps or 8, "stages": stages or 2,
                          "kernel_us_k10": round(ms / n * 1e3, 2), "gbs_k10": round(rows * 1536 / (ms / n) / 1e6, 1),
                          "kernel_us_k72": round(ms72 / n * 1e3, 2)}), flush=True)
    eng.close()
