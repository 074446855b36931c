#!/usr/bin/env python
"""Small-corpus latency of the fused scan (a real .mv2s vector index holds <= ~174 K rows of 384 floats,
Sources/WaxCore/Constants.swift:49): kernel time back to back and end-to-end call latency against the launch shape
(grid = CTAs, warps per CTA) -- what the fixed cost of a launch is made of."""
import itertools
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

rng = np.random.default_rng(0)
q = rng.standard_normal(384).astype(np.float32)
for rows in (10_000, 100_000, 174_000):
    eng = CUDAVectorEngine(VectorMetric.cosine, 384)
    eng.fill_synthetic(2, rows)
    for tail, grid, warps, chunk in itertools.product((0, 1), (0, 16, 32, 64, 96), (0, 4, 16), (8, 2)):
        eng.set_option("grid", grid); eng.set_option("warps", warps); eng.set_option("chunk_steps", chunk)
        eng.set_option("tail_select", tail)
        n = 300
        ms, _ = eng.time_search(10, n, warmup=10, n_queries=8)
        ms72, _ = eng.time_search(72, n, warmup=10, n_queries=8)
        rec = {"rows": rows, "tail_select": tail, "grid": grid or 148, "warps": warps or 8, "chunk_steps": chunk,
               "kernel_us_k10": round(ms / n * 1e3, 2), "kernel_us_k72": round(ms72 / n * 1e3, 2)}
        if grid == 0 and warps == 0 and chunk == 8:
            for delivery, inline in ((1, 1), (1, 0), (0, 0)):
                eng.set_option("host_delivery", delivery); eng.set_option("inline_query", inline)
                for _ in range(20):
                    eng.search(q, 10)
                t0 = time.perf_counter()
                for _ in range(500):
                    eng.search(q, 10)
                rec[f"e2e_us_delivery{delivery}_inline{inline}"] = round((time.perf_counter() - t0) / 500 * 1e6, 2)
            eng.set_option("host_delivery", 1); eng.set_option("inline_query", 1)
            if tail:                                   # what the selection tail does for the sharded form (world 1)
                blob = eng.shard_open(0, 1, 0)
                eng.shard_connect([blob])
                msx, _ = eng.time_shard_search(10, n, warmup=10, n_queries=8)
                rec["kernel_us_k10_with_exchange_world1"] = round(msx / n * 1e3, 2)
                eng.shard_close()
        print(json.dumps(rec), flush=True)
    eng.close()
