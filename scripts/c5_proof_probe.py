#!/usr/bin/env python
"""How often does level 1 of the batched path fail to PROVE a query on BASELINE configs[4] (10 M x 768 un-normalised,
batch 256, top-100 dot), and what does each unproven query cost end to end?  Heap size 16 vs 64 per (slice, query)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

rows, dims, batch, k = 10_000_000, 768, 256, 100
eng = CUDAVectorEngine(VectorMetric.dot, dims)
eng.fill_synthetic(5, rows, normalize=False)
for heap in (0, 64):
    eng.set_option("batch_heap", heap)
    eng.set_option("batch_bf16", 1)
    f0, r0, x0 = eng.counter("batch_filter_bf16_queries"), eng.counter("batch_retry_queries"), eng.batch_stats()[1]
    times = []
    for seed in range(12):
        rng = np.random.default_rng(1000 + seed)
        qs = rng.uniform(-1, 1, size=(batch, dims)).astype(np.float32)
        qs /= np.linalg.norm(qs, axis=1, keepdims=True)
        eng.search_batch_arrays(qs, k)
        t0 = time.perf_counter()
        eng.search_batch_arrays(qs, k)
        times.append((time.perf_counter() - t0) * 1e3)
        eng.set_option("batch_bf16", 1)          # re-arm (the adaptive level choice may have suspended bf16)
    ms, _, _ = eng.time_search_batch(batch, k, 10, warmup=2)
    print(json.dumps({"heap": heap or "auto(16)", "device_ms_per_batch": round(ms / 10, 3),
                      "e2e_ms_per_batch": [round(t, 2) for t in times],
                      "queries": 24 * batch, "bf16_filter_queries": eng.counter("batch_filter_bf16_queries") - f0,
                      "tf32_filter_queries": eng.counter("batch_retry_queries") - r0,
                      "exact_scans": eng.batch_stats()[1] - x0}), flush=True)
