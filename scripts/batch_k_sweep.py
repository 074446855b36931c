#!/usr/bin/env python
"""Batched path across k on BASELINE configs[2]'s corpus (10 M x 384, batch 1024, cosine): device time per batch, the
nominee-heap shape the engine picked vs forced ones, and how many queries level 1 left unproven (each unproven query
costs its batch a second pass).  One JSON line per (k, options)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

eng = CUDAVectorEngine(VectorMetric.cosine, 384)
eng.fill_synthetic(2, 10_000_000)
rng = np.random.default_rng(5)
for k in (10, 24, 32, 48, 72, 100, 128):
    for opts in ({}, {"batch_heap": 24}, {"batch_heap": 32}, {"batch_heap": 64}):
        for o in ("batch_heap", "batch_pair"):
            eng.set_option(o, 0)
        eng.set_option("batch_bf16", 1)
        for kk, vv in opts.items():
            eng.set_option(kk, vv)
        ms, launches, bad = eng.time_search_batch(1024, k, 10, warmup=2)
        f0, x0 = eng.counter("batch_filter_bf16_queries"), eng.batch_stats()[1]
        e2e = []
        for _ in range(3):
            qs = rng.uniform(-1, 1, size=(1024, 384)).astype(np.float32)
            qs /= np.linalg.norm(qs, axis=1, keepdims=True)
            t = time.perf_counter()
            eng.search_batch_arrays(qs, k)
            e2e.append(round((time.perf_counter() - t) * 1e3, 2))
            eng.set_option("batch_bf16", 1)
        print(json.dumps({"k": k, "options": opts, "heap": eng.counter("batch_last_heap"), "bump": eng.counter("batch_heap_bump"), "device_ms": round(ms / 10, 3), "unproven_last_step": bad, "e2e_ms_3_fresh_batches": e2e,
                          "bf16_filter_queries": eng.counter("batch_filter_bf16_queries") - f0, "exact_scans": eng.batch_stats()[1] - x0}), flush=True)
