#!/usr/bin/env python
"""Filtered search timings at 10 M x 384: unfiltered, the reference's over-fetch pattern, allow-list
(gather path / bitset path), deny-list."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

e = CUDAVectorEngine(VectorMetric.cosine, 384)
e.fill_synthetic(2, 10_000_000)
q = np.random.default_rng(0).standard_normal(384).astype(np.float32)


def t(fn, n=20):
    fn()
    s = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - s) / n * 1e3


rng = np.random.default_rng(1)
small = rng.choice(10_000_000, 1000, replace=False).astype(np.uint64)
big = rng.choice(10_000_000, 5_000_000, replace=False).astype(np.uint64)
deny = rng.choice(10_000_000, 100_000, replace=False).astype(np.uint64)
print(json.dumps({"unfiltered_ms": round(t(lambda: e.search(q, 24)), 3), "overfetch_72_ms": round(t(lambda: e.search(q, 72)), 3),
                  "allow_1000_ids_gather_ms": round(t(lambda: e.search_filtered(q, 24, allow=small)), 3),
                  "allow_5M_ids_bitset_ms": round(t(lambda: e.search_filtered(q, 24, allow=big), 3), 3),
                  "deny_100K_ids_bitset_ms": round(t(lambda: e.search_filtered(q, 24, deny=deny), 5), 3)}))
