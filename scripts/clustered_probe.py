#!/usr/bin/env python
"""How do the nomination levels behave on CLUSTERED embeddings (what real sentence embeddings look like), not the
uniform synthetic corpus of the headline configs?  Corpus: n unit rows = normalise(centre + sigma * noise) around C
random centres; queries: perturbed corpus rows.  Reports, per (C, sigma): queries proven by the bf16 level, retried on
TF32, re-run exactly; ms per 1024-query batch with the default levels and with TF32 only; equality of the results."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dims, batch, k = 384, 1024, 10
rng = np.random.default_rng(11)
for n_centres, sigma in ((200, 0.35), (20_000, 0.35), (200, 0.1), (2_000_000, 0.0)):
    if n_centres >= n:
        corpus = rng.standard_normal((n, dims), dtype=np.float32)          # no structure: the control
    else:
        centres = rng.standard_normal((n_centres, dims), dtype=np.float32)
        centres /= np.linalg.norm(centres, axis=1, keepdims=True)
        corpus = np.empty((n, dims), np.float32)
        for lo in range(0, n, 250_000):
            hi = min(n, lo + 250_000)
            corpus[lo:hi] = centres[rng.integers(0, n_centres, hi - lo)] + \
                sigma / np.sqrt(dims) * rng.standard_normal((hi - lo, dims), dtype=np.float32)
    corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    qs = corpus[rng.integers(0, n, batch)] + 0.2 / np.sqrt(dims) * rng.standard_normal((batch, dims), dtype=np.float32)
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    eng.reserve(n)
    for lo in range(0, n, 500_000):
        eng.add_batch(np.arange(lo, min(n, lo + 500_000), dtype=np.uint64), corpus[lo:lo + 500_000])
    line = {"rows": n, "centres": n_centres, "sigma": sigma}
    results = {}
    for mode, bf16 in (("levels bf16->tf32->exact", 1), ("tf32->exact", 0)):
        eng.set_option("batch_bf16", bf16)
        eng.search_batch_arrays(qs, k)                                   # warm-up (builds norms / shadow)
        eng.set_option("batch_bf16", bf16)                               # re-arm the bf16 level for the timed batch
        t0, f0 = eng.batch_stats(); r0 = eng.counter("batch_retry_queries"); b0 = eng.counter("batch_bf16_queries")
        t = time.perf_counter()
        ids, scores, ns = eng.search_batch_arrays(qs, k)
        dt = time.perf_counter() - t
        t1, f1 = eng.batch_stats()
        results[mode] = (ids.copy(), scores.copy())
        line[mode] = {"ms_per_batch_e2e": round(dt * 1e3, 3), "bf16_level_queries": eng.counter("batch_bf16_queries") - b0,
                      "tf32_retries": eng.counter("batch_retry_queries") - r0, "exact_fallbacks": f1 - f0}
    a, b = results["levels bf16->tf32->exact"], results["tf32->exact"]
    line["identical"] = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)))
    line["top10_score_range_q0"] = [float(a[1][0, 0]), float(a[1][0, k - 1])]
    print(json.dumps(line), flush=True)
    eng.close()
    del corpus
