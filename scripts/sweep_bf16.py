#!/usr/bin/env python
"""Batched path: TF32 vs bf16-shadow nominations (device-only timing, CUDA events inside the library), BASELINE
configs[2] and configs[4].  One JSON line per (config, mode); also cross-checks that the bf16 modes return exactly
what the TF32 mode returns at full size (ids and score bits of 256 queries)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

CONFIGS = [
    dict(name="configs[2]: 10M x 384, batch 1024, top-10 cosine", metric=VectorMetric.cosine, rows=10_000_000,
         dims=384, batch=1024, k=10, normalize=True, seed=2),
    dict(name="configs[4]: 10M x 768, batch 256, top-100 dot", metric=VectorMetric.dot, rows=10_000_000,
         dims=768, batch=256, k=100, normalize=False, seed=5),
]
MODES = [
    ("tf32", dict(batch_bf16=0, batch_pair=0)),
    ("bf16 ares", dict(batch_bf16=1, batch_pair=0, batch_ares=1)),
    ("bf16 stream", dict(batch_bf16=1, batch_pair=0, batch_ares=0)),
    ("bf16 pair ares", dict(batch_bf16=1, batch_pair=1, batch_ares=1)),
    ("bf16 pair stream", dict(batch_bf16=1, batch_pair=1, batch_ares=0)),
]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only = sys.argv[2] if len(sys.argv) > 2 else ""
for cfg in CONFIGS:
    if only and only not in cfg["name"]:
        continue
    eng = CUDAVectorEngine(cfg["metric"], cfg["dims"])
    eng.fill_synthetic(cfg["seed"], cfg["rows"], normalize=cfg["normalize"])
    flops = 2.0 * cfg["batch"] * cfg["rows"] * cfg["dims"]
    rng = np.random.default_rng(1)
    qs = rng.uniform(-1, 1, size=(256, cfg["dims"])).astype(np.float32)
    ref = None
    for name, opts in MODES:
        for key, v in opts.items():
            eng.set_option(key, v)
        line = {"config": cfg["name"], "mode": name}
        try:
            for noins in (0, 1):
                eng.set_option("batch_noinsert", noins)
                ms, launches, bad = eng.time_search_batch(cfg["batch"], cfg["k"], steps, warmup=2)
                per = ms / steps
                tag = "gemm_only" if noins else "full"
                line[tag] = {"ms_per_batch": per, "qps": cfg["batch"] / per * 1e3, "tflops": flops / (per * 1e-3) / 1e12,
                             "launches_per_batch": launches / steps, "unproven": bad}
            eng.set_option("batch_noinsert", 0)
            t0, f0 = eng.batch_stats()
            r0 = eng.counter("batch_retry_queries")
            ids, scores, ns = eng.search_batch_arrays(qs, cfg["k"])
            t1, f1 = eng.batch_stats()
            line["check"] = {"exact_fallbacks": f1 - f0, "tf32_retries": eng.counter("batch_retry_queries") - r0,
                             "bf16_queries": eng.counter("batch_bf16_queries"), "shadow_gb": eng.counter("shadow_bytes") / 1e9}
            if ref is None:
                ref = (ids.copy(), scores.copy())
            else:
                line["check"]["same_ids_as_tf32"] = bool(np.array_equal(ids, ref[0]))
                line["check"]["same_score_bits_as_tf32"] = bool(np.array_equal(scores.view(np.uint32), ref[1].view(np.uint32)))
        except Exception as ex:  # keep the sweep going: one bad mode must not cost the GPU call
            line["error"] = repr(ex)
        print(json.dumps(line), flush=True)
    eng.close()
