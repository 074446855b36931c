"""Corpus mutation throughput through the C-ABI (the reference's add path is O(N) per vector)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

dims, total, chunk = 384, 2_000_000, 100_000
rng = np.random.default_rng(0)
eng = CUDAVectorEngine(VectorMetric.cosine, dims)
block = rng.standard_normal((chunk, dims)).astype(np.float32)
t0 = time.perf_counter()
for start in range(0, total, chunk):
    eng.add_batch(np.arange(start, start + chunk, dtype=np.uint64), block)
t_append = time.perf_counter() - t0
t0 = time.perf_counter()
eng.add_batch(np.arange(0, chunk, dtype=np.uint64), block)              # upsert of existing ids (scatter path)
t_upsert = time.perf_counter() - t0
t0 = time.perf_counter()
eng.remove(1000)                                                        # order-preserving delete near the front
t_remove = time.perf_counter() - t0
t0 = time.perf_counter()
blob = eng.serialize()
t_ser = time.perf_counter() - t0
e2 = CUDAVectorEngine(VectorMetric.cosine, dims)
t0 = time.perf_counter()
e2.deserialize(blob)
t_de = time.perf_counter() - t0
print(json.dumps({"rows": total, "dims": dims, "append_rows_per_s": round(total / t_append), "append_gb_per_s": round(total * dims * 4 / t_append / 1e9, 2),
                  "upsert_100k_s": round(t_upsert, 4), "remove_one_s": round(t_remove, 4),
                  "serialize_s": round(t_ser, 3), "deserialize_s": round(t_de, 3), "blob_gb": round(len(blob) / 1e9, 2)}))
