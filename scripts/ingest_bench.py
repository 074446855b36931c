"""Corpus mutation / export throughput through the C-ABI (the reference's add path is O(N) per vector, its
remove one tail memmove per id).  Library-level numbers use pre-touched caller buffers (pageable, as a Swift `Data` /
`[Float]` would be); the mirror-level serialize()/deserialize() times are reported next to them."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric, _lib as L  # noqa: E402

dims, total, chunk = 384, 2_000_000, 100_000
rng = np.random.default_rng(0)
eng = CUDAVectorEngine(VectorMetric.cosine, dims)
block = rng.standard_normal((chunk, dims)).astype(np.float32)
big = np.ascontiguousarray(np.tile(block, (total // chunk, 1)))            # 3.07 GB pageable, touched
eng.add_batch(np.arange(10**9, 10**9 + 1000, dtype=np.uint64), block[:1000])   # warm-up: staging buffers, streams
eng.remove_batch(np.arange(10**9, 10**9 + 1000, dtype=np.uint64))
t0 = time.perf_counter()
for start in range(0, total, chunk):
    eng.add_batch(np.arange(start, start + chunk, dtype=np.uint64), block)
t_append = time.perf_counter() - t0
e0 = CUDAVectorEngine(VectorMetric.cosine, dims)
e0.reserve(total)                                                        # the same 20 calls into reserved capacity
t0 = time.perf_counter()
for start in range(0, total, chunk):
    e0.add_batch(np.arange(start, start + chunk, dtype=np.uint64), block)
t_append_reserved = time.perf_counter() - t0
e0.close()
e1 = CUDAVectorEngine(VectorMetric.cosine, dims)
e1.reserve(total)
e1.add_batch(np.arange(10, dtype=np.uint64), block[:10]); e1.remove_batch(np.arange(10, dtype=np.uint64))
t0 = time.perf_counter()
e1.add_batch(np.arange(total, dtype=np.uint64), big)                     # one call, 3.07 GB
t_one = time.perf_counter() - t0
t0 = time.perf_counter()
eng.add_batch(np.arange(0, chunk, dtype=np.uint64), block)               # upsert of existing ids (scatter path)
t_upsert = time.perf_counter() - t0
t0 = time.perf_counter()
eng.remove(1000)                                                         # order-preserving delete near the front
t_remove = time.perf_counter() - t0
gone = rng.choice(total - 1, 1000, replace=False).astype(np.uint64)
t0 = time.perf_counter()
n_gone = eng.remove_batch(gone)                                          # 1000 frames, one pass
t_remove_batch = time.perf_counter() - t0
qs = rng.standard_normal((64, dims)).astype(np.float32)
eng.search_batch(qs, 10)                                                 # builds norms + shadow
t0 = time.perf_counter()
eng.add_batch(np.arange(10**7, 10**7 + 1000, dtype=np.uint64), block[:1000])
eng.search_batch(qs, 10)                                                 # caches extended by 1000 rows, not rebuilt
t_append_search = time.perf_counter() - t0
# serialize / deserialize at the C-ABI with a pre-touched caller buffer
n = C.c_uint64(0)
L.lib().wax_vs_serialized_length(eng.handle, C.byref(n))
buf = np.ones(n.value, np.uint8)
out = C.c_uint64(0)
t0 = time.perf_counter()
rc = L.lib().wax_vs_serialize(eng.handle, buf.ctypes.data_as(C.POINTER(C.c_uint8)), buf.size, C.byref(out))
t_ser = time.perf_counter() - t0
assert rc == 0
e2 = CUDAVectorEngine(VectorMetric.cosine, dims)
e2.reserve(eng.count)
t0 = time.perf_counter()
rc = L.lib().wax_vs_deserialize(e2.handle, buf.ctypes.data_as(C.POINTER(C.c_uint8)), buf.size)
t_de = time.perf_counter() - t0
assert rc == 0 and e2.count == eng.count
t0 = time.perf_counter()
blob = eng.serialize()                                                   # the mirror: fresh bytearray, no extra copies
t_ser_mirror = time.perf_counter() - t0
gb = total * dims * 4 / 1e9
probe = (C.c_float * 7)()
assert L.lib().wax_vs_debug_transfer_probe(eng.handle, 1 << 30, probe) == 0
print(json.dumps({
    "rows": total, "dims": dims, "host_threads_for_staging": "min(8, cgroup cores)",
    "append_20x100k_rows_per_s": round(total / t_append), "append_20x100k_gb_per_s": round(gb / t_append, 2),
    "append_20x100k_reserved_gb_per_s": round(gb / t_append_reserved, 2),
    "append_one_call_gb_per_s": round(gb / t_one, 2),
    "upsert_100k_s": round(t_upsert, 4), "remove_one_s": round(t_remove, 4),
    "remove_batch_1000_s": round(t_remove_batch, 4), "remove_batch_removed": int(n_gone),
    "append_1000_then_batch_search_s": round(t_append_search, 4),
    "serialize_s": round(t_ser, 3), "serialize_gb_per_s": round(buf.size / 1e9 / t_ser, 2),
    "deserialize_s": round(t_de, 3), "deserialize_gb_per_s": round(buf.size / 1e9 / t_de, 2),
    "mirror_serialize_s": round(t_ser_mirror, 3), "blob_gb": round(len(blob) / 1e9, 2),
    "transfer_probe_gb_per_s": {"memcpy_1_thread": round(probe[0], 1), "memcpy_staging_threads": round(probe[1], 1),
                                "dma_h2d_pinned": round(probe[2], 1), "dma_d2h_pinned": round(probe[3], 1),
                                "upload_pipeline": round(probe[4], 1), "download_pipeline": round(probe[5], 1),
                                "staging_threads": int(probe[6])},
    "round1": {"append_gb_per_s": 5.46, "upsert_100k_s": 0.0577, "remove_one_s": 0.0043, "serialize_s": 1.813, "deserialize_s": 0.284},
}))
