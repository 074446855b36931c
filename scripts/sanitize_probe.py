#!/usr/bin/env python
"""Small cases of every kernel family touched in round 2, meant to run under `compute-sanitizer --tool memcheck` (or
racecheck): generic TMA shapes (R = 8 / 4 / 2, ragged dims), the unrolled 1536-dim shape, the direct-load kernel, the 24-entry
heap shape, the four-row exact re-score (finish / gather / filter re-score), masked nominations (filtered batch), the
large-k batch (filter level), the device-side merge of gathered lists.  Prints OK lines; any sanitizer report fails it."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric, sharded  # noqa: E402
from wax_b200 import _lib as L  # noqa: E402

rng = np.random.default_rng(3)
for dims, n in ((36, 3001), (64, 5000), (100, 2777), (300, 4099), (1000, 1501), (1536, 700), (2048, 333), (2560, 257), (4100, 65), (7, 900)):
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    corpus = rng.standard_normal((n, dims)).astype(np.float32)
    eng.add_batch(np.arange(n, dtype=np.uint64), corpus)
    for k in (1, 10, 72, 200):
        got = eng.search(corpus[5], k)
        assert got[0][0] == 5, (dims, k)
    eng.close()
    print("scan OK", dims, flush=True)
dims, n = 384, 40_000
eng = CUDAVectorEngine(VectorMetric.dot, dims)
eng.fill_synthetic(9, n, normalize=False)
qs = rng.standard_normal((130, dims)).astype(np.float32)
eng.set_option("batch_ares", 0); eng.set_option("batch_heap", 24)
a = eng.search_batch(qs, 100)
eng.set_option("batch_heap", 0); eng.set_option("batch_ares", 1)
b = eng.search_batch(qs, 100)
assert a == b
print("batch heap24 OK", flush=True)
deny = np.arange(0, n, 3, dtype=np.uint64)
f = eng.search_batch_filtered(qs, 10, deny=deny)
assert all(i % 3 != 0 for hits in f for i, _ in hits)
g = eng.search_batch_filtered(qs[:9], 10, allow=np.arange(7, 900, 5, dtype=np.uint64))
assert all(len(h) == 10 for h in g)
print("filtered batch OK", flush=True)
big = eng.search_batch(qs[:20], 300)
assert all(len(h) == 300 for h in big)
print("large-k batch OK", flush=True)
import torch  # noqa: E402
world, bq, k = 5, 33, 10
cands = np.zeros((world, bq, k), sharded.CAND_DTYPE)
for r in range(world):
    for q in range(bq):
        d = np.sort(rng.random(k).astype(np.float32))
        cands[r, q]["distance"] = d; cands[r, q]["valid"] = 1
        cands[r, q]["row"] = np.arange(k) + r * 100; cands[r, q]["frame_id"] = cands[r, q]["row"]
dev = torch.from_numpy(cands.view(np.uint8).reshape(-1).copy()).cuda()
out = torch.zeros(bq * k * 24, dtype=torch.uint8, device="cuda")
assert L.lib().wax_vs_merge_candidates_device(eng.handle, C.c_void_p(dev.data_ptr()), world, bq, k, k, C.c_void_p(out.data_ptr()), None) == 0
torch.cuda.synchronize()
got = out.cpu().numpy().view(sharded.CAND_DTYPE).reshape(bq, k)
want, _ = sharded.merge_candidates_batch(cands, k)
assert np.array_equal(got, want)
print("device merge OK", flush=True)
eng.close()
