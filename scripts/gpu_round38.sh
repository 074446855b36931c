#!/bin/bash
# GPU call 38: compute-sanitizer memcheck over the new batched kernels (bf16 / ARES / pair / FILTER) on small cases.
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from wax_b200 import CUDAVectorEngine, VectorMetric
rng = np.random.default_rng(0)
for dims, n, b, k, opts in ((384, 3000, 130, 10, {}), (384, 3000, 260, 10, {"batch_pair": 1}), (128, 700, 9, 10, {"batch_ares": 0}),
                            (768, 2000, 40, 100, {}), (384, 3000, 130, 10, {"batch_bf16": 0})):
    e = CUDAVectorEngine(VectorMetric.cosine, dims)
    cent = rng.standard_normal((3, dims)).astype(np.float32)
    rows = cent[rng.integers(0, 3, n)] + np.float32(0.005) * rng.standard_normal((n, dims)).astype(np.float32)   # tight clusters -> filter level
    rows[: n // 2] = rng.standard_normal((n // 2, dims)).astype(np.float32)
    e.add_batch(list(range(n)), rows)
    for key, v in opts.items():
        e.set_option(key, v)
    qs = rows[rng.integers(0, n, b)] + np.float32(0.01) * rng.standard_normal((b, dims)).astype(np.float32)
    got = e.search_batch(qs, k)
    e.set_option("batch_tensor", 0)
    ref = [e.search(q, k) for q in qs[:6]]
    assert got[:6] == ref, (dims, n, b, k, opts)
    print(dims, n, b, k, opts, "bf16", e.counter("batch_bf16_queries"), "filter", e.counter("batch_retry_queries"), "exact", e.batch_stats()[1], flush=True)
    e.close()
print("SANITIZER CASES DONE")
PY
timeout 120 python /tmp/san_case.py 2>&1 | tail -7
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_memcheck_r38.txt 2>&1; echo "rc=$?"
tail -25 gpurun_out/sanitizer_memcheck_r38.txt
