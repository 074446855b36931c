#!/usr/bin/env python
"""A handful of small-corpus searches (10 K rows, k = 10 and k = 72) for an ncu source-level capture of the kernel's
fixed costs (prologue, block merge, last-CTA merge):  ncu --set full --import-source on -k regex:scan_tma ..."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
eng = CUDAVectorEngine(VectorMetric.cosine, 384)
eng.fill_synthetic(2, rows)
q = np.random.default_rng(0).standard_normal(384).astype(np.float32)
for k in (10, 72, 10, 72, 10, 72):
    eng.search(q, k)
