"""One k=72 search at a given corpus size (ncu target)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402
rows = int(sys.argv[1]); k = int(sys.argv[2])
eng = CUDAVectorEngine(VectorMetric.cosine, 384)
eng.fill_synthetic(2, rows)
q = np.random.default_rng(0).standard_normal(384).astype(np.float32)
for _ in range(3):
    eng.search(q, k)
