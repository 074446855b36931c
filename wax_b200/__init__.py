"""wax_b200 -- B200-native (sm_100a) brute-force vector scan + top-k behind Wax's `VectorSearchEngine` surface.

Scope: the hot path named by BASELINE.json and nothing else.  The numeric work lives in
libwaxvs_cuda.so (wax_b200/csrc, C-ABI in include/wax_vs_cuda.h); this package is the host-side mirror of
the reference interface plus the row-sharded multi-GPU wrapper.
"""
from .engine import (CUDAVectorEngine, CapacityExceeded, EncodingError, InvalidToc, VectorEnginePreference,
                     VectorMetric, VectorSearchSession, WaxError, is_normalized_l2, normalize_l2)

__all__ = [
    "CUDAVectorEngine", "VectorSearchSession", "VectorMetric", "VectorEnginePreference", "WaxError",
    "EncodingError", "CapacityExceeded", "InvalidToc", "normalize_l2", "is_normalized_l2",
]
