#!/usr/bin/env python
"""Phase breakdown of one fused search against corpus size and k (wax_vs_debug_phase_trace)."""
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wax_b200 import CUDAVectorEngine, VectorMetric, _lib as L  # noqa: E402

for rows in (10_000, 174_000, 1_250_000, 10_000_000):
    eng = CUDAVectorEngine(VectorMetric.cosine, 384)
    eng.fill_synthetic(2, rows)
    for k in (10, 72):
        out = (C.c_float * 5)()
        assert L.lib().wax_vs_debug_phase_trace(eng.handle, k, 50, out) == 0, L.last_error()
        print(json.dumps({"rows": rows, "k": k, "scan_loop_done_us": round(out[0], 2), "cta_select_done_us": round(out[1], 2),
                          "grid_stage_start_us": round(out[2], 2), "kernel_end_us": round(out[3], 2),
                          "event_duration_us": round(out[4], 2)}), flush=True)
    eng.close()
