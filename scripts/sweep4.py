"""Sweep 4: batched tensor-core path after the epilogue rework; with/without nominations (pipeline floor)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

out_path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "sweep4.json"
only = sys.argv[2] if len(sys.argv) > 2 else "all"
results = []
rows = 10_000_000


def run(eng, label, dims, b, k, iters, noinsert=0):
    eng.set_option("batch_noinsert", noinsert)
    ms, launches, bad = eng.time_search_batch(b, k, iters, warmup=1)
    per = ms / iters
    rec = {"config": label, "batch": b, "noinsert": noinsert, "ms_per_batch": round(per, 3), "qps": round(b / per * 1e3, 1),
           "tflops_useful": round(2.0 * b * rows * dims / (per * 1e-3) / 1e12, 1), "unproven": bad}
    results.append(rec); print(json.dumps(rec), flush=True)
    eng.set_option("batch_noinsert", 0)


eng = CUDAVectorEngine(VectorMetric.cosine, 384)
eng.fill_synthetic(2, rows)
if only == "ncu":
    run(eng, "10Mx384 cos k10", 384, 256, 10, 1)
    sys.exit(0)
for b in (4, 128, 256, 1024):
    run(eng, "10Mx384 cos k10", 384, b, 10, 3)
    run(eng, "10Mx384 cos k10", 384, b, 10, 3, noinsert=1)
eng.close()
eng = CUDAVectorEngine(VectorMetric.dot, 768)
eng.fill_synthetic(5, rows, normalize=False)
run(eng, "10Mx768 dot k100", 768, 256, 100, 3)
run(eng, "10Mx768 dot k100", 768, 256, 100, 3, noinsert=1)
out_path.parent.mkdir(exist_ok=True)
out_path.write_text(json.dumps(results, indent=1))
