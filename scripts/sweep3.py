"""Sweep 3: dynamic chunk scheduling of the scan kernel + first timings of the batched tensor-core path."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

out_path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "sweep3.json"
results = {"scan": [], "batch": []}
rows, dims = 10_000_000, 384
eng = CUDAVectorEngine(VectorMetric.cosine, dims)
eng.fill_synthetic(2, rows)
print("stream read ceiling GB/s:", round(eng.stream_read_gbs(5), 1), flush=True)
for rep in range(2):
    for chunk in (0, 4, 8, 16, 32, 64):
        for stages in (2, 3):
            eng.set_option("chunk_steps", chunk); eng.set_option("stages", stages)
            ms, _ = eng.time_search(10, 40, warmup=3, n_queries=8)
            rec = {"chunk_steps": chunk, "stages": stages, "rep": rep, "ms": round(ms / 40, 4),
                   "gbs": round(rows * dims * 4 / (ms / 40) / 1e6, 1)}
            results["scan"].append(rec); print(json.dumps(rec), flush=True)
eng.set_option("chunk_steps", 16); eng.set_option("stages", 0)
# batched path, config 3: 10M x 384, top-10 cosine
for b in (4, 16, 64, 128, 256, 512, 1024):
    iters = 3 if b >= 256 else 5
    ms, launches, bad = eng.time_search_batch(b, 10, iters, warmup=1)
    per = ms / iters
    rec = {"config": "10Mx384 cos k10", "batch": b, "ms_per_batch": round(per, 3), "qps": round(b / per * 1e3, 1),
           "tflops_useful": round(2.0 * b * rows * dims / (per * 1e-3) / 1e12, 1), "unproven": bad, "launches": launches / iters}
    results["batch"].append(rec); print(json.dumps(rec), flush=True)
eng.close()
# config 5: 10M x 768 (rows not normalised), batch 256, top-100 dot
eng = CUDAVectorEngine(VectorMetric.dot, 768)
eng.fill_synthetic(5, rows, normalize=False)
for b in (64, 256):
    ms, launches, bad = eng.time_search_batch(b, 100, 3, warmup=1)
    per = ms / 3
    rec = {"config": "10Mx768 dot k100", "batch": b, "ms_per_batch": round(per, 3), "qps": round(b / per * 1e3, 1),
           "tflops_useful": round(2.0 * b * rows * 768 / (per * 1e-3) / 1e12, 1), "unproven": bad, "launches": launches / 3}
    results["batch"].append(rec); print(json.dumps(rec), flush=True)
ms, _ = eng.time_search(100, 5, warmup=1, n_queries=2)
print(json.dumps({"config": "10Mx768 dot k100 single-query", "ms": round(ms / 5, 3)}), flush=True)
out_path.parent.mkdir(exist_ok=True)
out_path.write_text(json.dumps(results, indent=1))
