"""Tuning sweep of the fused scan kernel on one GPU (experiments; results land in gpurun_out/)."""
import itertools
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
out_path = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "sweep.json"
dims, k, iters = 384, 10, 20
eng = CUDAVectorEngine(VectorMetric.cosine, dims)
t0 = time.time()
eng.fill_synthetic(2, rows)
print(f"fill {rows} rows: {time.time() - t0:.2f}s", flush=True)
bytes_per = rows * dims * 4
results = []


def run(opts):
    for key in ("variant", "rows_per_step", "stages", "warps", "grid", "l2_hint"):
        eng.set_option(key, opts.get(key, 0))
    if "ldg_ctas_per_sm" in opts:
        eng.set_option("ldg_ctas_per_sm", opts["ldg_ctas_per_sm"])
    try:
        ms, launches = eng.time_search(k, iters, warmup=3, n_queries=8)
    except Exception as exc:  # noqa: BLE001
        print(opts, "FAILED", exc, flush=True)
        return
    per = ms / iters
    rec = dict(opts, ms=per, gbs=bytes_per / per / 1e6, qps=1e3 / per)
    results.append(rec)
    print(json.dumps(rec), flush=True)


run({})  # default config
for R, stages, warps in itertools.product((4, 8), (2, 3, 4, 6), (4, 8, 12, 16)):
    smem = warps * stages * (R * dims * 4 + 8) + warps * 256
    if smem > 232448:
        continue
    run({"variant": 1, "rows_per_step": R, "stages": stages, "warps": warps})
for grid in (74, 148, 296):
    run({"variant": 1, "grid": grid})
run({"variant": 1, "l2_hint": 1})
for cps in (2, 4, 8):
    run({"variant": 2, "ldg_ctas_per_sm": cps})
results.sort(key=lambda r: r["ms"])
out_path.parent.mkdir(exist_ok=True)
out_path.write_text(json.dumps({"rows": rows, "dims": dims, "k": k, "results": results}, indent=1))
print("BEST", json.dumps(results[0]))
