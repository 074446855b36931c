"""Second-phase sweep around the sweet spot found in profiles/sweep_r01_call1.json, with repetitions."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
out_path = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "sweep2.json"
dims, k, iters = 384, 10, 40
eng = CUDAVectorEngine(VectorMetric.cosine, dims)
eng.fill_synthetic(2, rows)
bytes_per = rows * dims * 4
print("stream read ceiling GB/s:", [round(eng.stream_read_gbs(5), 1) for _ in range(3)], flush=True)
results = []
configs = [
    {}, {"rows_per_step": 4, "stages": 2, "warps": 8}, {"rows_per_step": 8, "stages": 2, "warps": 4},
    {"rows_per_step": 4, "stages": 2, "warps": 6}, {"rows_per_step": 4, "stages": 2, "warps": 7},
    {"rows_per_step": 4, "stages": 2, "warps": 9}, {"rows_per_step": 4, "stages": 2, "warps": 10},
    {"rows_per_step": 8, "stages": 2, "warps": 3}, {"rows_per_step": 8, "stages": 2, "warps": 5},
    {"rows_per_step": 8, "stages": 2, "warps": 6}, {"rows_per_step": 4, "stages": 3, "warps": 6},
    {"rows_per_step": 4, "stages": 2, "warps": 4, "grid": 296}, {"rows_per_step": 4, "stages": 2, "warps": 8, "grid": 296},
    {"rows_per_step": 8, "stages": 2, "warps": 2, "grid": 296}, {"rows_per_step": 4, "stages": 2, "warps": 8, "l2_hint": 1},
    {"rows_per_step": 4, "stages": 2, "warps": 8, "grid": 147}, {"rows_per_step": 4, "stages": 2, "warps": 8, "grid": 144},
]
for rep in range(3):
    for opts in configs:
        for key in ("variant", "rows_per_step", "stages", "warps", "grid", "l2_hint"):
            eng.set_option(key, opts.get(key, 0))
        ms, _ = eng.time_search(k, iters, warmup=3, n_queries=8)
        per = ms / iters
        rec = dict(opts, rep=rep, ms=round(per, 4), gbs=round(bytes_per / per / 1e6, 1))
        results.append(rec)
        print(json.dumps(rec), flush=True)
# k sensitivity on the default config
for key in ("variant", "rows_per_step", "stages", "warps", "grid", "l2_hint"):
    eng.set_option(key, 0)
for kk in (1, 10, 32, 33, 72, 100, 1000, 10000):
    ms, launches = eng.time_search(kk, 10, warmup=2, n_queries=4)
    rec = {"k": kk, "ms": round(ms / 10, 4), "launches_per_query": launches / 10}
    results.append(rec)
    print(json.dumps(rec), flush=True)
out_path.parent.mkdir(exist_ok=True)
out_path.write_text(json.dumps({"rows": rows, "results": results}, indent=1))
