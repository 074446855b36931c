"""Single-query scan throughput across embedding dimensions (fixed ~8 GB corpus), TMA-staged vs direct-load kernel."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import CUDAVectorEngine, VectorMetric  # noqa: E402

out = []
for dims in (64, 100, 128, 256, 384, 512, 768, 1000, 1024, 1536, 2048, 3072, 4096):
    rows = int(8e9 // (dims * 4))
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    eng.fill_synthetic(3, rows)
    rec = {"dims": dims, "rows": rows}
    for name, opts in (("tma", {"variant": 0}), ("ldg", {"variant": 2})):
        for k_, v_ in opts.items():
            eng.set_option(k_, v_)
        ms, _ = eng.time_search(10, 10, warmup=2, n_queries=4)
        rec[name + "_gbs"] = round(rows * dims * 4 / (ms / 10) / 1e6, 1)
    if dims == 768:
        for r in (2, 4):
            eng.set_option("variant", 0); eng.set_option("rows_per_step", r)
            ms, _ = eng.time_search(10, 10, warmup=2, n_queries=4)
            rec[f"tma_R{r}_gbs"] = round(rows * dims * 4 / (ms / 10) / 1e6, 1)
    out.append(rec); print(json.dumps(rec), flush=True)
    eng.close()
Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "dims_sweep.json").write_text(json.dumps(out, indent=1))
