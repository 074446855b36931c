#!/bin/bash
# GPU call 29: epilogue fixes (threshold prefetch, no barrier without scales) + default-on bf16: whole suite, sweep, batch bench.
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -o timeout_method=thread --timeout 300 2>&1 | tail -12 | tee $OUT/pytest_gpu_r29.txt
echo "== sweep"; timeout 500 python scripts/sweep_bf16.py 5 > $OUT/sweep_bf16_r29.jsonl 2> $OUT/sweep_bf16_r29.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_bf16_r29.jsonl'):
    d = json.loads(l)
    print(d['config'][:11], d['mode'], 'full %.2f ms %.0f TF' % (d['full']['ms_per_batch'], d['full']['tflops']) if 'full' in d else d.get('error'),
          'gemm %.2f ms' % d['gemm_only']['ms_per_batch'] if 'gemm_only' in d else '', d.get('check'))
PY
tail -3 $OUT/sweep_bf16_r29.err
echo "== batch bench"; timeout 600 python scripts/bench_batch.py 10 2>&1 | tee $OUT/bench_batch_r29.jsonl | cut -c1-400
