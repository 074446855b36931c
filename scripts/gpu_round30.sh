#!/bin/bash
# GPU call 30: compact rare path in the nomination epilogue (200 KB -> 23 KB of SASS): batch tests + sweep.
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== batch + random gpu tests"; timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_random.py -q -o timeout_method=thread --timeout 300 2>&1 | tail -6 | tee $OUT/pytest_gpu_r30.txt
echo "== sweep"; timeout 700 python scripts/sweep_bf16.py 20 > $OUT/sweep_bf16_r30.jsonl 2> $OUT/sweep_bf16_r30.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_bf16_r30.jsonl'):
    d = json.loads(l)
    print(d['config'][:11], d['mode'], 'full %.2f ms %.0f TF' % (d['full']['ms_per_batch'], d['full']['tflops']) if 'full' in d else d.get('error'),
          'gemm %.2f ms' % d['gemm_only']['ms_per_batch'] if 'gemm_only' in d else '', d.get('check'))
PY
tail -3 $OUT/sweep_bf16_r30.err
