#!/bin/bash
# GPU call 31: ncu --set full of the bf16 nomination kernel + finish kernel (B = 1024, 10 M x 384), and the launch list.
set -u
mkdir -p gpurun_out; OUT=gpurun_out
cat > /tmp/batch_once.py <<'PY'
import sys; sys.path.insert(0, '.')
from wax_b200 import CUDAVectorEngine, VectorMetric
e = CUDAVectorEngine(VectorMetric.cosine, 384); e.fill_synthetic(2, 10_000_000)
print(e.time_search_batch(1024, 10, 2, warmup=1))
PY
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_batch_bf16_r31.csv python /tmp/batch_once.py > $OUT/ncu_l.log 2>&1; tail -1 $OUT/ncu_l.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batch_ -s 2 -c 2 -f -o $OUT/prof_batch_bf16_r31 python /tmp/batch_once.py > $OUT/ncu_full.log 2>&1; tail -2 $OUT/ncu_full.log
ls -la $OUT/*.ncu-rep
