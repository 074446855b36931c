#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu10.txt
echo "== latency"; timeout 600 python scripts/latency.py $OUT/latency.json 2>&1 | tail -8
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 2>&1 | tail -1 | tee $OUT/bench10.json | cut -c1-300
echo "== ncu launches (final default config)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_r01b.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
grep -c scan_tma $OUT/launches_r01b.csv
echo "== ncu full scan kernel (final default config)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 2 -c 2 -f -o $OUT/prof_scan_r01b \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
echo "== ncu full batch kernel B=1024"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batch_ -s 2 -c 2 -f -o $OUT/prof_batch_r01b \
    python -c "
import sys; sys.path.insert(0,'.')
from wax_b200 import CUDAVectorEngine, VectorMetric
e=CUDAVectorEngine(VectorMetric.cosine,384); e.fill_synthetic(2,10_000_000)
print(e.time_search_batch(1024,10,1,warmup=1))" > $OUT/ncu_batch.log 2>&1
tail -2 $OUT/ncu_batch.log
