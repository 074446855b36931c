#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== sweep2"; timeout 900 python scripts/sweep2.py 10000000 $OUT/sweep2.json 2>&1 | tail -80
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -2 | tee $OUT/bench2.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_r01.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
tail -2 $OUT/ncu_bench.log; grep -c scan_tma $OUT/launches_r01.csv
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 2 -c 2 -f -o $OUT/prof_scan_r01 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ls -la $OUT
