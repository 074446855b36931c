#!/bin/bash
# One GPU session: smoke -> parity tests -> bench -> tuning sweep -> ncu.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
{ nvidia-smi; nproc; free -g; } > $OUT/box.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "== pytest gpu"; timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -3 | tee $OUT/bench.json
if [ "${SWEEP:-1}" = "1" ]; then echo "== sweep"; timeout 900 python scripts/sweep.py 10000000 $OUT/sweep.json 2>&1 | tail -60; fi
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launches"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
  tail -3 $OUT/ncu_bench.log
  echo "== ncu full"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 2 -c 2 -o $OUT/prof_scan -f \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
  tail -3 $OUT/ncu_full.log
  ls -la $OUT
fi
