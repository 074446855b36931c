#!/bin/bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out
echo "== batch tests"; timeout 600 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -8 | tee $OUT/pytest_batch4.txt
echo "== sweep4"; timeout 900 python scripts/sweep4.py $OUT/sweep4.json 2>&1 | tail -30
echo "== ncu batch kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batch_tf32 -s 1 -c 1 -f -o $OUT/prof_batch_r01 \
    python scripts/sweep4.py $OUT/sweep4_ncu.json ncu > $OUT/ncu_batch.log 2>&1
tail -3 $OUT/ncu_batch.log
