#!/bin/bash
# Final single-GPU validation of a round: full GPU suite, smoke, headline bench, batch bench.
set -u
mkdir -p gpurun_out; OUT=gpurun_out
{ cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc; } > $OUT/host_cpu.txt; cat $OUT/host_cpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== all gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu_final.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_final.json | cut -c1-300
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | tee $OUT/bench_reference_final.json | cut -c1-300
echo "== batch bench"; timeout 600 python scripts/bench_batch.py 5 2>&1 | tee $OUT/bench_batch_final.jsonl | cut -c1-200
