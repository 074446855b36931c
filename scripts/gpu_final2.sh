#!/bin/bash
# Final single-GPU validation of the round (after the bf16 / filter-level work): smoke, full GPU suite, bench pair,
# batch bench (bf16 and TF32 nominations), ncu launch list of the bench command.
set -u
mkdir -p gpurun_out; OUT=gpurun_out
{ cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv,noheader; } > $OUT/box_final2.txt; cat $OUT/box_final2.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -o timeout_method=thread --timeout 300 2>&1 | tail -4 | tee $OUT/pytest_gpu_final2.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_final2.json | cut -c1-400
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | tee $OUT/bench_reference_final2.json | cut -c1-300
echo "== batch bench bf16"; timeout 600 python scripts/bench_batch.py 20 2>&1 | tee $OUT/bench_batch_final2.jsonl | cut -c1-260
echo "== batch bench tf32"; timeout 600 python scripts/bench_batch.py 20 tf32 2>&1 | tee $OUT/bench_batch_tf32_final2.jsonl | cut -c1-260
echo "== ncu launch list of bench.py"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_final2.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_final2.log 2>&1
grep -c scan_tma $OUT/launches_final2.csv
