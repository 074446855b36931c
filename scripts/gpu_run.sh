#!/bin/bash
# One parameterised runner for everything that goes to the GPU box (replaces the per-call scripts of round 1).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_run.sh <tag> <step> [<step> ...]'
# Outputs land in gpurun_out/<name>_<tag>.*; copy what should be judged into profiles/.
# Steps:
#   box        host/GPU facts                      smoke      __graft_entry__.smoke()
#   tests      pytest -m gpu (whole suite)         bench      bench.py (N=1) + --impl reference
#   batch      scripts/bench_batch.py (configs[2],[4]; bf16 then tf32)
#   ingest     scripts/ingest_bench.py             latency    scripts/latency.py
#   ncu-small  ncu --set full of the fused scan at 10 K rows (k = 10, k = 72): where the fixed cost goes
#   shape      scripts/shape_sweep.py (warps x tail at 174 K ... 10 M rows)
#   c5-launches  ncu launch list of the configs[4] batch (nominate vs finish kernel time)
#   small-n    scripts/small_n_sweep.py (launch shape vs latency at 10 K .. 174 K rows)
#   batch-sweep  kernel-shape options of the batched path on configs[2] / [4]
#   launches   ncu launch list of bench.py         ncu-scan   ncu --set full of the fused scan kernel
#   ncu-batch  ncu --set full of the batched nominate + finish kernels (configs[2] and [4] shapes)
#   sharded:N  torchrun -N tests/check_sharded_torchrun.py + bench.py --gpus N      (needs gpurun --gpus N)
#   sanitize   compute-sanitizer memcheck over scripts/sanitize_probe.py      ncu-dims  ncu --set full of the scan at 1000 / 1536 dims
#   parity:N   torchrun -N tests/check_sharded_torchrun.py only      batch-sharded:N  scripts/bench_batch_sharded.py (2 and 1 batches in flight)
#   c4         torchrun -8 tests/check_sharded_torchrun.py 100000000 light          (needs gpurun --gpus 8)
#   ncu-batch-c5 / ncu-batch-c3  ncu --set full of the nominate + finish kernels of configs[4] / configs[2] (NCU_OPTS="batch_pair=1" ...)
#   power      clocks + power draw sampled while the batched configs run (is the tensor path power-capped?)
#   sass       cuobjdump opcode histogram of libwaxvs_cuda.so -> profiles-style text
set -u
TAG=${1:?tag}; shift
mkdir -p gpurun_out; OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533"
for step in "$@"; do
  echo "=== $step"
  case "$step" in
    box) { cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc; free -g | sed -n 2p; nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,power.limit --format=csv,noheader; nvidia-smi topo -m 2>/dev/null | head -12; } > $OUT/box_$TAG.txt 2>&1; cat $OUT/box_$TAG.txt ;;
    smoke) timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 ;;
    tests) timeout 1500 python -m pytest tests -q -m gpu --durations=8 > $OUT/pytest_gpu_$TAG.txt 2>&1; tail -14 $OUT/pytest_gpu_$TAG.txt ;;
    bench) timeout 900 python bench.py 2> $OUT/bench_$TAG.err | tail -1 > $OUT/bench_$TAG.json; cut -c1-600 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
           timeout 600 python bench.py --impl reference 2>/dev/null | tail -1 > $OUT/bench_reference_$TAG.json; cut -c1-300 $OUT/bench_reference_$TAG.json ;;
    batch) timeout 600 python scripts/bench_batch.py 20 2>&1 | tee $OUT/bench_batch_$TAG.jsonl | cut -c1-300
           timeout 600 python scripts/bench_batch.py 20 tf32 2>&1 | tee $OUT/bench_batch_tf32_$TAG.jsonl | cut -c1-300 ;;
    ingest) timeout 900 python scripts/ingest_bench.py 2> $OUT/ingest_$TAG.err | tail -1 | tee $OUT/ingest_$TAG.json; tail -3 $OUT/ingest_$TAG.err
            WAXVS_TRACE_INGEST=1 timeout 900 python scripts/ingest_bench.py 2> $OUT/ingest_trace_$TAG.txt > /dev/null; grep -c waxvs $OUT/ingest_trace_$TAG.txt ;;
    ncu-small) timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 4 -c 2 -o $OUT/ncu_small_$TAG -f \
                python scripts/small_n_probe.py 10000 > $OUT/ncu_small_$TAG.log 2>&1
              ncu -i $OUT/ncu_small_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/ncu_small_${TAG}_summary.csv; cut -c1-300 $OUT/ncu_small_${TAG}_summary.csv ;;
    shape) timeout 900 python scripts/shape_sweep.py 2>&1 | tee $OUT/shape_sweep_$TAG.jsonl | cut -c1-200 ;;
    c5-launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/ncu_launches_c5_$TAG.csv \
                python scripts/bench_batch.py 3 bf16 only=1 > $OUT/ncu_launches_c5_$TAG.log 2>&1; grep -E "batch_|shadow" $OUT/ncu_launches_c5_$TAG.csv | tail -8 | cut -c1-260 ;;
    c5-proof) timeout 600 python scripts/c5_proof_probe.py 2>&1 | tee $OUT/c5_proof_$TAG.jsonl | cut -c1-400 ;;
    phases) timeout 300 python scripts/phase_trace.py 2>&1 | tee $OUT/phase_trace_$TAG.jsonl ;;
    small-n) timeout 900 python scripts/small_n_sweep.py 2>&1 | tee $OUT/small_n_$TAG.jsonl | cut -c1-260 ;;
    batch-sweep) for o in "only=1 batch_pair=1" "only=1 batch_pair=1 batch_heap=16" "only=1 batch_heap=16" "only=0 batch_pair=1" "only=0 batch_ares=0" "only=0 batch_pair=1 batch_ares=0"; do
             timeout 400 python scripts/bench_batch.py 10 bf16 $o 2>&1 | tail -1 | tee -a $OUT/batch_sweep_$TAG.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['options'], d['ms_per_batch'], d['roofline']['frac'], d['exact_fallback_queries'])"; done ;;
    latency) timeout 600 python scripts/latency.py 2>&1 | tail -1 | tee $OUT/latency_$TAG.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/ncu_launches_$TAG.csv \
                python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_$TAG.log 2>&1; grep -c scan_tma $OUT/ncu_launches_$TAG.csv ;;
    ncu-scan) timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 4 -c 1 -o $OUT/ncu_scan_$TAG -f \
                python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-shadow > $OUT/ncu_scan_$TAG.log 2>&1
              ncu -i $OUT/ncu_scan_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/ncu_scan_${TAG}_summary.csv; cat $OUT/ncu_scan_${TAG}_summary.csv | cut -c1-400 ;;
    ncu-batch) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'batch_nominate|batch_finish' -s 6 -c 2 -o $OUT/ncu_batch_$TAG -f \
                python scripts/bench_batch.py 2 > $OUT/ncu_batch_$TAG.log 2>&1
              ncu -i $OUT/ncu_batch_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/ncu_batch_${TAG}_summary.csv; cut -c1-400 $OUT/ncu_batch_${TAG}_summary.csv ;;
    ncu-batch-c5) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'batch_nominate|batch_finish' -s 4 -c 2 -o $OUT/ncu_batch_c5_$TAG -f \
                python scripts/bench_batch.py 2 bf16 only=1 > $OUT/ncu_batch_c5_$TAG.log 2>&1
              ncu -i $OUT/ncu_batch_c5_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/ncu_batch_c5_${TAG}_summary.csv; cut -c1-400 $OUT/ncu_batch_c5_${TAG}_summary.csv ;;
    ncu-batch-c3) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'batch_nominate|batch_finish' -s 4 -c 2 -o $OUT/ncu_batch_c3_$TAG -f \
                python scripts/bench_batch.py 2 bf16 only=0 ${NCU_OPTS:-} > $OUT/ncu_batch_c3_$TAG.log 2>&1
              ncu -i $OUT/ncu_batch_c3_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/ncu_batch_c3_${TAG}_summary.csv; cut -c1-400 $OUT/ncu_batch_c3_${TAG}_summary.csv ;;
    power) # SM clock / power draw / throttle reasons sampled every 50 ms WHILE the batched configs run 60 steps each
           nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw,power.limit,clocks_throttle_reasons.active,temperature.gpu --format=csv,noheader -lms 50 > $OUT/power_$TAG.csv 2>&1 &
           SMI=$!
           timeout 600 python scripts/bench_batch.py 60 bf16 2>&1 | tee $OUT/bench_batch_power_$TAG.jsonl | cut -c1-260
           kill $SMI; sort -t, -k3 -n $OUT/power_$TAG.csv | tail -3 ;;
    sharded:*) N=${step#sharded:}
           timeout 900 $TR --nproc-per-node $N tests/check_sharded_torchrun.py > $OUT/sharded_parity_${TAG}_n$N.txt 2>&1; tail -12 $OUT/sharded_parity_${TAG}_n$N.txt
           timeout 900 $TR --nproc-per-node $N bench.py --gpus $N 2> $OUT/bench_${TAG}_n$N.err | tail -1 > $OUT/bench_${TAG}_n$N.json; cut -c1-700 $OUT/bench_${TAG}_n$N.json; tail -3 $OUT/bench_${TAG}_n$N.err ;;
    sanitize) timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py > $OUT/sanitizer_memcheck_$TAG.txt 2>&1; echo "rc=$?" >> $OUT/sanitizer_memcheck_$TAG.txt; tail -8 $OUT/sanitizer_memcheck_$TAG.txt ;;
    ncu-dims) for d in 1000 1536; do
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 4 -c 1 -o $OUT/ncu_scan_d${d}_$TAG -f python scripts/small_dims_sweep.py $d > $OUT/ncu_scan_d${d}_$TAG.log 2>&1
             ncu -i $OUT/ncu_scan_d${d}_$TAG.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/ncu_scan_d${d}_${TAG}_summary.csv; cut -c1-330 $OUT/ncu_scan_d${d}_${TAG}_summary.csv | tail -1
             rm -f $OUT/ncu_scan_d${d}_$TAG.ncu-rep; done ;;   # the reports are ~30 MB each: only the summaries travel back
    parity:*) N=${step#parity:}
           timeout 900 $TR --nproc-per-node $N tests/check_sharded_torchrun.py > $OUT/sharded_parity_${TAG}_n$N.txt 2>&1; tail -14 $OUT/sharded_parity_${TAG}_n$N.txt ;;
    batch-sharded:*) N=${step#batch-sharded:}
           for depth in 2 1; do timeout 600 $TR --nproc-per-node $N scripts/bench_batch_sharded.py 20 $depth 2> $OUT/bench_batch_sharded_${TAG}_n${N}_d$depth.err | tail -1 | tee $OUT/bench_batch_sharded_${TAG}_n${N}_d$depth.json | cut -c1-400; tail -2 $OUT/bench_batch_sharded_${TAG}_n${N}_d$depth.err; done ;;
    c4) timeout 1200 $TR --nproc-per-node 8 tests/check_sharded_torchrun.py 100000000 light > $OUT/sharded_parity_${TAG}_c4_100m_n8.txt 2>&1; tail -12 $OUT/sharded_parity_${TAG}_c4_100m_n8.txt ;;
    sass) cuobjdump -sass wax_b200/libwaxvs_cuda.so | python scripts/sass_histogram.py > $OUT/sass_opcodes_$TAG.txt; head -40 $OUT/sass_opcodes_$TAG.txt ;;
    *) echo "unknown step $step" ;;
  esac
done
