#!/usr/bin/env python
"""Condense `ncu -i <report> --page raw --csv` into the columns the profiles/ summaries quote.
usage: ncu -i report.ncu-rep --page raw --csv | python scripts/ncu_summary.py > profiles/<name>_summary.csv"""
import csv
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_active.max",
        "sm__cycles_elapsed.max", "gpc__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__cluster_size",
        "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(sys.stdin))
rows = [r for r in rows if r and not r[0].startswith("==")]
hdr = rows[0]
idx = [hdr.index(k) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] for i in idx])
