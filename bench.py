#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json `metric`):

    queries/sec, cosine top-10 @ 10 M x 384 fp32, single query per step; % of the HBM roofline; 1/2/4/8 GPUs.

A "step" is one pass of the hot path over the corpus for one query (BASELINE configs[1]).
    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      (the CPU restatement of the reference path on the host cores)

N > 1 is STRONG scaling: the same 10 M-row corpus row-sharded N ways (contiguous ranges), the query
replicated, one all-gather (NCCL) of the per-shard top-k per step and a host-side merge -- the only exchange
the path has.

Prints ONE JSON line (rank 0).  Keys follow the driver's contract; see DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ROWS, DIMS, TOP_K = 10_000_000, 384, 10
CORPUS_SEED, QUERY_SEED = 2, 1002
METRIC_NAME = "queries/sec cosine top-10 @ 10Mx384 fp32"
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback ("of fallback")


# ---------------------------------------------------------------------------------------------------------
def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        # NVML in-process first (a sample every 5 ms: the default timed region is only ~100 ms long), the
        # nvidia-smi loop of the recipe as the fallback.
        self.nvml = None
        self.stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self.nvml = (pynvml, handle)
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _poll_nvml(self):
        nv, h = self.nvml
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = int(get_reasons(h)) if get_reasons else 0
                self.rows.append([str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self.stop.wait(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        self.stop.set()
        if self.nvml:
            try:
                self.thread.join(timeout=1)
                self.nvml[0].nvmlShutdown()
            except Exception:
                pass
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if getattr(self, "nvml", None) else "nvidia-smi"}


def host_queries(n: int) -> np.ndarray:
    """Seeded synthetic unit queries built on the host (uniform[-1,1] then L2-normalised, the reference
    benchmark embedder's distribution, RAGBenchmarkSupport.swift:130-156)."""
    rng = np.random.default_rng(QUERY_SEED)
    q = rng.uniform(-1.0, 1.0, size=(n, DIMS)).astype(np.float32)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


def host_mem_available_gb() -> float:
    try:
        for line in Path("/proc/meminfo").read_text().splitlines():
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_arm(rows: int, steps: int, warmup: int, budget_s: float = 25.0):
    """The reference's CPU path, restated (oracle ACC_F32_TREE, all host threads; `kind: port` because the
    Swift/USearch reference cannot be built in this image).  Bounded sample: as many corpus rows as fit the
    host and the time budget, materialised once; q/s is scaled to the full corpus by rows (the scan is linear)."""
    from oracle import oracle as o
    o.build()
    threads = o.host_threads()
    avail = host_mem_available_gb()
    sample_rows = rows
    max_rows_mem = int(max(avail - 8.0, 1.0) * 1e9 * 0.6 / (DIMS * 4))
    sample_rows = max(100_000, min(sample_rows, max_rows_mem, 10_000_000))
    t0 = time.perf_counter()
    corpus = o.synth_rows(CORPUS_SEED, 0, sample_rows, DIMS, normalize=True, threads=threads)
    gen_s = time.perf_counter() - t0
    qs = o.synth_rows(QUERY_SEED, 0, max(steps + warmup, 1), DIMS, normalize=True, threads=1)
    per = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t = time.perf_counter()
        o.search(o.COSINE, corpus, qs[i % len(qs)], TOP_K, mode=o.ACC_F32_TREE, threads=threads)
        dt = time.perf_counter() - t
        if i >= warmup:
            per.append(dt)
        if time.perf_counter() - t_start > budget_s and len(per) >= 3:
            break
    sec_per_query_full = float(np.mean(per)) * (rows / sample_rows)
    return {
        "value": 1.0 / sec_per_query_full, "unit": "queries/s", "cores": threads, "kind": "port",
        "sample": (f"{len(per)} queries x exact scan of {sample_rows} of {rows} rows x {DIMS} (oracle ACC_F32_TREE, "
                   f"{threads} threads, corpus materialised in {gen_s:.1f}s outside the timed region"
                   + (", time scaled by rows" if sample_rows != rows else "") + ")"),
        "ms_per_query_full_corpus": sec_per_query_full * 1e3, "steps_timed": len(per),
    }


def run_reference(args, rank: int):
    if rank != 0:
        return
    base = cpu_reference_arm(args.rows, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": base["value"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": base["steps_timed"], "warmup": args.warmup,
        "ms_per_step": base["ms_per_query_full_corpus"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1)),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference path (exact scan, USearch metric formulas); the shipped reference "
                "CPU engine is USearch HNSW (approximate) and cannot be built here (no Swift toolchain).",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world: int):
    return {
        "workload": f"{args.rows} x {DIMS} fp32 corpus (BASELINE configs[1]), 1 query per step, top-{TOP_K} cosine",
        "rows": args.rows, "dims": DIMS, "top_k": TOP_K,
        "sharding": f"{world} contiguous row shard(s), {args.rows // world} rows per GPU" if world > 1 else "single GPU",
        "l2": f"corpus {args.rows * DIMS * 4 / world / 1e9:.2f} GB per GPU vs 126 MB L2: inputs larger than L2, no flush"
              if args.rows * DIMS * 4 / world > 4 * 126e6 else "corpus per GPU not >> L2: L2 flushed between steps",
        "corpus_seed": CORPUS_SEED, "query_seed": QUERY_SEED,
    }


# ---------------------------------------------------------------------------------------------------------
def run_single(args):
    import torch
    from wax_b200 import CUDAVectorEngine, VectorMetric
    torch.cuda.set_device(0)
    eng = CUDAVectorEngine(VectorMetric.cosine, DIMS, device=0)
    eng.fill_synthetic(CORPUS_SEED, args.rows)
    for key, val in (kv.split("=") for kv in args.opt):
        eng.set_option(key, int(val))
    n_distinct = min(64, args.steps + args.warmup)

    small = args.rows * DIMS * 4 <= 4 * 126e6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    # ---- value: kernel-only, inputs resident in HBM, CUDA events on the launching stream (inside the library)
    torch.cuda.synchronize()
    with ClockSampler(0) as clk:
        if small:   # flush L2 between steps: time each step separately
            ms_total, launches = 0.0, 0
            eng.time_search(TOP_K, 1, warmup=max(args.warmup, 3), n_queries=n_distinct, seed=QUERY_SEED)
            for i in range(args.steps):
                flush.fill_(i & 0xFF); torch.cuda.synchronize()
                ms, ln = eng.time_search(TOP_K, 1, warmup=0, n_queries=1, seed=QUERY_SEED + i)
                ms_total += ms; launches += ln
        else:
            ms_total, launches = eng.time_search(TOP_K, args.steps, warmup=max(args.warmup, 3),
                                                 n_queries=n_distinct, seed=QUERY_SEED)
        torch.cuda.synchronize()
        # ---- e2e: the public call a user makes (wax_vs_search through the mirror): HOST query in, HOST result out
        qs = host_queries(n_distinct)
        for i in range(max(args.warmup, 3)):
            eng.search(qs[i % n_distinct], TOP_K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        last = None
        for i in range(args.steps):
            if small:
                flush.fill_(i & 0xFF)
            last = eng.search(qs[i % n_distinct], TOP_K)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    clocks = clk.summary()
    ms_per_step = ms_total / args.steps
    value = args.steps / (ms_total / 1e3)

    peak, peak_src = measured_peak()
    read_ceiling = eng.stream_read_gbs(5)                # plain LDG.128 read of the same bytes, same box
    alg_bytes = args.rows * DIMS * 4                     # N*D*4 algorithmic bytes per launch
    achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic_r01.json"
    if tp.exists():
        try:
            t = json.loads(tp.read_text())
            if t.get("rows") == args.rows:
                traffic = t.get("dram_bytes_per_launch")
        except Exception:
            pass
    line = {
        "metric": METRIC_NAME, "value": value, "unit": "queries/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "stream_read_ceiling_gbs": read_ceiling, "frac_of_read_ceiling": achieved / read_ceiling if read_ceiling else None,
                     "nominal_hbm_gbs": 7700.0, "frac_of_nominal": achieved / 7700.0,
                     "kernel": "scan_tma_kernel<C=3,cosine> (fused scan+top-k, 1 launch per query)",
                     "algorithmic_bytes_per_launch": alg_bytes},
        "e2e": {"value": args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": DIMS * 4,
                "d2h_bytes_per_step": TOP_K * 24, "ms_per_step": e2e_s / args.steps * 1e3,
                "api": "wax_vs_search (host query -> host ids/scores, synchronous)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "check": {"top1_frame_id": last[0][0] if last else None, "top1_score": last[0][1] if last else None},
    }
    if not args.no_shadow and not small:
        try:                        # an extra measurement must never cost the headline line
            line["shadow_filtered"] = shadow_filtered_arm(eng, args, qs, n_distinct)
        except Exception as ex:     # noqa: BLE001
            line["shadow_filtered"] = {"error": repr(ex)}
    if not args.no_shadow and not small:
        try:                        # BASELINE configs[2] on the same resident corpus: the tensor-bound companion number
            line["batched"] = batched_arm(eng, args)
        except Exception as ex:     # noqa: BLE001
            line["batched"] = {"error": repr(ex)}
    if not args.no_cpu_baseline:
        base = cpu_reference_arm(args.rows, steps=5, warmup=1, budget_s=20.0)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)


def batched_arm(eng, args, batch: int = 1024, steps: int = 10):
    """NOT the headline: BASELINE configs[2] (batch of 1024 queries, top-10 cosine) on the corpus already resident for
    the headline -- tcgen05 nominations over the bf16 shadow + exact fp32 re-score + completeness proof,
    results identical to 1024 single-query scans.  Device-only, CUDA events inside the library; the full companion
    (end to end, configs[4], TF32) is scripts/bench_batch.py."""
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:  # noqa: BLE001
        pass
    ms, launches, unproven = eng.time_search_batch(batch, TOP_K, steps, warmup=2, seed=QUERY_SEED)
    per = ms / steps
    flops = 2.0 * batch * args.rows * DIMS
    tf = flops / (per * 1e-3) / 1e12
    bf16 = eng.counter("shadow_bytes") > 0
    peak = float(peaks.get("bf16_tflops", 2250.0)) if bf16 else 1100.0
    return {
        "workload": f"{args.rows} x {DIMS} fp32 corpus (BASELINE configs[2]), batch {batch}, top-{TOP_K} cosine",
        "value": batch / per * 1e3, "unit": "queries/s", "ms_per_step": per, "steps": steps,
        "dtype": ("bf16" if bf16 else "tf32") + " nominations + f32 exact re-score",
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                     "peak_source": "measured cuBLAS bf16 burst (MEASURED_PEAKS.json bf16_tflops)" if bf16 and "bf16_tflops" in peaks
                     else ("nominal dense bf16" if bf16 else "nominal dense TF32"),
                     "frac_of_sustained": tf / float(peaks["bf16_tflops_sustained"]) if bf16 and "bf16_tflops_sustained" in peaks else None,
                     "useful_flops_per_launch": flops},
        "gpu_launches_per_step": launches / steps, "unproven_queries_last_step": unproven,
    }


def shadow_filtered_arm(eng, args, qs, n_distinct):
    """NOT the headline (`value` / `e2e` above are the fused fp32 scan north_star names): the same single-query
    workload with the opt-in `single_shadow` mode -- the query is first ranked against the bf16 shadow of the corpus
    (half the HBM bytes), the best nominees are re-scored exactly in fp32 and a completeness proof
    guards the result, so ids and score bits are identical to the fp32 scan (checked here on every step)."""
    import torch
    expect = [eng.search(qs[i % n_distinct], TOP_K) for i in range(min(args.steps, n_distinct))]
    eng.set_option("single_shadow", 1)
    try:
        ms, launches, unproven = eng.time_search_batch(1, TOP_K, args.steps, warmup=3, seed=QUERY_SEED)
        for i in range(3):
            eng.search(qs[i % n_distinct], TOP_K)
        f0 = eng.batch_stats()[1]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = [eng.search(qs[i % n_distinct], TOP_K) for i in range(args.steps)]
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        same = all(got[i] == expect[i % len(expect)] for i in range(args.steps))
        return {
            "value": args.steps / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms / args.steps,
            "e2e": {"value": args.steps / e2e_s, "unit": "queries/s", "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches_per_query": launches / args.steps, "identical_to_fp32_scan": bool(same),
            "exact_fallbacks": eng.batch_stats()[1] - f0, "shadow_gb": eng.counter("shadow_bytes") / 1e9,
            "hbm_bytes_per_query": args.rows * DIMS * 2,
            "achieved_gbs_on_shadow_bytes": args.rows * DIMS * 2 / (ms / args.steps / 1e3) / 1e9,
            "note": "opt-in mode (wax_vs_debug_set_option single_shadow=1); bf16 shadow nominates on the tensor path, "
                    "fp32 re-score + proof make the result identical; costs dims*2 B/row of extra HBM",
        }
    finally:
        eng.set_option("single_shadow", 0)


def run_sharded(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from wax_b200 import VectorMetric, sharded
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = sharded.ShardedVectorEngine(VectorMetric.cosine, DIMS, total_rows=args.rows)
    eng.fill_synthetic(CORPUS_SEED)
    for key, val in (kv.split("=") for kv in args.opt):
        eng.engine.set_option(key, int(val))
    n_distinct = min(64, args.steps + args.warmup)
    qs_host = host_queries(n_distinct)
    qs_dev = torch.from_numpy(qs_host).cuda()
    qs_pinned = torch.from_numpy(qs_host).pin_memory()
    warm = max(args.warmup, 3)
    shard_bytes = (eng.row_hi - eng.row_lo) * DIMS * 4
    small = shard_bytes <= 4 * 126e6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    depth = max(1, args.pipeline)      # independent queries in flight (throughput metric): host merge of
                                       # query i overlaps the scan + all-gather of query i+1

    micro = max(1, args.micro)         # queries per exchange: one all-gather carries `micro` queries' candidates

    def run_steps(n, queries_of):
        """n single-query steps, issued in micro-batches; queries_of(i0, g) -> [g, DIMS] device tensor."""
        from collections import deque
        pending, last, i, b = deque(), None, 0, 0
        while i < n:
            g = min(micro, n - i)
            if small:
                flush.fill_(i & 0xFF)
            pending.append(eng.search_many_async(queries_of(i, g), TOP_K, slot=b % depth))
            i += g; b += 1
            if len(pending) == depth:
                last = eng.finish_many(pending.popleft())[-1]
        while pending:
            last = eng.finish_many(pending.popleft())[-1]
        return last

    def resident(i0, g):
        j0 = i0 % n_distinct
        if j0 + g <= n_distinct:
            return qs_dev[j0:j0 + g]                       # a view: no kernel, queries already resident in HBM
        return torch.cat([qs_dev[j0:], qs_dev[: g - (n_distinct - j0)]])

    # ---- value: inputs resident in HBM; per step = local fused kernel + all-gather + D2H + host merge
    run_steps(warm, resident)
    dist.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record()
        last = run_steps(args.steps, resident)
        ev1.record()
        torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())

        # ---- e2e: host query -> H2D -> scan -> all-gather -> D2H -> host merge, every step
        d_qs = [torch.empty((micro, DIMS), dtype=torch.float32, device="cuda") for _ in range(depth)]
        pin_stage = [torch.empty((micro, DIMS), dtype=torch.float32).pin_memory() for _ in range(depth)]
        counter = [0]

        def host_query(i0, g):
            slot = counter[0] % depth
            counter[0] += 1
            for j in range(g):                                    # host queries arrive one by one
                pin_stage[slot][j].copy_(qs_pinned[(i0 + j) % n_distinct])
            d_qs[slot][:g].copy_(pin_stage[slot][:g], non_blocking=True)
            return d_qs[slot][:g]
        run_steps(warm, host_query)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        last = run_steps(args.steps, host_query)
        torch.cuda.synchronize(); dist.barrier()
        e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e.item())
    clocks = clk.summary()
    if rank == 0:
        peak, peak_src = measured_peak()
        ms_per_step = ms_total / args.steps
        achieved = shard_bytes / (ms_per_step / 1e3) / 1e9
        line = {
            "metric": METRIC_NAME, "value": args.steps / (ms_total / 1e3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "note": "per-GPU: shard bytes / whole step time (scan + all-gather + D2H + host merge)",
                         "algorithmic_bytes_per_launch": shard_bytes},
            "e2e": {"value": args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": DIMS * 4,
                    "d2h_bytes_per_step": world * TOP_K * 24, "ms_per_step": e2e_s / args.steps * 1e3,
                    "api": "ShardedVectorEngine.search_async/finish (host query -> host ids/scores on every rank)"},
            "gpu_launches": args.steps, "queries_in_flight": depth * micro,
            "collective": f"1 all_gather_into_tensor of {micro} x {TOP_K * 24} B per rank per {micro} steps (nccl)",
            "clocks": clocks,
            "check": {"top1_frame_id": last[0][0], "top1_score": last[0][1]},
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS, help="override the corpus size (experiments only)")
    ap.add_argument("--opt", action="append", default=[], help="engine tuning option key=value (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shadow", action="store_true", help="skip the extra `shadow_filtered` (opt-in mode) measurement")
    ap.add_argument("--pipeline", type=int, default=2, help="N>1: micro-batches in flight per rank")
    ap.add_argument("--micro", type=int, default=4, help="N>1: queries per exchange (one all-gather carries them all)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if world > 1:
        return run_sharded(args, rank, world, local_rank)
    return run_single(args)


if __name__ == "__main__":
    main()
