#!/usr/bin/env python
"""BASELINE configs[2] (10 M x 384, batch 1024, top-10 cosine) row-sharded over the ranks of a torchrun launch:
every rank nominates on its shard (tensor-core levels), one all-gather of batch x k candidates, one host merge.
Rank 0 prints one JSON line; time = wall clock between barriers + device synchronisation, max over ranks."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from wax_b200 import VectorMetric, sharded  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows, dims, batch, k = 10_000_000, 384, 1024, 10
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
eng = sharded.ShardedVectorEngine(VectorMetric.cosine, dims, total_rows=rows)
eng.fill_synthetic(2)
rng = np.random.default_rng(1)
qs = rng.uniform(-1, 1, size=(batch, dims)).astype(np.float32)
d_qs = torch.from_numpy(qs).cuda()
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2     # batches in flight: 1 = synchronous search_batch_arrays
for _ in range(3):
    eng.search_batch_arrays(d_qs, k)
sync_ids = eng.search_batch_arrays(d_qs, k)[0]


def run(n):
    if depth <= 1:
        out = None
        for _ in range(n):
            out = eng.search_batch_arrays(d_qs, k)
        return out
    from collections import deque
    pending, out = deque(), None
    for _ in range(n):
        pending.append(eng.search_batch_submit(d_qs, k))
        if len(pending) == depth:
            out = eng.finish_batch(pending.popleft())
    while pending:
        out = eng.finish_batch(pending.popleft())
    return out


run(2)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
ids, scores, ns = run(steps)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
per = float(dt.item()) / steps
assert np.array_equal(ids, sync_ids), "pipelined batches returned different ids than the synchronous call"
if rank == 0:
    print(json.dumps({"metric": "queries/sec (batched, row-sharded)", "n_gpus": world, "value": batch / per, "unit": "queries/s",
                      "ms_per_batch": per * 1e3, "config": {"workload": "10M x 384 fp32, batch 1024, top-10 cosine", "rows_per_gpu": rows // world},
                      "exchange_bytes_per_rank": batch * k * 24, "batches_in_flight": depth, "check_top1": [int(ids[0, 0]), float(scores[0, 0])],
                      "bf16_queries": eng.engine.counter("batch_bf16_queries"),
                      "exact_fallbacks": eng.engine.batch_stats()[1]}), flush=True)
if world > 1:
    dist.destroy_process_group()
