"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL; lives under tests/ because it uses the oracle
as the checker -- not collected by pytest): the row-sharded search must return exactly
what a single exact scan of the whole corpus returns (bit-exact against the streaming oracle)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as o  # noqa: E402  (checker)
from wax_b200 import VectorMetric, sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total, dims, seed = 2_000_003, 384, 31
eng = sharded.ShardedVectorEngine(VectorMetric.cosine, dims, total_rows=total)
eng.fill_synthetic(seed)
ok = True
for qi, k in ((0, 10), (1, 32), (2, 72), (3, 1)):
    q = o.synth_row(777, qi, dims, True)
    got = eng.search(q, k)
    rows, d, s = o.search_synth(o.COSINE, seed, 0, total, dims, True, q, k, mode=o.ACC_F32_TREE, threads=16)
    same = [g[0] for g in got] == rows.tolist() and np.array_equal(np.float32([g[1] for g in got]).view(np.uint32), s.view(np.uint32))
    ok = ok and same
    if rank == 0:
        print(f"k={k}: {'OK' if same else 'MISMATCH'} top1={got[0]}", flush=True)
qs = o.synth_rows(778, 0, 4, dims, normalize=True)
many = eng.finish_many(eng.search_many_async(torch.from_numpy(qs).cuda(), 10, slot=0))
same = many == [eng.search(q, 10) for q in qs]
ok = ok and same
if rank == 0:
    print(f"micro-batched exchange (4 queries, one all-gather): {'OK' if same else 'MISMATCH'}", flush=True)
# batched form: tensor-core levels per shard, one all-gather for the whole batch, vectorised merge
qb = o.synth_rows(779, 0, 300, dims, normalize=True)
batch = eng.search_batch(qb, 10)
same = all(batch[i] == eng.search(qb[i], 10) for i in (0, 1, 127, 128, 299))
rows, d, s = o.search_synth(o.COSINE, seed, 0, total, dims, True, qb[5], 10, mode=o.ACC_F32_TREE, threads=16)
same = same and [g[0] for g in batch[5]] == rows.tolist() and \
    np.array_equal(np.float32([g[1] for g in batch[5]]).view(np.uint32), s.view(np.uint32))
ok = ok and same
if rank == 0:
    print(f"sharded search_batch (300 queries, bf16 queries {eng.engine.counter('batch_bf16_queries')}): "
          f"{'OK' if same else 'MISMATCH'}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED PARITY", "PASS" if flag.item() == 1 else "FAIL", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
