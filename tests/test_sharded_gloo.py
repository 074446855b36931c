"""CPU, world_size 2, gloo: the N>1 host path (shard ranges, all-gather of per-shard candidates, host merge)
with the local scan injected (the oracle stands in for the GPU step -- test infrastructure only)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, dims, k, seed, q, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import wax_b200
        from oracle import oracle as o
        from wax_b200 import sharded
        lo, hi = sharded.shard_range(total, world, rank)

        def local_search(query, kk):
            rows, d, _ = o.search_synth(o.COSINE, seed, lo, hi - lo, dims, True, query, kk, mode=o.ACC_F32_TREE)
            c = np.zeros(kk, sharded.CAND_DTYPE)
            c["distance"][: rows.size] = d
            c["row"][: rows.size] = rows
            c["frame_id"][: rows.size] = rows + 7          # ids distinct from rows
            c["valid"][: rows.size] = 1
            return c

        eng = sharded.ShardedVectorEngine(wax_b200.VectorMetric.cosine, dims, total_rows=total,
                                          local_search=local_search)
        assert (eng.row_lo, eng.row_hi) == (lo, hi)
        hits = eng.search(q, k)
        np.save(Path(out_dir) / f"r{rank}.npy", np.array(hits, dtype=np.float64))
        # the batched form: one exchange for the whole batch, vectorised merge -- the same answer per query
        qs = np.stack([q, o.synth_row(78, 0, dims, True), q * np.float32(2.0)])
        batch = eng.search_batch(qs, k)
        assert batch[0] == hits and batch[2] == hits
        ids, scores, ns = eng.search_batch_arrays(qs, k)
        assert ids.shape == (3, min(k, total)) and ns.tolist() == [len(b) for b in batch]
        np.save(Path(out_dir) / f"b{rank}.npy", np.array(batch[1], dtype=np.float64))
        # the filtered collective search exists only on the fused peer-memory transport: a loud error here, not a fallback
        try:
            eng.search_filtered(q, k, allow=[1, 2, 3])
            raise AssertionError("search_filtered must refuse the all-gather transport")
        except wax_b200.InvalidToc:
            pass
        # pipelined form: three batches of different sizes in flight on the worker thread -- the all-gathers must
        # line up across the ranks (submission order), and each batch must come back as the synchronous call has it
        handles = [eng.search_batch_submit(qs[:n], k) for n in (3, 1, 2)]
        for n, h in zip((3, 1, 2), handles):
            ids2, scores2, ns2 = eng.finish_batch(h)
            assert np.array_equal(ids2, ids[:n]) and np.array_equal(scores2, scores[:n]) and ns2.tolist() == ns[:n].tolist()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,k", [(5001, 10), (3, 10), (2000, 72)])
def test_two_rank_sharded_search_equals_single_scan(tmp_path, oracle, total, k):
    o, dims, seed = oracle, 32, 21
    q = o.synth_row(77, 0, dims, True)
    mp.spawn(_worker, args=(2, _free_port(), total, dims, k, seed, q, str(tmp_path)), nprocs=2, join=True)
    rows, d, s = o.search_synth(o.COSINE, seed, 0, total, dims, True, q, k, mode=o.ACC_F32_TREE)
    for rank in range(2):                      # every rank holds the full merged answer
        got = np.load(tmp_path / f"r{rank}.npy").reshape(-1, 2)
        assert got[:, 0].astype(np.int64).tolist() == (rows.astype(np.int64) + 7).tolist()
        assert np.array_equal(got[:, 1].astype(np.float32), s)
    q2 = o.synth_row(78, 0, dims, True)
    rows2, _, s2 = o.search_synth(o.COSINE, seed, 0, total, dims, True, q2, k, mode=o.ACC_F32_TREE)
    for rank in range(2):
        got = np.load(tmp_path / f"b{rank}.npy").reshape(-1, 2)
        assert got[:, 0].astype(np.int64).tolist() == (rows2.astype(np.int64) + 7).tolist()
        assert np.array_equal(got[:, 1].astype(np.float32), s2)
