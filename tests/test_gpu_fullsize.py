"""GPU, BASELINE.json full sizes (10 M x 384): size-independent properties + a full streaming oracle scan.

The oracle regenerates the synthetic corpus row by row (bit-identical generator), so no 15 GB host buffer is
needed; on the GPU box's host cores the 10 M-row scan takes a few seconds."""
import ctypes as C

import numpy as np
import pytest

from wax_b200 import CUDAVectorEngine, VectorMetric, sharded

pytestmark = pytest.mark.gpu

N_FULL, DIMS, SEED = 10_000_000, 384, 2      # BASELINE configs[1]: "10M x 384 fp32 corpus, 1 query, top-10 cosine"


@pytest.fixture(scope="module")
def full_engine():
    eng = CUDAVectorEngine(VectorMetric.cosine, DIMS)
    eng.fill_synthetic(SEED, N_FULL)
    yield eng
    eng.close()


def test_full_size_top10_bit_exact_against_streaming_oracle(oracle, full_engine):
    q = oracle.synth_row(1002, 0, DIMS, True)
    got = full_engine.search(q, 10)
    rows, d, s = oracle.search_synth(oracle.COSINE, SEED, 0, N_FULL, DIMS, True, q, 10,
                                     mode=oracle.ACC_F32_TREE, threads=oracle.host_threads())
    assert [g[0] for g in got] == rows.tolist()
    assert np.array_equal(np.float32([g[1] for g in got]).view(np.uint32), s.view(np.uint32))
    # fp64 truth for the returned rows: scores within 1e-4, order identical
    exact = [1.0 - oracle.distance(oracle.COSINE, oracle.ACC_F64, q, oracle.synth_row(SEED, r, DIMS, True)) for r in rows]
    assert np.max(np.abs(np.float64([g[1] for g in got]) - np.float64(exact))) <= 1e-4
    assert all(a > b for a, b in zip(exact, exact[1:]))


def test_full_size_k72_and_k100(oracle, full_engine):
    q = oracle.synth_row(1002, 1, DIMS, True)
    top100 = full_engine.search(q, 100)
    assert len(top100) == 100 and all(a[1] >= b[1] for a, b in zip(top100, top100[1:]))
    assert full_engine.search(q, 72) == top100[:72]              # select path is prefix-consistent
    assert full_engine.search(q, 10) == top100[:10]              # fused-list path agrees with the select path
    for fid, score in top100[::9]:                                # re-score sampled hits from regenerated rows
        row = oracle.synth_row(SEED, fid, DIMS, True)
        assert np.float32(1.0) - np.float32(oracle.distance(oracle.COSINE, oracle.ACC_F32_TREE, q, row)) == np.float32(score)


def test_full_size_planted_neighbours_are_found(oracle, full_engine):
    """Rows built to be the query's nearest neighbours must come back first, in the planted order, wherever
    they sit in the 15 GB stream (first row, a middle row, the last row)."""
    q = oracle.synth_row(1003, 0, DIMS, True)
    rng = np.random.default_rng(0)
    planted = {}
    for rank, row in enumerate((N_FULL - 1, 0, 4_999_999, 7_777_777)):
        noise = rng.standard_normal(DIMS).astype(np.float32)
        planted[row] = (q + np.float32(0.02 * (rank + 1)) * noise / np.linalg.norm(noise)).astype(np.float32)
    before = full_engine.search(q, 10)
    full_engine.add_batch(list(planted), np.stack(list(planted.values())))     # upsert in place (ids == rows)
    assert full_engine.count == N_FULL
    got = full_engine.search(q, 10)
    assert [g[0] for g in got[:4]] == list(planted)
    assert [g for g in got[4:]] == [b for b in before if b[0] not in planted][:6]
    for (fid, score), vec in zip(got[:4], planted.values()):
        assert abs(score - (1.0 - oracle.distance(oracle.COSINE, oracle.ACC_F64, q, vec))) <= 1e-4
    full_engine.fill_synthetic(SEED, N_FULL)                                    # restore for other tests


def test_shard_invariance_through_the_device_entry_point(oracle):
    """Three engines holding contiguous shards + wax_vs_search_device + the host merge give exactly the
    single-engine answer (what the NCCL all-gather path computes, here inside one process)."""
    import torch
    from wax_b200 import _lib as L
    total, k = 300_007, 10
    q = oracle.synth_row(1004, 0, DIMS, True)
    single = CUDAVectorEngine(VectorMetric.cosine, DIMS)
    single.fill_synthetic(9, total)
    expect = single.search(q, k)
    d_q = torch.from_numpy(q).cuda()
    parts = []
    for r in range(3):
        lo, hi = sharded.shard_range(total, 3, r)
        eng = CUDAVectorEngine(VectorMetric.cosine, DIMS)
        eng.fill_synthetic(9, hi - lo, first_row=lo, id_base=lo)
        buf = torch.zeros(k * 24, dtype=torch.uint8, device="cuda")
        rc = L.lib().wax_vs_search_device(eng.handle, C.c_void_p(d_q.data_ptr()), 1, k, lo, C.c_void_p(buf.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, L.last_error()
        torch.cuda.synchronize()
        parts.append(buf.cpu().numpy().view(sharded.CAND_DTYPE).copy())
    best = sharded.merge_candidates(np.concatenate(parts), k)
    scores = sharded.score_from_distance(0, best["distance"])
    assert [(int(i), float(s)) for i, s in zip(best["frame_id"], scores)] == expect
    assert best["row"].tolist() == [e[0] for e in expect]          # id_base == shard offset -> id == global row


def test_single_rank_sharded_engine(oracle):
    """ShardedVectorEngine with world_size 1 (no process group): same answer as the plain engine."""
    import torch
    eng = sharded.ShardedVectorEngine(VectorMetric.cosine, DIMS, total_rows=100_000)
    eng.fill_synthetic(11)
    q = oracle.synth_row(1005, 0, DIMS, True)
    got = eng.search(q, 10)
    rows, _, s = oracle.search_synth(oracle.COSINE, 11, 0, 100_000, DIMS, True, q, 10, mode=oracle.ACC_F32_TREE, threads=8)
    assert [g[0] for g in got] == rows.tolist() and np.array_equal(np.float32([g[1] for g in got]), s)


def test_single_rank_micro_batched_exchange(oracle):
    """search_many_async: several queries, one exchange -- same answers as one query at a time."""
    import torch
    eng = sharded.ShardedVectorEngine(VectorMetric.cosine, DIMS, total_rows=60_000)
    eng.fill_synthetic(12)
    qs = oracle.synth_rows(1006, 0, 5, DIMS)
    d_qs = torch.from_numpy(qs).cuda()
    many = eng.finish_many(eng.search_many_async(d_qs, 10, slot=0))
    assert many == [eng.search(q, 10) for q in qs]
    h1 = eng.search_many_async(d_qs[:2], 10, slot=0)
    h2 = eng.search_many_async(d_qs[2:], 10, slot=1)             # two micro-batches in flight
    assert eng.finish_many(h1) + eng.finish_many(h2) == many


def test_single_rank_sharded_search_batch(oracle):
    """ShardedVectorEngine.search_batch (wax_vs_search_batch_device + vectorised merge): identical to one query at a
    time, through the tensor-core levels (the counters say so), with a shard offset in the global rows."""
    import torch
    eng = sharded.ShardedVectorEngine(VectorMetric.cosine, DIMS, total_rows=90_000)
    eng.fill_synthetic(13)
    qs = oracle.synth_rows(1007, 0, 140, DIMS)
    one_by_one = [eng.search(q, 10) for q in qs]
    assert eng.search_batch(qs, 10) == one_by_one
    assert eng.engine.counter("batch_bf16_queries") == 140
    assert eng.search_batch(torch.from_numpy(qs).cuda(), 10) == one_by_one      # device-resident queries
    # global rows: a second engine holding the same rows as rows [50_000, 140_000) of a larger corpus
    eng.row_lo, eng.row_hi = 50_000, 140_000
    ids, scores, ns = eng.search_batch_arrays(qs[:8], 10)
    assert ns.tolist() == [10] * 8 and [int(i) for i in ids[0]] == [g[0] for g in one_by_one[0]]   # frame ids unchanged
    assert np.array_equal(scores[0], np.float32([g[1] for g in one_by_one[0]]))
    # pipelined form: two batches in flight on the worker thread, same answers
    eng.row_lo, eng.row_hi = 0, 90_000
    h1 = eng.search_batch_submit(qs[:70], 10)
    h2 = eng.search_batch_submit(torch.from_numpy(qs[70:]).cuda(), 10)
    i1, s1, n1 = eng.finish_batch(h1)
    i2, s2, n2 = eng.finish_batch(h2)
    got = [[(int(i), float(s)) for i, s in zip(ii, ss)] for ii, ss in zip(np.concatenate([i1, i2]), np.concatenate([s1, s2]))]
    assert got == one_by_one and n1.tolist() == [10] * 70 and n2.tolist() == [10] * 70
    # k larger than the shard: padded by the scan path
    tiny = sharded.ShardedVectorEngine(VectorMetric.cosine, DIMS, total_rows=6)
    tiny.fill_synthetic(14)
    got = tiny.search_batch(qs[:5], 10)
    assert [len(g) for g in got] == [6] * 5 and got == [tiny.search(q, 10) for q in qs[:5]]
