"""GPU: bulk ingest / export at device speed keeps the reference's mutation semantics
(MetalVectorEngine.swift:330-444, :682-815): chunked pinned double-buffered transfers, the one-pass remove_batch
compaction, and the incrementally extended per-row caches (1/|v|, bf16 shadow) of the batched path."""
import numpy as np
import pytest

from wax_b200 import CUDAVectorEngine, VectorMetric

from helpers import EngineModel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", [VectorMetric.cosine, VectorMetric.l2])
def test_remove_batch_equals_one_remove_per_id(oracle, metric):
    """Random op sequences with remove_batch (unknown ids, repeated ids, everything, nothing) against the list model:
    row order, ids, raw rows and search results."""
    dims = 16
    rng = np.random.default_rng(11 + metric.value)
    eng, model = CUDAVectorEngine(metric, dims), EngineModel(oracle, metric.value, dims)
    pool = list(range(500, 620))
    for step in range(60):
        op = rng.integers(0, 10)
        if op < 6:
            n = int(rng.integers(1, 30))
            ids = [int(x) for x in rng.choice(pool, n)]
            vs = rng.standard_normal((n, dims)).astype(np.float32)
            eng.add_batch(ids, vs); model.add_batch(ids, list(vs))
        else:
            n = int(rng.integers(1, 25))
            ids = [int(x) for x in rng.choice(pool + [7, 8, 9], n)]        # 7..9 never exist; duplicates allowed
            present = len({i for i in ids if i in model.ids})
            assert eng.remove_batch(ids) == present
            for i in ids:
                model.remove(i)
        assert eng.count == len(model.ids)
        if model.ids:
            assert np.array_equal(eng.read_rows(0, eng.count), model.corpus())
            q = rng.standard_normal(dims).astype(np.float32)
            got, exp = eng.search(q, 20), model.search(q, 20)
            assert [g[0] for g in got] == [e[0] for e in exp]
            assert np.array_equal(np.float32([g[1] for g in got]), np.float32([e[1] for e in exp]))
    assert eng.remove_batch(list(model.ids) + [1, 2]) == len(model.ids)     # everything
    assert eng.count == 0 and eng.search(np.ones(dims, np.float32), 3) == []
    assert eng.remove_batch([1, 2, 3]) == 0                                  # empty engine: no-op (:425)


def test_remove_batch_compacts_a_large_matrix_through_the_bounce_buffer(oracle):
    """Enough rows that the compaction runs over several 256 MB slabs, with gaps at the start, inside and at the end,
    on a synthetic-filled engine (implicit ids): rows and ids after == numpy delete; blob == the oracle's encoding."""
    dims, n = 384, 400_000                                   # 614 MB: three slabs
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    eng.fill_synthetic(81, n, id_base=1000)
    rng = np.random.default_rng(2)
    gone = np.unique(np.concatenate([[0, 1, n - 1, n - 2, 200_000], rng.integers(0, n, 5000)]))
    before = eng.read_rows(0, n)
    assert eng.remove_batch([int(1000 + r) for r in gone] + [5]) == gone.size
    keep = np.setdiff1d(np.arange(n), gone)
    assert eng.count == keep.size
    after = eng.read_rows(0, eng.count)
    assert np.array_equal(after, before[keep])
    q = oracle.synth_row(82, 0, dims, True)
    r, _, s = oracle.search(oracle.COSINE, after, q, 10, mode=oracle.ACC_F32_TREE, threads=8)
    got = eng.search(q, 10)
    assert [g[0] for g in got] == [int(1000 + keep[int(i)]) for i in r]
    assert np.array_equal(np.float32([g[1] for g in got]).view(np.uint32), s.view(np.uint32))
    blob = eng.serialize()
    assert blob == oracle.mv2v_encode(0, after, 1000 + keep)
    # a contiguous tail move (single id near the front) takes the plain-copy branch
    eng.remove(int(1000 + keep[3]))
    assert np.array_equal(eng.read_rows(0, 10), np.delete(after[:11], 3, axis=0))


def test_chunked_transfers_round_trip_bit_exactly(oracle):
    """add_batch / serialize / deserialize of more than one 64 MB staging chunk, from pageable and from pinned host
    memory; the upsert (scatter) path through the persistent device staging."""
    import torch
    dims, n = 384, 150_000                                   # 230 MB: four chunks
    rows = oracle.synth_rows(83, 0, n, dims)
    ids = np.arange(10, 10 + n, dtype=np.uint64)
    eng = CUDAVectorEngine(VectorMetric.dot, dims)
    eng.add_batch(ids[:100_000], rows[:100_000])                           # pageable source
    pinned = torch.from_numpy(rows[100_000:]).pin_memory()
    eng.add_batch(ids[100_000:], pinned.numpy())                           # already pinned: handed to the DMA engine as is
    assert eng.counter("ingest_h2d_bytes") == n * dims * 4
    assert np.array_equal(eng.read_rows(0, n), rows)
    blob = eng.serialize()
    assert eng.counter("ingest_d2h_bytes") == n * dims * 4
    assert blob == oracle.mv2v_encode(1, rows, ids)
    other = CUDAVectorEngine(VectorMetric.dot, dims)
    other.deserialize(blob)
    assert other.count == n and np.array_equal(other.read_rows(0, n), rows)
    # upsert 40 000 existing rows (reversed order, with an in-batch duplicate) + 10 new ones in one call
    up_ids = np.concatenate([ids[:40_000][::-1], ids[:1], np.arange(5_000_000, 5_000_010, dtype=np.uint64)])
    up_rows = oracle.synth_rows(84, 0, up_ids.size, dims)
    other.add_batch(up_ids, up_rows)
    expect = rows.copy()
    expect[:40_000] = up_rows[:40_000][::-1]
    expect[0] = up_rows[40_000]                                            # the later duplicate wins
    assert other.count == n + 10
    got = other.read_rows(0, n + 10)
    assert np.array_equal(got[:n], expect) and np.array_equal(got[n:], up_rows[40_001:])


def test_appends_extend_the_cached_norms_and_shadow_instead_of_rebuilding(oracle):
    """The batched path's per-row caches follow the corpus: an append computes only the new rows (counters), an
    overwrite or a remove resets them -- and the batched results stay identical to the single-query path throughout."""
    dims = 384
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    eng.reserve(55_000)
    rows = oracle.synth_rows(85, 0, 60_000, dims) * np.float32(3.0)
    qs = oracle.synth_rows(86, 0, 40, dims)

    def check():
        got = eng.search_batch(qs, 10)
        eng.set_option("batch_tensor", 0)
        exp = [eng.search(q, 10) for q in qs]
        eng.set_option("batch_tensor", 1)
        assert got == exp
    eng.add_batch(list(range(30_000)), rows[:30_000])
    check()
    assert eng.counter("norms_rows") == 30_000 and eng.counter("shadow_rows") == 30_000
    eng.add_batch(list(range(30_000, 50_000)), rows[30_000:50_000])        # pure append
    assert eng.counter("norms_rows") == 30_000 and eng.counter("shadow_rows") == 30_000   # prefix kept, tail pending
    check()
    assert eng.counter("norms_rows") == 50_000 and eng.counter("shadow_rows") == 50_000
    eng.add(5, rows[59_999] * np.float32(100.0))                           # overwrite: everything is recomputed
    assert eng.counter("norms_rows") == 0
    check()
    eng.remove_batch([40_000, 45_000])                                     # rows below 40 000 did not move
    assert eng.counter("norms_rows") == 40_000 and eng.counter("shadow_rows") == 40_000
    check()
    assert eng.counter("norms_rows") == 49_998
    eng.add_batch(list(range(50_000, 60_000)), rows[50_000:])              # append past the reserved capacity: grows
    check()
    assert eng.count == 59_998 and eng.counter("shadow_rows") == 59_998


def test_sorted_id_fast_path_and_the_switch_to_the_hash_table(oracle):
    """Frame ids normally arrive in increasing order: appends then touch no hash table and lookups are binary searches in
    the id array (order-preserving removes and in-place upserts keep it sorted).  The first out-of-order id switches the
    engine to the hash table.  The list model must agree before, across and after the switch."""
    dims = 24
    rng = np.random.default_rng(21)
    eng, model = CUDAVectorEngine(VectorMetric.cosine, dims), EngineModel(oracle, 0, dims)

    def both(fn_name, *args):
        getattr(eng, fn_name)(*args)
        if fn_name == "remove_batch":
            for i in args[0]:
                model.remove(i)
        else:
            getattr(model, fn_name)(*args)

    def check():
        assert eng.count == len(model.ids)
        assert np.array_equal(eng.read_rows(0, eng.count), model.corpus())
        q = rng.standard_normal(dims).astype(np.float32)
        assert [g[0] for g in eng.search(q, 25)] == [m[0] for m in model.search(q, 25)]
        allow = [int(x) for x in rng.choice(model.ids, 12, replace=False)]
        got = eng.search_filtered(q, 5, allow=allow)          # the filtered search resolves ids the same way
        assert set(g[0] for g in got) <= set(allow) and len(got) == 5
    vecs = lambda n: rng.standard_normal((n, dims)).astype(np.float32)   # noqa: E731
    both("add_batch", list(range(1000, 1400)), list(vecs(400)))           # increasing: the fast path
    both("add_batch", list(range(2000, 2100)), list(vecs(100)))
    check()
    both("remove_batch", [1000, 1399, 2050, 1200, 777])
    both("add_batch", [1100, 2099, 3000, 3001], list(vecs(4)))            # upserts + larger ids: still sorted
    both("add", 1100, vecs(1)[0])
    check()
    both("add_batch", [50, 3002, 1300, 50], list(vecs(4)))                # 50 is out of order: hash table from here on
    check()
    both("remove_batch", [50, 3002, 1001])
    both("add_batch", list(range(10, 40)) + [3001], list(vecs(31)))
    check()
    blob = eng.serialize()
    other = CUDAVectorEngine(VectorMetric.cosine, dims)
    other.deserialize(blob)                                               # unsorted ids in the blob: detected on load
    other.add(1300, model.rows[model.ids.index(1300)])
    assert other.count == eng.count and other.search(model.rows[5], 3) == eng.search(model.rows[5], 3)
