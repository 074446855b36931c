"""GPU: the fused scan + peer-memory exchange + in-kernel merge of the row-sharded search (wax_vs_shard_*,
wax_b200/csrc/waxvs_shard.cuh), exercised on ONE device: `world` engines in this process act as the ranks (their
mailboxes are reached with plain pointers instead of CUDA IPC handles -- the kernels, flags, acknowledgements and the
merge are the production ones), one thread per rank issues the collective calls.  The multi-process / multi-GPU form
(IPC handles over NVLink) is covered by tests/check_sharded_torchrun.py under torchrun.

Every sharded answer must equal the single-engine answer over the whole corpus: same ids, same score bits, ties
across shards broken by GLOBAL row."""
import threading

import numpy as np
import pytest

from wax_b200 import CUDAVectorEngine, InvalidToc, VectorMetric, sharded

pytestmark = pytest.mark.gpu


class Group:
    """`world` engines holding contiguous shards of one corpus, connected as a shard group inside this process."""

    def __init__(self, metric, dims, corpus=None, ids=None, world=3, synth=None, fused=1):
        self.world = world
        total = len(corpus) if corpus is not None else synth[1]
        self.engines, self.ranges = [], []
        for r in range(world):
            lo, hi = sharded.shard_range(total, world, r)
            eng = CUDAVectorEngine(metric, dims)
            if corpus is not None:
                if hi > lo:
                    eng.add_batch(ids[lo:hi], corpus[lo:hi])
            else:
                eng.fill_synthetic(synth[0], hi - lo, first_row=lo, id_base=lo)
            eng.set_option("shard_fused", fused)
            eng.set_option("shard_timeout_ms", 8000)         # a protocol bug must fail the test quickly, not hold the GPU
            self.engines.append(eng)
            self.ranges.append((lo, hi))
        blobs = [e.shard_open(r, world, self.ranges[r][0]) for r, e in enumerate(self.engines)]
        for e in self.engines:
            e.shard_connect(blobs)

    def collective(self, fn):
        """Run fn(rank, engine) on one thread per rank; returns the per-rank results."""
        out, errors = [None] * self.world, []

        def work(r):
            try:
                out[r] = fn(r, self.engines[r])
            except Exception as exc:  # noqa: BLE001
                errors.append((r, exc))
        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        [t.start() for t in threads]; [t.join() for t in threads]
        assert not errors, errors[:1]
        return out

    def search(self, q, k):
        res = self.collective(lambda r, e: e.shard_search(q, k))
        assert all(x == res[0] for x in res), "ranks disagree on the merged result"
        return res[0]

    def close(self):
        for e in self.engines:
            e.close()


@pytest.mark.parametrize("metric", [VectorMetric.cosine, VectorMetric.dot, VectorMetric.l2])
@pytest.mark.parametrize("world,total,k,fused", [(3, 60_007, 10, 1), (2, 20_000, 72, 1), (8, 100_003, 10, 1),
                                                 (3, 60_007, 10, 0), (5, 33_333, 128, 1), (16, 40_000, 32, 1)])
def test_sharded_search_equals_the_single_engine_answer(oracle, metric, world, total, k, fused):
    dims, seed = 384, 40 + world
    single = CUDAVectorEngine(metric, dims)
    single.fill_synthetic(seed, total)
    grp = Group(metric, dims, world=world, synth=(seed, total), fused=fused)
    try:
        for qi in range(3):
            q = oracle.synth_row(2000 + qi, 0, dims, True) * np.float32(1.0 + qi)      # also un-normalised queries
            assert grp.search(q, k) == single.search(q, k), (qi,)
        q = oracle.synth_row(2000, 0, dims, True)
        r, _, s = oracle.search_synth(metric.value, seed, 0, total, dims, True, q, k, mode=oracle.ACC_F32_TREE, threads=8)
        got = grp.search(q, k)
        assert [g[0] for g in got] == r.tolist()
        assert np.array_equal(np.float32([g[1] for g in got]).view(np.uint32), s.view(np.uint32))
    finally:
        grp.close(); single.close()


def test_ties_across_shards_break_by_global_row(oracle):
    """Period-256 duplicates (the reference benchmark's degenerate generator, MetalVectorEngineBenchmark.swift:33-38):
    every distance occurs in every shard, so the merge must order equal distances by rank, then local row."""
    dims, n = 64, 3000
    base = oracle.synth_rows(50, 0, 256, dims)
    corpus = np.ascontiguousarray(base[np.arange(n) % 256])
    ids = list(range(1000, 1000 + n))
    single = CUDAVectorEngine(VectorMetric.cosine, dims)
    single.add_batch(ids, corpus)
    grp = Group(VectorMetric.cosine, dims, corpus=corpus, ids=ids, world=4)
    try:
        for k in (10, 40, 100):
            got = grp.search(base[7], k)
            assert got == single.search(base[7], k)
            assert [g[0] for g in got[:11]] == [1000 + 7 + 256 * i for i in range(11)][:len(got[:11])]
    finally:
        grp.close(); single.close()


def test_short_and_empty_shards_pad_correctly(oracle):
    """Fewer rows than k in a shard (padding entries), an EMPTY shard (no scan: the stand-alone exchange kernel), fewer
    rows than k in the whole corpus (fewer results)."""
    dims = 128
    corpus = oracle.synth_rows(51, 0, 7, dims)
    ids = [5, 6, 7, 8, 9, 10, 11]
    single = CUDAVectorEngine(VectorMetric.cosine, dims)
    single.add_batch(ids, corpus)
    q = oracle.synth_row(52, 0, dims, True)
    for world in (2, 3, 8):            # 8 ranks over 7 rows: one rank holds nothing
        grp = Group(VectorMetric.cosine, dims, corpus=corpus, ids=ids, world=world)
        try:
            assert grp.search(q, 10) == single.search(q, 10)     # 7 results
            assert grp.search(q, 3) == single.search(q, 3)
        finally:
            grp.close()
    single.close()


def test_many_queries_reuse_the_mailbox_slots(oracle):
    """Many more collective searches than the mailbox has slots (8): slot reuse is guarded by the acknowledgements,
    results stay exact.  (Each rank waits for its result before the next launch: several ranks share ONE device here, and
    queuing launches back to back on streams that may share a hardware queue could order a rank's first kernel behind
    another rank's blocked second one.  The back-to-back form runs under torchrun, one device per rank:
    tests/check_sharded_torchrun.py.)"""
    dims, total = 384, 50_000
    single = CUDAVectorEngine(VectorMetric.cosine, dims)
    single.fill_synthetic(60, total)
    grp = Group(VectorMetric.cosine, dims, world=3, synth=(60, total))
    try:
        qs = oracle.synth_rows(61, 0, 40, dims)
        expect = [single.search(q, 10) for q in qs]
        res = grp.collective(lambda r, e: [e.shard_search(q, 10) for q in qs])
        for r in range(3):
            assert res[r] == expect, r
    finally:
        grp.close(); single.close()


def test_absent_peer_times_out_instead_of_hanging(oracle):
    dims = 128
    grp = Group(VectorMetric.cosine, dims, world=2, synth=(70, 5000))
    try:
        grp.engines[0].set_option("shard_timeout_ms", 300)
        q = oracle.synth_row(71, 0, dims, True)
        with pytest.raises(InvalidToc, match="timed out"):
            grp.engines[0].shard_search(q, 10)                  # rank 1 never calls
    finally:
        grp.close()


def test_sharded_engine_uses_the_fused_transport_on_one_rank(oracle):
    """ShardedVectorEngine without a process group (world 1): the p2p-fused transport end to end (push to self, merge)."""
    eng = sharded.ShardedVectorEngine(VectorMetric.cosine, 384, total_rows=100_000)
    eng.fill_synthetic(11)
    assert eng.transport == "p2p-fused"
    q = oracle.synth_row(1005, 0, 384, True)
    rows, _, s = oracle.search_synth(oracle.COSINE, 11, 0, 100_000, 384, True, q, 10, mode=oracle.ACC_F32_TREE, threads=8)
    got = eng.search(q, 10)
    assert [g[0] for g in got] == rows.tolist() and np.array_equal(np.float32([g[1] for g in got]), s)
    assert eng.search(q, 200) == eng.engine.search(q, 200)     # k > 128: the all-gather transport
    ms, launches = eng.time_search(10, 10, warmup=3, n_queries=4)
    assert ms > 0 and launches == 10                              # one launch per query: the exchange is inside the scan


@pytest.mark.parametrize("metric", [VectorMetric.cosine, VectorMetric.l2])
def test_sharded_filtered_search_equals_the_single_engine_filtered_search(oracle, metric):
    """wax_vs_shard_search_filtered: every rank passes the same ids, resolves the ones its shard holds, applies the row
    filter inside its fused scan; the in-kernel exchange merges.  Equal to the single-engine filtered search (itself
    oracle-checked in test_gpu_filtered.py) for allow- and deny-lists, including lists that leave whole shards empty."""
    dims, total, world = 256, 50_011, 4
    corpus = oracle.synth_rows(3100, 0, total, dims, normalize=True)
    ids = np.arange(total, dtype=np.uint64) * 5 + 9
    single = CUDAVectorEngine(metric, dims)
    single.add_batch(ids, corpus)
    grp = Group(metric, dims, corpus=corpus, ids=ids, world=world)
    rng = np.random.default_rng(5)
    try:
        def both(q, k, **kw):
            res = grp.collective(lambda r, e: e.shard_search_filtered(q, k, **kw))
            assert all(x == res[0] for x in res), "ranks disagree on the merged result"
            return res[0], single.search_filtered(q, k, **kw)
        for qi in range(3):
            q = oracle.synth_row(3101 + qi, 0, dims, True)
            allow = ids[rng.choice(total, 20_000, replace=False)]
            got, want = both(q, 10, allow=allow)
            assert got == want
            got, want = both(q, 72, deny=[i for i, _ in single.search(q, 40)])
            assert got == want
            first_shard_only = ids[: grp.ranges[0][1]][rng.choice(grp.ranges[0][1], 300, replace=False)]
            got, want = both(q, 10, allow=first_shard_only)             # every other shard contributes nothing
            assert got == want and len(got) == 10
            got, want = both(q, 10, allow=ids[[3, total - 1]])          # fewer allowed rows than k
            assert got == want and len(got) == 2
            assert both(q, 10, allow=[1, 2, 3])[0] == []                # unknown ids only
        assert grp.search(q, 10) == single.search(q, 10)                # the unfiltered path still works afterwards
    finally:
        grp.close()
        single.close()


def test_device_merge_of_gathered_lists_equals_the_host_merge():
    """wax_vs_merge_candidates_device (the sharded search_batch's merge kernel) against the numpy merge it replaces:
    random sorted per-rank lists with exact distance ties inside and across ranks, padding (valid = 0) at the tails,
    ranks owning ascending row ranges -- identical records in identical order, every k_out."""
    import ctypes as C
    import torch
    from wax_b200 import _lib as L
    rng = np.random.default_rng(77)
    eng = CUDAVectorEngine(VectorMetric.cosine, 8)
    try:
        for world, b, k in [(8, 300, 10), (2, 17, 128), (16, 5, 72), (1, 9, 10), (3, 40, 1)]:
            cands = np.zeros((world, b, k), sharded.CAND_DTYPE)
            for r in range(world):
                for q in range(b):
                    n_valid = int(rng.integers(0, k + 1)) if rng.random() < 0.3 else k
                    d = np.round(rng.random(n_valid).astype(np.float32), 2)               # coarse: many exact ties
                    rows = rng.choice(1000, n_valid, replace=False).astype(np.uint64) + np.uint64(r * 1000)
                    order = np.lexsort((rows, d))
                    cands[r, q, :n_valid]["distance"] = d[order]
                    cands[r, q, :n_valid]["row"] = rows[order]
                    cands[r, q, :n_valid]["frame_id"] = rows[order] * np.uint64(3)
                    cands[r, q, :n_valid]["valid"] = 1
            dev = torch.from_numpy(cands.view(np.uint8).reshape(-1).copy()).cuda()
            for k_out in sorted({k, max(1, k // 2), 1}):
                out = torch.zeros(b * k_out * 24, dtype=torch.uint8, device="cuda")
                rc = L.lib().wax_vs_merge_candidates_device(eng.handle, C.c_void_p(dev.data_ptr()), world, b, k, k_out,
                                                            C.c_void_p(out.data_ptr()), None)
                assert rc == 0, L.last_error()
                torch.cuda.synchronize()
                got = out.cpu().numpy().view(sharded.CAND_DTYPE).reshape(b, k_out)
                want, n_valid = sharded.merge_candidates_batch(cands, k_out)
                for q in range(b):
                    m = int(n_valid[q])
                    assert np.array_equal(got[q, :m], want[q, :m]), (world, b, k, k_out, q)
                    assert not got[q, m:]["valid"].any()
    finally:
        eng.close()
