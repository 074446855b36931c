"""GPU: filtered search (allow-/deny-list pushed below the top-k) against the oracle run on the
allowed subset of the corpus.  Same bars as everywhere: identical ids, identical score bits."""
import numpy as np
import pytest

from wax_b200 import CUDAVectorEngine, VectorMetric

pytestmark = pytest.mark.gpu


def _expect(oracle, metric, corpus, ids, allowed_rows, q, k):
    rows = np.array(sorted(allowed_rows), np.int64)
    if rows.size == 0:
        return []
    r, d, s = oracle.search(metric.value, corpus[rows], q, k, mode=oracle.ACC_F32_TREE, threads=4)
    return [(int(ids[rows[int(i)]]), float(sc)) for i, sc in zip(r, s)]


@pytest.mark.parametrize("metric", list(VectorMetric))
def test_allow_and_deny_lists_match_the_oracle_on_the_subset(oracle, metric):
    n, dims = 60_000, 384
    rng = np.random.default_rng(11 + metric.value)
    corpus = oracle.synth_rows(1300, 0, n, dims, normalize=(metric is not VectorMetric.dot))
    ids = (np.arange(n, dtype=np.uint64) * 7 + 1000)                     # frameIds distinct from rows
    eng = CUDAVectorEngine(metric, dims)
    eng.add_batch(ids, corpus)
    q = oracle.synth_row(1301, 0, dims, True)
    # small allow-list -> gather path; with duplicates and unknown ids mixed in
    allow_rows = rng.choice(n, 500, replace=False)
    allow = np.concatenate([ids[allow_rows], ids[allow_rows[:20]], np.array([5, 6, 2**60], np.uint64)])
    for k in (1, 10, 72, 200, 10_000):
        assert eng.search_filtered(q, k, allow=allow) == _expect(oracle, metric, corpus, ids, allow_rows, q, k)
    # large allow-list -> bitset consulted inside the fused scan (k <= 128) and in the emit path (k > 128)
    allow_rows = rng.choice(n, 30_000, replace=False)
    for k in (10, 72, 200):
        assert eng.search_filtered(q, k, allow=ids[allow_rows]) == _expect(oracle, metric, corpus, ids, allow_rows, q, k)
    # deny-list: the unfiltered best hits are exactly the ones removed
    best = [i for i, _ in eng.search(q, 25)]
    deny_rows = set(((np.array(best, np.uint64) - 1000) // 7).astype(np.int64).tolist())
    got = eng.search_filtered(q, 10, deny=best + [3, 4])
    assert got == _expect(oracle, metric, corpus, ids, set(range(n)) - deny_rows, q, 10)
    assert not set(i for i, _ in got) & set(best)
    assert eng.search_filtered(q, 10, deny=[]) == eng.search(q, 10)


def test_filter_edge_cases(oracle):
    dims = 128
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    assert eng.search_filtered(np.ones(dims, np.float32), 5, allow=[1, 2]) == []          # empty engine
    eng.fill_synthetic(1400, 5000, id_base=100)                                           # implicit ids 100..5099
    corpus = oracle.synth_rows(1400, 0, 5000, dims)
    ids = np.arange(5000, dtype=np.uint64) + 100
    q = oracle.synth_row(1401, 0, dims, True)
    assert eng.search_filtered(q, 10, allow=[]) == []                                     # nothing allowed
    assert eng.search_filtered(q, 10, allow=[7, 8, 99, 5100]) == []                       # only unknown ids
    assert eng.search_filtered(q, 10, allow=[100, 5099]) == _expect(oracle, VectorMetric.cosine, corpus, ids, [0, 4999], q, 10)
    assert eng.search_filtered(q, 10, deny=ids) == []                                     # everything denied
    got = eng.search_filtered(q, 10, deny=ids[:-3])                                       # three rows left
    assert got == _expect(oracle, VectorMetric.cosine, corpus, ids, [4997, 4998, 4999], q, 10) and len(got) == 3
    with pytest.raises(Exception):
        eng.search_filtered(q[:5], 10, allow=[100])
    with pytest.raises(ValueError):
        eng.search_filtered(q, 10)
    # after a mutation the ids are explicit: the filter follows the moved rows
    eng.remove(101)
    assert eng.search_filtered(q, 5, allow=[100, 101, 102]) == _expect(oracle, VectorMetric.cosine, corpus, ids, [0, 2], q, 5)


def test_filter_replaces_the_reference_overfetch(oracle):
    """UnifiedSearchTests.swift:133-158 (filtersAllowResultsBeyondTopK): with an allow-list of {id2, id3} and topK 2
    the reference has to ask the engine for more than topK and filter afterwards; here the engine returns exactly
    the two allowed frames."""
    eng = CUDAVectorEngine(VectorMetric.cosine, 4)
    eng.add_batch([0, 1, 2, 3], [[1, 0, 0, 0], [0.9, 0.1, 0, 0], [0.5, 0.5, 0, 0], [0, 1, 0, 0]])
    assert [i for i, _ in eng.search([1, 0, 0, 0], 2)] == [0, 1]
    assert [i for i, _ in eng.search_filtered([1, 0, 0, 0], 2, allow=[2, 3])] == [2, 3]


@pytest.mark.parametrize("metric", list(VectorMetric))
def test_batched_filtered_search_equals_the_per_query_filtered_search(oracle, metric):
    """wax_vs_search_batch_filtered: one filter, a batch of queries, one pass.  Small allow-lists take the batched gather,
    larger ones the tensor-core levels with the row filter below the top-k (cosine / dot) or the fused scan (l2); every
    answer equals the single-query filtered search (itself checked against the oracle above), and the first query is
    checked against the oracle directly."""
    n, dims, b = 80_000, 384, 130
    rng = np.random.default_rng(21 + metric.value)
    corpus = oracle.synth_rows(1500, 0, n, dims, normalize=(metric is not VectorMetric.dot))
    ids = (np.arange(n, dtype=np.uint64) * 3 + 77)
    eng = CUDAVectorEngine(metric, dims)
    eng.add_batch(ids, corpus)
    qs = oracle.synth_rows(1501, 0, b, dims, normalize=True)
    small = rng.choice(n, 700, replace=False)
    large = rng.choice(n, 40_000, replace=False)
    deny = rng.choice(n, 50_000, replace=False)
    cases = [("allow", small, set(small.tolist())), ("allow", large, set(large.tolist())),
             ("deny", deny, set(range(n)) - set(deny.tolist()))]
    for kind, rows, allowed in cases:
        for k in (10, 72):
            kw = {kind: ids[rows]}
            t0, f0 = eng.batch_stats()
            got = eng.search_batch_filtered(qs, k, **kw)
            t1, f1 = eng.batch_stats()
            assert len(got) == b
            assert got[0] == _expect(oracle, metric, corpus, ids, allowed, qs[0], k)
            for qi in range(0, b, 9):
                assert got[qi] == eng.search_filtered(qs[qi], k, **kw), (kind, len(rows), k, qi)
            if metric is not VectorMetric.l2 and len(rows) > 16_384:
                assert (t1 - t0) + (f1 - f0) == b, "a large filtered batch must take the tensor-core levels"
                assert f1 - f0 <= 3, f"{f1 - f0} of {b} filtered queries fell back to the exact scan"
    # the best unfiltered hits of every query denied at once: none of them may come back
    best = sorted({i for q in qs[:8] for i, _ in eng.search(q, 10)})
    got = eng.search_batch_filtered(qs[:8], 10, deny=best)
    assert all(not (set(i for i, _ in hits) & set(best)) for hits in got)
    assert eng.search_batch_filtered(qs[:8], 10, deny=[]) == eng.search_batch(qs[:8], 10)
    assert eng.search_batch_filtered(qs[:8], 10, allow=[]) == [[] for _ in range(8)]
    assert eng.search_batch_filtered(qs[:8], 10, allow=ids[:3]) == [eng.search_filtered(q, 10, allow=ids[:3]) for q in qs[:8]]
    assert eng.search_batch_filtered([], 10, allow=ids[:3]) == []
