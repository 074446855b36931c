"""GPU: the two tails of the fused scan -- pairwise bitonic merges of sorted lists vs. exact radix SELECTION over the keys
(option `tail_select`) -- must produce the same bits in every regime: all k of the fused range, ties, fewer real rows
than k, non-finite rows, tiny and ragged corpora, every metric, and under the sharded exchange."""
import numpy as np
import pytest

from wax_b200 import CUDAVectorEngine, VectorMetric

from test_gpu_sharded import Group

pytestmark = pytest.mark.gpu


def _both(eng, q, k):
    out = []
    for sel in (0, 1):
        eng.set_option("tail_select", sel)
        out.append(eng.search(q, k))
    assert out[0] == out[1], (k, out[0][:3], out[1][:3])
    return out[1]


@pytest.mark.parametrize("metric", list(VectorMetric))
@pytest.mark.parametrize("dims,n", [(384, 50_001), (384, 7), (128, 3_000), (1024, 9_000), (640, 20_000), (768, 300)])
def test_selection_tail_equals_merge_tail(oracle, metric, dims, n):
    eng = CUDAVectorEngine(metric, dims)
    eng.fill_synthetic(500 + dims, n)
    qs = oracle.synth_rows(501, 0, 3, dims) * np.float32(1.3)
    for q in qs:
        for k in (1, 2, 10, 31, 32, 33, 72, 100, 128):
            _both(eng, q, k)
    eng.set_option("tail_select", 1)
    r, _, s = oracle.search_synth(metric.value, 500 + dims, 0, n, dims, True, qs[0], 72, mode=oracle.ACC_F32_TREE, threads=8)
    got = eng.search(qs[0], 72)
    assert [g[0] for g in got] == r.tolist()
    assert np.array_equal(np.float32([g[1] for g in got]).view(np.uint32), s.view(np.uint32))


def test_selection_tail_with_ties_nans_and_short_corpora(oracle):
    dims = 256
    base = oracle.synth_rows(510, 0, 64, dims)
    corpus = np.ascontiguousarray(base[np.arange(5000) % 64])       # every distance 78 times: ties everywhere
    corpus[17] = np.nan
    corpus[18, 3] = np.inf
    corpus[19] = 0.0
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    eng.add_batch(list(range(100, 5100)), corpus)
    for k in (10, 72, 128):
        got = _both(eng, base[5], k)
        assert [g[0] for g in got[:5]] == [105 + 64 * i for i in range(5)]      # equal distances: ascending row
    ones = CUDAVectorEngine(VectorMetric.cosine, 128)                             # the all-ones fixture: one distance
    ones.add_batch(list(range(3000)), np.ones((3000, 128), np.float32))
    assert [g[0] for g in _both(ones, np.ones(128, np.float32), 100)] == list(range(100))
    nan_only = CUDAVectorEngine(VectorMetric.l2, 128)                             # fewer finite rows than k
    rows = np.full((40, 128), np.nan, np.float32)
    rows[5] = 1.0; rows[30] = 2.0
    nan_only.add_batch(list(range(40)), rows)
    assert [g[0] for g in _both(nan_only, np.ones(128, np.float32), 10)] == [5, 30]


def test_selection_tail_under_the_sharded_exchange(oracle):
    dims, total = 384, 40_000
    single = CUDAVectorEngine(VectorMetric.cosine, dims)
    single.fill_synthetic(520, total)
    grp = Group(VectorMetric.cosine, dims, world=4, synth=(520, total))
    try:
        for e in grp.engines:
            e.set_option("tail_select", 1)
        for k in (10, 72, 128):
            q = oracle.synth_row(521, k, dims, True)
            assert grp.search(q, k) == single.search(q, k)
    finally:
        grp.close(); single.close()
