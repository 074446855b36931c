"""CPU: host-side logic of the mirror (no device needed) checked against the oracle."""
import numpy as np
import pytest

import wax_b200
from wax_b200 import sharded


def test_normalize_matches_oracle(oracle):
    for seed in range(5):
        v = oracle.synth_row(40 + seed, 0, 384, False) * np.float32(seed + 0.5)
        assert (wax_b200.normalize_l2(v).view(np.uint32) == oracle.normalize_l2(v).view(np.uint32)).all()
        assert wax_b200.is_normalized_l2(v) == oracle.is_normalized_l2(v)
    assert wax_b200.is_normalized_l2([1.0, 0.0, 0.0]) and not wax_b200.is_normalized_l2([2.0, 0.0, 0.0])
    assert not wax_b200.is_normalized_l2([])
    assert wax_b200.normalize_l2([0.0, 0.0]).tolist() == [0.0, 0.0]


def test_metric_score_matches_oracle(oracle):
    for m in wax_b200.VectorMetric:
        for d in (0.0, 0.25, -3.5, 1e-7, float("nan"), float("inf")):
            assert m.score(d) == oracle.score_from_distance(m.value, d)
        assert (sharded.score_from_distance(m.value, np.array([0.25, np.nan], np.float32)).tolist()
                == [m.score(0.25), 0.0])


def test_shard_ranges_partition_the_corpus():
    for total in (0, 1, 7, 10_000_000, 100_000_000):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_merge_candidates_total_order():
    c = np.zeros(7, sharded.CAND_DTYPE)
    c["distance"] = [0.5, 0.1, 0.5, 0.1, 0.9, 0.0, 0.1]
    c["row"] = [10, 30, 5, 20, 1, 99, 25]
    c["frame_id"] = c["row"] + 1000
    c["valid"] = [1, 1, 1, 1, 1, 0, 1]            # the 0.0 entry is padding
    best = sharded.merge_candidates(c, 4)
    assert best["row"].tolist() == [20, 25, 30, 5]
    assert sharded.merge_candidates(c[5:6], 4).size == 0


def test_engine_argument_errors_are_wax_errors():
    with pytest.raises(wax_b200.InvalidToc):
        wax_b200.CUDAVectorEngine(wax_b200.VectorMetric.cosine, 0)
    with pytest.raises(wax_b200.CapacityExceeded):
        wax_b200.CUDAVectorEngine(wax_b200.VectorMetric.cosine, 1_000_001)


def test_batched_merge_equals_the_per_query_merge():
    """merge_candidates_batch (vectorised, two stable sorts) == merge_candidates per query: ties on distance broken by
    the global row, invalid candidates dropped, fewer than k valid results allowed."""
    from wax_b200 import sharded
    rng = np.random.default_rng(5)
    world, batch, k = 4, 9, 12
    c = np.zeros((world, batch, k), sharded.CAND_DTYPE)
    c["distance"] = rng.integers(0, 6, size=c.shape).astype(np.float32) / 4          # many exact ties
    c["row"] = rng.permutation(c.size).reshape(c.shape)
    c["frame_id"] = c["row"] + 1000
    c["valid"] = rng.integers(0, 3, size=c.shape) > 0
    c["valid"][:, 0, :] = 0                                                           # a query with no valid candidate
    c["valid"][1:, 1, :] = 0; c["valid"][0, 1, 3:] = 0; c["valid"][0, 1, :3] = 1     # a query with 3
    for top in (1, 5, 12, 40):
        best, n = sharded.merge_candidates_batch(c, top)
        assert best.shape == (batch, min(top, world * k))
        for i in range(batch):
            ref = sharded.merge_candidates(c[:, i, :], top)
            assert n[i] == ref.size
            assert np.array_equal(best[i, : ref.size]["row"], ref["row"])
            assert np.array_equal(best[i, : ref.size]["frame_id"], ref["frame_id"])
    assert sharded.merge_candidates_batch(c, 5)[1][0] == 0 and sharded.merge_candidates_batch(c, 5)[1][1] == 3


def test_shard_merge_rule_model():
    """The in-kernel merge of the row-sharded search (wax_b200/csrc/waxvs_shard.cuh) places candidate j of rank r at
    j + sum over lower ranks of upper_bound(distance) + sum over higher ranks of lower_bound(distance).  Model of that
    rule on random sorted lists full of ties and padding: it is exactly the (distance, rank, index) = (distance, global
    row) order, every output position is written once, padding ends up last."""
    import bisect
    rng = np.random.default_rng(0)
    none = 0xFFFFFFFF
    for _ in range(1500):
        world, k = int(rng.integers(1, 9)), int(rng.integers(1, 12))
        lists = []
        for _r in range(world):
            n_valid = int(rng.integers(0, k + 1))
            lists.append(sorted(int(x) for x in rng.integers(0, 6, n_valid)) + [none] * (k - n_valid))
        out = [None] * k
        for r, lst in enumerate(lists):
            for j, key in enumerate(lst):
                pos = j + sum((bisect.bisect_right if r2 < r else bisect.bisect_left)(l2, key)
                              for r2, l2 in enumerate(lists) if r2 != r)
                if pos < k:
                    assert out[pos] is None
                    out[pos] = (key, r, j)
        assert all(o is not None for o in out)
        ref = sorted((key, r, j) for r, l in enumerate(lists) for j, key in enumerate(l) if key != none)[:k]
        assert [o for o in out if o[0] != none] == ref
        assert all(o[0] == none for o in out[len(ref):])
