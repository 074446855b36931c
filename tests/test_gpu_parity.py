"""GPU parity tests proper: CUDA path (through the C-ABI) vs the CPU oracle on the same seeded inputs.

Bars (north_star): identical top-k id ordering; cosine scores within 1e-4 fp32.  Two comparisons per case:
  * oracle ACC_F32_TREE mirrors the kernels' accumulation order -> ids AND score bits must be IDENTICAL;
  * oracle ACC_F64 / ACC_F32_SEQ (reference order) -> scores within TOL, ids identical outside fp64 near-ties.
"""
import hashlib

import numpy as np
import pytest

from wax_b200 import CUDAVectorEngine, VectorMetric

from helpers import assert_tie_aware_order, load_json

pytestmark = pytest.mark.gpu
TOL = 1e-4   # north_star: "cosine scores within 1e-4 fp32"


def _engine_from(metric, corpus, ids=None):
    eng = CUDAVectorEngine(metric, corpus.shape[1])
    eng.add_batch(list(range(corpus.shape[0])) if ids is None else ids, corpus)
    return eng


def _check(oracle, eng, metric, corpus, q, k, rel_scale=1.0):
    got = eng.search(q, k)
    rows, d, s = oracle.search(metric.value, corpus, q, k, mode=oracle.ACC_F32_TREE, threads=4)
    assert [g[0] for g in got] == rows.tolist(), "ids differ from the bit-exact (tree-order) oracle"
    assert np.array_equal(np.float32([g[1] for g in got]).view(np.uint32), s.view(np.uint32)), "score bits differ"
    k_eff = len(got)
    r64, _, s64 = oracle.search(metric.value, corpus, q, k_eff + 1, mode=oracle.ACC_F64, threads=4)
    tol = TOL * rel_scale
    assert np.max(np.abs(np.float64([g[1] for g in got]) - s64[:k_eff].astype(np.float64)), initial=0.0) <= tol
    assert_tie_aware_order([g[0] for g in got], r64[:k_eff].tolist(), s64.astype(np.float64), 2e-6 * rel_scale)
    rseq, _, sseq = oracle.search(metric.value, corpus, q, k, mode=oracle.ACC_F32_SEQ, threads=4)
    assert np.max(np.abs(np.float32([g[1] for g in got]) - sseq), initial=0.0) <= tol
    return got


def test_golden_c1_10k_384(oracle):
    """BASELINE configs[0]: 10K x 384, committed golden vectors (ids + score bits, k = 72 and top-10)."""
    g = load_json("c1_10k_384.json")
    qs = {"unit": oracle.synth_row(g["query_seed"], 0, g["dims"], True),
          "raw": oracle.synth_row(g["query_seed"], 1, g["dims"], False) * np.float32(3.5)}
    for metric in VectorMetric:
        eng = CUDAVectorEngine(metric, g["dims"])
        eng.fill_synthetic(g["seed"], g["rows"])
        corpus = eng.read_rows(0, g["rows"])
        assert hashlib.sha256(corpus.tobytes()).hexdigest() == g["corpus_sha256"]   # device generator == oracle's
        for qname, q in qs.items():
            exp = g["queries"][qname]["metrics"][metric.name]
            for k in (72, 10):
                got = eng.search(q, k)
                assert [i for i, _ in got] == exp["f32_tree"]["rows"][:k]
                assert [int(x) for x in np.float32([s for _, s in got]).view(np.uint32)] == exp["f32_tree"]["score_bits"][:k]
                f64 = np.array(exp["f64"]["score_bits"], np.uint32).view(np.float32)[:k]
                scale = 1.0 if metric is VectorMetric.cosine else max(1.0, float(np.max(np.abs(f64))))
                assert np.max(np.abs(np.float32([s for _, s in got]) - f64)) <= TOL * scale
                assert [i for i, _ in got][:10] == exp["f64"]["rows"][:10]


@pytest.mark.parametrize("dims", [1, 2, 3, 4, 7, 100, 128, 130, 256, 384, 512, 768, 1024, 1536, 4100])
@pytest.mark.parametrize("metric", list(VectorMetric))
def test_parity_across_dimensions(oracle, dims, metric):
    n = 3001 if dims <= 1024 else 700
    corpus = oracle.synth_rows(100 + dims, 0, n, dims, normalize=(metric is not VectorMetric.dot))
    q = oracle.synth_row(200 + dims, 0, dims, metric is VectorMetric.cosine)
    eng = _engine_from(metric, corpus)
    scale = 1.0 if metric is VectorMetric.cosine else float(dims)
    for k in (1, 10, 32):
        _check(oracle, eng, metric, corpus, q, k, rel_scale=scale)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 1000, 4736, 4737, 50_001])
def test_ragged_row_counts(oracle, n):
    corpus = oracle.synth_rows(300 + n, 0, n, 384)
    q = oracle.synth_row(301, 0, 384, True)
    eng = _engine_from(VectorMetric.cosine, corpus)
    for k in (1, 10, 32, 33, 72):
        got = _check(oracle, eng, VectorMetric.cosine, corpus, q, k)
        assert len(got) == min(k, n)


@pytest.mark.parametrize("k", [33, 64, 72, 100, 256, 1000, 4096, 10_000])
@pytest.mark.parametrize("metric", [VectorMetric.cosine, VectorMetric.dot])
def test_large_k_select_path(oracle, k, metric):
    """k > 32 takes the emit + radix-select path; k = 72 is the production candidateLimit
    (UnifiedSearch.swift:1195-1200), 100 is BASELINE configs[4], 10 000 the API clamp."""
    dims = 384 if k <= 1000 else 128
    corpus = oracle.synth_rows(400, 0, 20_011, dims, normalize=(metric is VectorMetric.cosine))
    q = oracle.synth_row(401, 3, dims, True)
    eng = _engine_from(metric, corpus)
    _check(oracle, eng, metric, corpus, q, k, rel_scale=1.0 if metric is VectorMetric.cosine else float(dims))


def test_unnormalised_query_is_divided_by_its_own_norm(oracle):
    """The fused in-kernel |q|: a scaled query gives the same ids and (to 1e-6) the same cosine scores."""
    corpus = oracle.synth_rows(500, 0, 5000, 384)
    eng = _engine_from(VectorMetric.cosine, corpus)
    q = oracle.synth_row(501, 0, 384, True)
    base = eng.search(q, 10)
    for scale in (12.0, 1.0009, 1e-3, 3e4):
        got = _check(oracle, eng, VectorMetric.cosine, corpus, q * np.float32(scale), 10)
        assert [g[0] for g in got] == [b[0] for b in base]
        assert np.max(np.abs(np.float32([g[1] for g in got]) - np.float32([b[1] for b in base]))) < 1e-6


def test_exact_ties_break_by_row(oracle):
    """Degenerate generators of the reference: all-ones fixture (Fixtures/minilm_baseline_embeddings.json) and
    the period-256 duplicate rows of MetalVectorEngineBenchmark.swift:33-38."""
    ones = np.ones((300, 384), np.float32)
    eng = _engine_from(VectorMetric.cosine, ones, ids=[1000 + i for i in range(300)])
    for k in (5, 32, 40):
        got = eng.search(np.ones(384, np.float32), k)
        assert [g[0] for g in got] == [1000 + i for i in range(k)]
    dims, n = 128, 2000
    rows = np.array([[((i + d) % 256) / 255.0 for d in range(dims)] for i in range(n)], np.float32)
    eng = _engine_from(VectorMetric.cosine, rows)
    q = rows[17].copy()
    for k in (8, 24, 100):
        _check(oracle, eng, VectorMetric.cosine, rows, q, k)
        got = eng.search(q, k)
        assert [g[0] for g in got[:8]] == [17 + 256 * j for j in range(8)]     # exact duplicates, ascending row


def test_zero_and_nonfinite_rows(oracle):
    corpus = oracle.synth_rows(600, 0, 400, 384)
    corpus[5] = 0.0                                  # zero-norm document: cosine distance 1 (score 0)
    corpus[6, 3] = np.nan                            # dropped
    corpus[7, 9] = np.inf                            # dropped (inf/inf -> nan)
    corpus[8] = -corpus[0]
    q = corpus[0].copy()
    for metric in VectorMetric:
        eng = _engine_from(metric, corpus)
        got = _check(oracle, eng, metric, corpus, q, 400)
        ids = [g[0] for g in got]
        assert 6 not in ids and 7 not in ids and len(got) == 398
        if metric is VectorMetric.cosine:
            assert dict(got)[5] == 0.0 and ids[0] == 0 and ids[-1] == 8
    # zero query: every finite non-zero row has distance 1, the zero row distance 0 (USearch rules)
    eng = _engine_from(VectorMetric.cosine, corpus)
    got = eng.search(np.zeros(384, np.float32), 3)
    assert got[0] == (5, 1.0) and [g[0] for g in got[1:]] == [0, 1] and got[1][1] == 0.0


def test_search_batch_equals_single_searches(oracle):
    corpus = oracle.synth_rows(700, 0, 9000, 384)
    eng = _engine_from(VectorMetric.cosine, corpus)
    qs = oracle.synth_rows(701, 0, 5, 384)
    batch = eng.search_batch(qs, 10)
    assert batch == [eng.search(q, 10) for q in qs]
    assert eng.search_batch(qs, 50) == [eng.search(q, 50) for q in qs]


def test_device_generator_is_bit_identical_to_the_oracle(oracle):
    for dims, normalize in ((384, True), (384, False), (7, True), (768, True)):
        eng = CUDAVectorEngine(VectorMetric.cosine, dims)
        eng.fill_synthetic(42, 2000, first_row=12345, id_base=900, normalize=normalize)
        exp = oracle.synth_rows(42, 12345, 2000, dims, normalize=normalize)
        assert np.array_equal(eng.read_rows(0, 2000).view(np.uint32), exp.view(np.uint32))
        got = eng.search(exp[10], 1)
        assert got[0][0] == 910                       # frameId = id_base + row
    eng.remove(905)                                   # first mutation materialises the implicit ids
    assert eng.count == 1999 and eng.search(exp[10], 1)[0][0] == 910 and 905 not in [i for i, _ in eng.search(exp[5], 50)]


def test_kernel_variants_agree(oracle):
    """TMA-staged and direct-load kernels, every ring geometry: identical bits."""
    corpus = oracle.synth_rows(800, 0, 40_000, 384)
    q = oracle.synth_row(801, 0, 384, True)
    eng = _engine_from(VectorMetric.cosine, corpus)
    ref = eng.search(q, 32)
    for opts in ({"variant": 2}, {"variant": 1, "rows_per_step": 8, "stages": 2}, {"variant": 1, "rows_per_step": 4, "stages": 3, "warps": 10},
                 {"variant": 1, "rows_per_step": 4, "stages": 2, "warps": 16}, {"variant": 1, "chunk_steps": 0}, {"variant": 1, "chunk_steps": 3}, {"variant": 1, "warps": 4, "grid": 7}, {"variant": 1, "l2_hint": 1}):
        for key in ("variant", "rows_per_step", "stages", "warps", "grid", "l2_hint"):
            eng.set_option(key, opts.get(key, 0))
        eng.set_option("chunk_steps", opts.get("chunk_steps", 8))
        assert eng.search(q, 32) == ref, opts
        assert eng.search(q, 100)[:32] == ref, opts


@pytest.mark.parametrize("metric", list(VectorMetric))
def test_fused_k128_lists_agree_with_the_select_path(oracle, metric):
    """33 <= k <= 128 stays in the fused launch (4 list slots per lane); forcing the emit + radix-select path must
    give the same bits.  k = 72 is the production candidateLimit (UnifiedSearch.swift:1195-1200)."""
    corpus = oracle.synth_rows(1200, 0, 30_011, 384, normalize=(metric is not VectorMetric.dot))
    corpus[100:164] = corpus[99]                       # 65 exact duplicates straddling list slots
    eng = _engine_from(metric, corpus)
    q = corpus[99] + oracle.synth_row(1201, 0, 384, True) * np.float32(0.05)
    for k in (33, 64, 65, 72, 96, 97, 128):
        eng.set_option("fused_k_max", 128)
        fused = _check(oracle, eng, metric, corpus, q, k, rel_scale=1.0 if metric is VectorMetric.cosine else 384.0)
        eng.set_option("fused_k_max", 32)
        assert eng.search(q, k) == fused
    eng.set_option("fused_k_max", 128)
    eng.set_option("variant", 2)                       # generic (direct-load) kernel, same lists
    assert eng.search(q, 72) == fused[:72] or eng.search(q, 128) == fused
    assert eng.search(q, 128) == fused
