"""CPU tests of the oracle itself: pinned against the reference's known-answer tests (reference_kats.json),
against the committed golden vectors, and against the semantics checklist of SURVEY.md section 8c."""
import hashlib
import math

import numpy as np
import pytest

from helpers import EngineModel, load_json

METRICS = {"cosine": 0, "dot": 1, "l2": 2}


def _run_kat_steps(case, make_engine, search_fn, roundtrip_fn):
    eng = make_engine()
    remembered = {}
    for step in case["steps"]:
        op = step[0]
        if op == "add":
            eng.add(step[1], step[2])
        elif op == "add_batch":
            eng.add_batch(step[1], step[2])
        elif op == "remove":
            eng.remove(step[1])
        elif op == "roundtrip":
            eng = roundtrip_fn(eng)
        elif op == "check":
            c = step[1]
            hits = search_fn(eng, c["query"], c["top_k"], case)
            ids = [h[0] for h in hits]
            if c.get("non_empty"):
                assert hits
            for i in c.get("contains", []):
                assert i in ids
            for i in c.get("not_contains", []):
                assert i not in ids
            if "first" in c:
                assert ids and ids[0] == c["first"]
            if "remember_first_score" in c:
                remembered[c["remember_first_score"]] = hits[0][1]
            if "first_score_within" in c:
                name, tol = c["first_score_within"]
                assert abs(hits[0][1] - remembered[name]) < tol
    return eng


@pytest.mark.parametrize("case", [c for c in load_json("reference_kats.json")["cases"] if c["steps"]],
                         ids=lambda c: c["name"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_oracle_reference_kats(oracle, case, mode):
    o = oracle

    def search(eng, query, k, case):
        q = np.asarray(query, np.float32)
        if case.get("session") and not o.is_normalized_l2(q):   # VectorSearchSession.swift:70-76
            q = o.normalize_l2(q)
        if case.get("normalize_query"):                          # UnifiedSearchTests.swift:75
            q = o.normalize_l2(q)
        return eng.search(q, k, mode=mode)

    def roundtrip(eng):
        blob = o.mv2v_encode(eng.metric, eng.corpus() if eng.rows else np.zeros((0, eng.dims), np.float32), eng.ids)
        rc, vec, ids = o.mv2v_decode(blob, eng.metric, eng.dims)
        assert rc == 0
        fresh = EngineModel(o, eng.metric, eng.dims)
        fresh.add_batch([int(i) for i in ids], list(vec))
        return fresh

    _run_kat_steps(case, lambda: EngineModel(o, METRICS[case["metric"]], case["dims"]), search, roundtrip)


def test_oracle_search_after_add_kat(oracle):
    """MetalVectorEngineBenchmark.swift:131-172: count == topK; newly added closer vectors appear."""
    o, dims, k = oracle, 128, 5
    m = EngineModel(o, o.COSINE, dims)
    for i in range(100):
        v = np.full(dims, i / 100.0, np.float32); v[0] = 1.0
        m.add(i, v)
    q = np.full(dims, 0.5, np.float32)
    assert len(m.search(q, k)) == k
    for i in range(100, 200):
        v = np.full(dims, i / 200.0, np.float32); v[0] = 0.5
        m.add(i, v)
    r2 = m.search(q, k)
    assert len(r2) == k and any(100 <= i < 200 for i, _ in r2)


def test_oracle_matches_golden(oracle):
    o = oracle
    g = load_json("c1_10k_384.json")
    corpus = o.synth_rows(g["seed"], 0, g["rows"], g["dims"], normalize=True, threads=2)
    assert hashlib.sha256(corpus.tobytes()).hexdigest() == g["corpus_sha256"]
    qs = {"unit": o.synth_row(g["query_seed"], 0, g["dims"], True),
          "raw": o.synth_row(g["query_seed"], 1, g["dims"], False) * np.float32(3.5)}
    for qname, q in qs.items():
        assert hashlib.sha256(q.tobytes()).hexdigest() == g["queries"][qname]["query_sha256"]
        for mname, metric in METRICS.items():
            for modename, mode in (("f32_seq", 0), ("f64", 1), ("f32_tree", 2)):
                exp = g["queries"][qname]["metrics"][mname][modename]
                rows, d, s = o.search(metric, corpus, q, g["k"], mode=mode, threads=3)
                assert [int(r) for r in rows] == exp["rows"]
                assert [int(x) for x in d.view(np.uint32)] == exp["distance_bits"]
                assert [int(x) for x in s.view(np.uint32)] == exp["score_bits"]


def test_oracle_modes_agree_within_tolerance(oracle):
    """fp32 orders differ from fp64 truth by far less than the 1e-4 bar (north_star tolerance)."""
    o = oracle
    g = load_json("c1_10k_384.json")
    for mname in METRICS:
        m = g["queries"]["unit"]["metrics"][mname]
        f64 = np.array(m["f64"]["score_bits"], np.uint32).view(np.float32)
        for mode in ("f32_seq", "f32_tree"):
            s = np.array(m[mode]["score_bits"], np.uint32).view(np.float32)
            assert m[mode]["rows"] == m["f64"]["rows"]
            assert np.max(np.abs(s - f64)) <= 1e-5


def test_oracle_streaming_scan_equals_materialised(oracle):
    o = oracle
    corpus = o.synth_rows(5, 100, 3000, 96, normalize=False, threads=1)
    q = o.synth_row(6, 0, 96, True)
    for metric in (0, 1, 2):
        a = o.search(metric, corpus, q, 40, mode=2, row_base=100)
        b = o.search_synth(metric, 5, 100, 3000, 96, False, q, 40, mode=2, threads=4)
        assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all()


def test_score_from_distance(oracle):
    o = oracle  # VectorMetric.swift:32-43
    assert o.score_from_distance(o.COSINE, 0.25) == pytest.approx(0.75)
    assert o.score_from_distance(o.DOT, 0.25) == -0.25
    assert o.score_from_distance(o.L2, 3.0) == -3.0
    for bad in (float("nan"), float("inf"), float("-inf")):
        for m in (0, 1, 2):
            assert o.score_from_distance(m, bad) == 0.0


def test_clamp_topk(oracle):
    o = oracle  # MetalVectorEngine.swift:842-846
    assert [o.clamp_topk(k) for k in (-5, 0, 1, 72, 10_000, 10_001, 2**40)] == [1, 1, 1, 72, 10_000, 10_000, 10_000]


def test_usearch_zero_norm_rules(oracle):
    o = oracle
    z, v = np.zeros(8, np.float32), np.arange(1, 9, dtype=np.float32)
    for mode in (0, 1, 2):
        assert o.distance(o.COSINE, mode, z, z) == 0.0     # both zero -> 0
        assert o.distance(o.COSINE, mode, z, v) == 1.0     # one zero -> 1
        assert o.distance(o.COSINE, mode, v, z) == 1.0
        assert o.distance(o.COSINE, mode, v, v) == pytest.approx(0.0, abs=1e-6)
        assert o.distance(o.DOT, mode, v, v) == pytest.approx(1.0 - 204.0)
        assert o.distance(o.L2, mode, v, z) == pytest.approx(204.0)
    # documented deviation: the Metal kernel maps 0 < |v| <= 1e-6 to distance 1, USearch does not
    tiny = np.full(8, 1e-8, np.float32)
    assert o.metal_cosine_distance(o.normalize_l2(v), tiny) == 1.0
    assert o.distance(o.COSINE, 1, v, tiny) == pytest.approx(1.0 - float(v.sum()) / (math.sqrt(204.0) * math.sqrt(8.0)), abs=1e-5)


def test_metal_kernel_does_not_divide_by_query_norm(oracle):
    """CosineDistance.metal:233-328 assumes |q| = 1; the CUDA path/USearch divide by the real |q|."""
    o = oracle
    v = o.synth_row(3, 0, 384, True)
    q = o.synth_row(3, 1, 384, True) * np.float32(1.0009)   # inside the isNormalizedL2 1e-3 slack
    assert o.is_normalized_l2(q)
    exact = o.distance(o.COSINE, 1, q, v)
    metal = o.metal_cosine_distance(q, v)
    assert abs((1 - metal) - (1 - exact)) == pytest.approx(abs(1 - exact) * 0.0009, rel=0.05)


def test_nonfinite_rows_are_dropped(oracle):
    o = oracle  # MetalVectorEngine.swift:597
    corpus = np.eye(4, dtype=np.float32)
    corpus[1, 0] = np.nan
    corpus[2, 1] = np.inf
    rows, d, s = o.search(o.COSINE, corpus, [1, 1, 1, 1], 10)
    assert sorted(rows.tolist()) == [0, 3] and np.isfinite(d).all()


def test_total_order_on_exact_ties(oracle):
    o = oracle
    ones = np.ones((8, 384), np.float32)   # Fixtures/minilm_baseline_embeddings.json: 8 identical all-ones rows
    for mode in (0, 1, 2):
        rows, d, s = o.search(o.COSINE, ones, np.ones(384, np.float32), 5, mode=mode, threads=3)
        assert rows.tolist() == [0, 1, 2, 3, 4] and len(set(d.tolist())) == 1


def test_k_larger_than_n_and_empty(oracle):
    o = oracle
    corpus = o.synth_rows(9, 0, 7, 16)
    rows, _, _ = o.search(o.DOT, corpus, corpus[3], 10_000)
    assert len(rows) == 7 and rows[0] == 3
    rows, _, _ = o.search(o.DOT, np.zeros((0, 16), np.float32), corpus[3], 10)
    assert len(rows) == 0


def test_metal_cpu_heap_tie_rule(oracle):
    o = oracle  # MetalVectorEngine.swift:671: later equal distances never displace
    d = np.array([0.5, 0.1, 0.5, 0.1, 0.5, 0.0], np.float32)
    rows, dist = o.metal_cpu_topk(d, 3)
    assert rows.tolist() == [5, 1, 3] and dist.tolist() == pytest.approx([0.0, 0.1, 0.1])
    rows, _ = o.metal_cpu_topk(np.full(6, 0.25, np.float32), 3)
    assert rows.tolist() == [0, 1, 2]


def test_normalize_l2(oracle):
    o = oracle  # VectorMath.swift:15-33,123-127; VectorSearchEngineTests.swift:73-76
    assert o.is_normalized_l2([1.0, 0.0, 0.0]) and not o.is_normalized_l2([2.0, 0.0, 0.0])
    assert not o.is_normalized_l2([])
    n = o.normalize_l2([12.0, 0.0])
    assert n.tolist() == [1.0, 0.0]
    assert o.normalize_l2([0.0, 0.0]).tolist() == [0.0, 0.0]
    v = o.synth_row(1, 2, 384, False)
    assert abs(float(np.linalg.norm(o.normalize_l2(v).astype(np.float64))) - 1.0) < 1e-6


def test_mv2v_layout_and_errors(oracle):
    o = oracle  # MetalVectorEngine.swift:682-815; VectorSerializer.swift:175-251
    vec = np.array([[1.0, -2.0], [0.5, 0.25]], np.float32)
    blob = o.mv2v_encode(o.COSINE, vec, [7, 9])
    assert blob[:4] == b"MV2V" and blob[4:6] == b"\x01\x00" and blob[6] == 2 and blob[7] == 0
    assert blob[8:12] == (2).to_bytes(4, "little") and blob[12:20] == (2).to_bytes(8, "little")
    assert blob[20:28] == (16).to_bytes(8, "little") and blob[28:36] == bytes(8)
    assert blob[36:40] == bytes([0x00, 0x00, 0x80, 0x3F])          # 1.0f LE, as WALEmbeddingCodecTests.swift:27-28
    assert blob[40:44] == bytes([0x00, 0x00, 0x00, 0xC0])          # -2.0f LE (:30-31)
    assert blob[52:60] == (16).to_bytes(8, "little") and len(blob) == 36 + 16 + 8 + 16
    rc, v2, ids = o.mv2v_decode(blob, o.COSINE, 2)
    assert rc == 0 and (v2 == vec).all() and ids.tolist() == [7, 9]

    def corrupt(i, b):
        x = bytearray(blob); x[i] = b; return bytes(x)
    assert o.mv2v_decode(blob[:20], 0, 2)[0] == -1
    assert o.mv2v_decode(corrupt(0, 0x58), 0, 2)[0] == -2
    assert o.mv2v_decode(corrupt(4, 2), 0, 2)[0] == -3
    assert o.mv2v_decode(corrupt(6, 1), 0, 2)[0] == -4     # usearch encoding is not readable here (:743)
    assert o.mv2v_decode(blob, o.DOT, 2)[0] == -5
    assert o.mv2v_decode(blob, o.COSINE, 3)[0] == -6
    assert o.mv2v_decode(corrupt(30, 1), 0, 2)[0] == -7
    assert o.mv2v_decode(corrupt(20, 17), 0, 2)[0] == -8
    assert o.mv2v_decode(corrupt(52, 8), 0, 2)[0] == -10
    assert o.mv2v_decode(blob + b"\x00", 0, 2)[0] == -11
    empty = o.mv2v_encode(o.L2, np.zeros((0, 5), np.float32), [])
    assert len(empty) == 44 and o.mv2v_decode(empty, o.L2, 5)[0] == 0


def test_synth_generator_properties(oracle):
    o = oracle  # RAGBenchmarkSupport.swift:130-156 distribution: uniform[-1,1] then L2-normalised
    raw = o.synth_rows(11, 0, 2000, 64, normalize=False)
    assert -1.0 <= raw.min() and raw.max() <= 1.0 and abs(float(raw.mean())) < 0.01
    assert abs(float(raw.std()) - 1 / math.sqrt(3)) < 0.01
    nrm = np.linalg.norm(o.synth_rows(11, 0, 2000, 64).astype(np.float64), axis=1)
    assert np.max(np.abs(nrm - 1.0)) < 1e-6
    a = o.synth_rows(11, 500, 10, 64)
    b = o.synth_rows(11, 0, 2000, 64)[500:510]
    assert (a.view(np.uint32) == b.view(np.uint32)).all()    # addressable by (seed, row)
