"""GPU: the reference's own vector-engine tests, replayed through the CUDA engine via the C-ABI, plus the
semantics checklist of SURVEY.md section 8c (edge cases the reference code defines but does not test)."""
import threading

import numpy as np
import pytest

import wax_b200
from wax_b200 import CUDAVectorEngine, VectorMetric, VectorSearchSession

from helpers import EngineModel, load_json
from test_oracle import METRICS, _run_kat_steps

pytestmark = pytest.mark.gpu
M = {"cosine": VectorMetric.cosine, "dot": VectorMetric.dot, "l2": VectorMetric.l2}


@pytest.mark.parametrize("case", [c for c in load_json("reference_kats.json")["cases"] if c["steps"]],
                         ids=lambda c: c["name"])
def test_reference_kats_through_cuda_engine(case):
    def search(eng, query, k, case):
        if case.get("session"):
            return VectorSearchSession(eng).search(query, k)          # VectorSearchSession.swift:70-76
        if case.get("normalize_query"):
            query = wax_b200.normalize_l2(query)
        return eng.search(query, k)

    def roundtrip(eng):
        blob = eng.serialize()
        assert blob
        fresh = CUDAVectorEngine(eng.metric, eng.dimensions)
        fresh.deserialize(blob)
        return fresh

    eng = _run_kat_steps(case, lambda: CUDAVectorEngine(M[case["metric"]], case["dims"]), search, roundtrip)
    if case.get("pool_reuse"):       # MetalVectorEnginePoolTests.swift:7-20
        a1, r1 = eng.debug_buffer_pool_stats()
        eng.search([1.0, 0.0], 1)
        a2, r2 = eng.debug_buffer_pool_stats()
        assert a2 == a1 and r2 >= r1 + 1


def test_search_after_add_correctness():
    """MetalVectorEngineBenchmark.swift:131-172."""
    dims, k = 128, 5
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    for i in range(100):
        v = np.full(dims, i / 100.0, np.float32); v[0] = 1.0
        eng.add(i, v)
    q = np.full(dims, 0.5, np.float32)
    assert len(eng.search(q, k)) == k
    for i in range(100, 200):
        v = np.full(dims, i / 200.0, np.float32); v[0] = 0.5
        eng.add(i, v)
    r2 = eng.search(q, k)
    assert len(r2) == k and any(100 <= i < 200 for i, _ in r2)


def test_empty_engine_returns_empty_without_validation():
    eng = CUDAVectorEngine(VectorMetric.cosine, 4)
    assert eng.search([1, 0, 0, 0], 10) == []
    assert eng.search([1, 0], 10) == []          # `guard vectorCount > 0` precedes validate (:448-449)
    eng.remove(5)                                # remove on empty: no-op (:425)
    assert eng.count == 0


def test_dimension_mismatch_is_encoding_error():
    eng = CUDAVectorEngine(VectorMetric.cosine, 4)
    eng.add(1, [1, 0, 0, 0])
    with pytest.raises(wax_b200.EncodingError, match="vector dimension mismatch: expected 4, got 3"):
        eng.search([1, 0, 0], 1)
    with pytest.raises(wax_b200.EncodingError):
        eng.add(2, [1, 0, 0])
    with pytest.raises(wax_b200.EncodingError):
        eng.add_batch([2, 3], [[1, 0, 0, 0], [1, 0, 0]])
    with pytest.raises(wax_b200.EncodingError, match="frameIds.count != vectors.count"):
        eng.add_batch([2, 3], [[1, 0, 0, 0]])
    eng.add_batch([], [])                        # empty batch: no-op (:360)
    assert eng.count == 1


def test_topk_clamp_and_min_k_n():
    eng = CUDAVectorEngine(VectorMetric.dot, 8)
    rng = np.random.default_rng(0)
    eng.add_batch(list(range(50)), rng.standard_normal((50, 8)).astype(np.float32))
    q = rng.standard_normal(8).astype(np.float32)
    assert len(eng.search(q, 0)) == 1 and len(eng.search(q, -7)) == 1   # clamp to 1 (:843)
    assert len(eng.search(q, 10)) == 10
    assert len(eng.search(q, 50_000)) == 50                             # clamp to 10000 then min(k, N)
    full = eng.search(q, 10_000)
    assert [i for i, _ in full[:10]] == [i for i, _ in eng.search(q, 10)]
    assert all(a[1] >= b[1] for a, b in zip(full, full[1:]))            # best first


@pytest.mark.parametrize("metric", list(VectorMetric))
def test_mutations_match_reference_semantics(oracle, metric):
    """Upsert overwrites in place, new ids append, remove keeps relative order (MetalVectorEngine.swift:330-444):
    checked by replaying a random op sequence on the list model and comparing full results + the raw rows."""
    dims = 12
    rng = np.random.default_rng(1 + metric.value)
    eng, model = CUDAVectorEngine(metric, dims), EngineModel(oracle, metric.value, dims)
    pool = list(range(100, 140))
    for step in range(120):
        op = rng.integers(0, 10)
        if op < 5:
            i, v = int(rng.choice(pool)), rng.standard_normal(dims).astype(np.float32)
            eng.add(i, v); model.add(i, v)
        elif op < 8:
            n = int(rng.integers(1, 9))
            ids = [int(x) for x in rng.choice(pool, n)]            # duplicates inside a batch allowed
            vs = rng.standard_normal((n, dims)).astype(np.float32)
            eng.add_batch(ids, vs); model.add_batch(ids, list(vs))
        else:
            i = int(rng.choice(pool + [999]))                      # 999: unknown id -> no-op
            eng.remove(i); model.remove(i)
        assert eng.count == len(model.ids)
        if step % 10 == 9 and model.ids:
            q = rng.standard_normal(dims).astype(np.float32)
            got, exp = eng.search(q, 15), model.search(q, 15)
            assert [g[0] for g in got] == [e[0] for e in exp]
            assert np.array_equal(np.float32([g[1] for g in got]), np.float32([e[1] for e in exp]))
            assert np.array_equal(eng.read_rows(0, eng.count), model.corpus())


def test_serialize_is_byte_identical_to_the_mv2v_layout(oracle):
    rng = np.random.default_rng(5)
    for metric in VectorMetric:
        eng = CUDAVectorEngine(metric, 6)
        vec = rng.standard_normal((9, 6)).astype(np.float32)
        ids = [int(x) for x in rng.integers(0, 2**63, 9, dtype=np.uint64)]
        eng.add_batch(ids, vec)
        blob = eng.serialize()
        assert blob == oracle.mv2v_encode(metric.value, vec, ids)      # MetalVectorEngine.swift:682-714
        other = CUDAVectorEngine(metric, 6)
        other.deserialize(blob)
        assert other.count == 9 and other.serialize() == blob
        q = rng.standard_normal(6).astype(np.float32)
        assert other.search(q, 9) == eng.search(q, 9)
    empty = CUDAVectorEngine(VectorMetric.cosine, 5)
    assert empty.serialize() == oracle.mv2v_encode(0, np.zeros((0, 5), np.float32), [])
    empty2 = CUDAVectorEngine(VectorMetric.cosine, 5)
    empty2.deserialize(empty.serialize())
    assert empty2.count == 0


def test_deserialize_rejects_malformed_blobs(oracle):
    vec = np.array([[1.0, -2.0], [0.5, 0.25]], np.float32)
    blob = oracle.mv2v_encode(0, vec, [7, 9])
    eng = CUDAVectorEngine(VectorMetric.cosine, 2)

    def corrupt(i, b):
        x = bytearray(blob); x[i] = b; return bytes(x)
    cases = [(blob[:20], "too small"), (corrupt(0, 0x58), "magic mismatch"), (corrupt(4, 2), "version"),
             (corrupt(6, 1), "encoding"), (corrupt(7, 1), "Metric mismatch"), (corrupt(8, 3), "Dimension mismatch"),
             (corrupt(30, 1), "reserved bytes must be zero"), (corrupt(20, 17), "Vector data length mismatch"),
             (corrupt(52, 8), "FrameId data length mismatch"), (blob + b"\x00", "length mismatch")]
    for bad, reason in cases:
        with pytest.raises(wax_b200.InvalidToc, match=reason):
            eng.deserialize(bad)
        assert oracle.mv2v_decode(bad, 0, 2)[0] != 0            # the oracle rejects the same blobs
    eng.deserialize(blob)
    assert eng.count == 2 and eng.search([1.0, -2.0], 1)[0][0] == 7


def test_stage_for_commit_follows_the_dirty_flag():
    class FakeWax:
        def __init__(self): self.calls = []
        def stage_vec_index_for_next_commit(self, **kw): self.calls.append(kw)
    wax, eng = FakeWax(), CUDAVectorEngine(VectorMetric.cosine, 2)
    eng.stage_for_commit(wax)
    assert wax.calls == []                       # not dirty -> nothing staged (:819)
    eng.add(3, [1.0, 0.0])
    eng.stage_for_commit(wax); eng.stage_for_commit(wax)
    assert len(wax.calls) == 1 and wax.calls[0]["vector_count"] == 1 and wax.calls[0]["dimension"] == 2
    assert wax.calls[0]["bytes"][:4] == b"MV2V" and wax.calls[0]["similarity"] == 0


def test_load_replays_pending_embeddings_over_the_committed_blob():
    """`static load(from:)` (MetalVectorEngine.swift:318-328): committed blob first, then every pending embedding
    as an upsert in order -- the situation of WaxSessionTests.swift:73-123 (search before and after commit)."""
    from collections import namedtuple
    PutEmbedding = namedtuple("PutEmbedding", "frame_id vector")
    src = CUDAVectorEngine(VectorMetric.cosine, 4)
    src.add_batch([0, 1, 2], [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]])

    class FakeWax:
        def __init__(self, blob, pending): self.blob, self.pending = blob, pending
        def read_committed_vec_index_bytes(self): return self.blob
        def pending_embedding_mutations(self): return self.pending

    pending = [PutEmbedding(3, [0.0, 0.0, 0.0, 1.0]), PutEmbedding(1, [0.6, 0.8, 0.0, 0.0]),   # 1: overwrite committed
               PutEmbedding(3, [0.0, 0.0, 0.6, 0.8]), (9, [0.5, 0.5, 0.5, 0.5])]                # 3: twice, last wins
    eng = CUDAVectorEngine.load(FakeWax(src.serialize(), pending), VectorMetric.cosine, 4)
    ref = CUDAVectorEngine(VectorMetric.cosine, 4)
    ref.deserialize(src.serialize())
    for m in pending:                              # the reference's loop: one add per pending embedding
        ref.add(m[0], m[1])
    assert eng.count == ref.count == 5
    assert eng.serialize() == ref.serialize()
    assert eng.search([0.0, 0.0, 0.6, 0.8], 1)[0][0] == 3
    assert eng.search([0.6, 0.8, 0.0, 0.0], 1)[0][0] == 1
    # nothing committed yet, nothing pending
    empty = CUDAVectorEngine.load(FakeWax(None, []), VectorMetric.cosine, 4)
    assert empty.count == 0 and empty.search([1, 0, 0, 0], 3) == []
    # a pending embedding of the wrong dimension surfaces as the reference's encodingError, engine released
    with pytest.raises(wax_b200.EncodingError):
        CUDAVectorEngine.load(FakeWax(None, [PutEmbedding(1, [1.0, 2.0])]), VectorMetric.cosine, 4)


def test_concurrent_searches_are_reentrant(oracle):
    eng = CUDAVectorEngine(VectorMetric.cosine, 64)
    eng.fill_synthetic(3, 20_000)
    qs = oracle.synth_rows(4, 0, 16, 64)
    expected = [eng.search(q, 10) for q in qs]
    errors = []

    def work(t):
        try:
            for rep in range(20):
                i = (t + rep) % len(qs)
                if eng.search(qs[i], 10) != expected[i]:
                    errors.append((t, rep))
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)
    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in threads]; [t.join() for t in threads]
    assert not errors
    allocs, reuses = eng.debug_buffer_pool_stats()
    assert allocs <= 9 and reuses > 0


def test_search_concurrent_with_add_never_reports_a_short_buffer(oracle):
    """A search whose result buffers were sized while another thread grows the corpus must still succeed (the mirror
    allocates clamp(topK) entries / re-sizes on ERR_BUFFER), and every answer must be a valid answer for SOME prefix of
    the add sequence: best-first, no duplicate ids, scores of the returned ids exact."""
    dims = 32
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    rows = oracle.synth_rows(21, 0, 600, dims)
    eng.add_batch([0, 1, 2], rows[:3])
    qs = oracle.synth_rows(22, 0, 4, dims)
    errors, stop = [], threading.Event()

    def searcher(t):
        try:
            while not stop.is_set():
                got = eng.search(qs[t % 4], 50)              # k > count at first: the buffer-size race window
                assert len({g[0] for g in got}) == len(got)
                assert all(a[1] >= b[1] for a, b in zip(got, got[1:]))
                batch = eng.search_batch(qs, 50)
                assert all(len({g[0] for g in b}) == len(b) for b in batch)
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)
    threads = [threading.Thread(target=searcher, args=(t,)) for t in range(4)]
    [t.start() for t in threads]
    for i in range(3, 600):
        eng.add(i, rows[i])
    stop.set()
    [t.join() for t in threads]
    assert not errors, errors[:1]
    r, _, s = oracle.search(oracle.COSINE, rows, qs[0], 50, mode=oracle.ACC_F32_TREE)
    got = eng.search(qs[0], 50)
    assert [g[0] for g in got] == r.tolist()


def test_mutator_drains_in_flight_device_path_scans(oracle):
    """wax_vs_search_device returns with its kernel still in flight on the caller's (non-blocking) stream; a mutator
    that follows must wait for it before moving rows: the candidates written by the in-flight scan are the
    pre-mutation answer, bit for bit."""
    import ctypes as C
    import torch
    from wax_b200 import _lib as L, sharded
    dims, n, k = 384, 2_000_000, 10
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    eng.fill_synthetic(23, n)
    q = oracle.synth_row(24, 0, dims, True)
    expect = eng.search(q, k)
    d_q = torch.from_numpy(q).cuda()
    stream = torch.cuda.Stream()
    buf = torch.zeros(k * 24, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    rc = L.lib().wax_vs_search_device(eng.handle, C.c_void_p(d_q.data_ptr()), 1, k, 0, C.c_void_p(buf.data_ptr()),
                                      C.c_void_p(stream.cuda_stream))
    assert rc == 0, L.last_error()
    eng.remove(expect[0][0])          # shifts ~all rows down by one: must not start under the scan
    stream.synchronize()
    cands = buf.cpu().numpy().view(sharded.CAND_DTYPE)
    assert [int(c["frame_id"]) for c in cands] == [e[0] for e in expect]
    after = eng.search(q, k)
    assert after[0][0] != expect[0][0] and [a[0] for a in after[:k - 1]] == [e[0] for e in expect[1:]]


@pytest.mark.parametrize("dims", [384, 640, 100, 1000, 2048])
def test_host_delivery_and_inline_query_do_not_change_results(oracle, dims):
    """The synchronous single-query entry point normally sends the query in the kernel parameters (<= 512 floats, TMA
    kernels) and lets the kernel store the result in mapped host memory + raise a flag.  Every combination of those two
    switches -- and the shapes that cannot use them (long rows, the direct-load kernel) -- returns the same bits."""
    eng = CUDAVectorEngine(VectorMetric.cosine, dims)
    eng.fill_synthetic(31, 30_000)
    qs = oracle.synth_rows(32, 0, 6, dims) * np.float32(1.7)
    r, _, s = oracle.search_synth(oracle.COSINE, 31, 0, 30_000, dims, True, qs[0], 10, mode=oracle.ACC_F32_TREE, threads=8)
    results = []
    for delivery in (1, 0):
        for inline in (1, 0):
            eng.set_option("host_delivery", delivery); eng.set_option("inline_query", inline)
            results.append([eng.search(q, k) for q in qs for k in (10, 72, 1)])
    assert all(x == results[0] for x in results[1:])
    assert [g[0] for g in results[0][0]] == r.tolist()
    assert np.array_equal(np.float32([g[1] for g in results[0][0]]).view(np.uint32), s.view(np.uint32))


def test_growth_from_initial_reserve(oracle):
    eng = CUDAVectorEngine(VectorMetric.l2, 3)           # initialReserve 64, doubling (:19, :857-871)
    rows = oracle.synth_rows(8, 0, 1000, 3, normalize=False)
    for start in range(0, 1000, 37):
        eng.add_batch(list(range(start, min(start + 37, 1000))), rows[start:start + 37])
    assert eng.count == 1000 and np.array_equal(eng.read_rows(0, 1000), rows)
    r, _, s = oracle.search(oracle.L2, rows, rows[500], 5, mode=oracle.ACC_F32_TREE)
    got = eng.search(rows[500], 5)
    assert [g[0] for g in got] == r.tolist() and got[0][0] == 500 and got[0][1] == 0.0


def test_cpp_mirror_runs_the_reference_tests_on_the_gpu(tmp_path):
    """The header-only C++ mirror over the C-ABI (the compiled-language host side): build and run its GPU test."""
    import subprocess
    from pathlib import Path
    from wax_b200 import build
    root = Path(__file__).resolve().parents[1]
    lib = build.build()
    exe = tmp_path / "cpp_engine_gpu"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", str(root / "tests" / "cpp_engine_gpu.cpp"), f"-L{lib.parent}",
                    "-lwaxvs_cuda", f"-Wl,-rpath,{lib.parent}", "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "cpp mirror ok" in out.stdout, (out.returncode, out.stdout, out.stderr)


def test_engines_on_two_devices_in_one_process(oracle):
    """`wax_vs_create(devices=[d])`: opt-in shared-memory limits (cudaFuncSetAttribute) are per device, so every
    kernel family must work on a second device of the same process (skipped on single-GPU boxes)."""
    import ctypes as C
    from wax_b200 import _lib as L
    n = C.c_int32(0)
    assert L.lib().wax_vs_device_count(C.byref(n)) == 0
    if n.value < 2:
        pytest.skip("needs two GPUs")
    dims, rows = 384, 30_000
    qs = oracle.synth_rows(71, 0, 130, dims)
    results = []
    for dev in (0, 1):
        eng = CUDAVectorEngine(VectorMetric.cosine, dims, device=dev)
        eng.fill_synthetic(70, rows)
        results.append((eng.search(qs[0], 10), eng.search(qs[0], 300), eng.search_batch(qs, 10),
                        eng.search_filtered(qs[0], 5, allow=list(range(0, rows, 7)))))
        eng.close()
    assert results[0] == results[1]
