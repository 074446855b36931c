"""Host-side mirror of Wax's `VectorSearchEngine` surface over the CUDA C-ABI.

The reference's host language is Swift, which this image cannot compile; `swift/CUDAVectorEngine.swift`
(uncompiled) is the literal binding.  This module is the same surface in Python so the
parity tests read like the reference's own tests (Tests/WaxIntegrationTests/VectorSearchEngineTests.swift):

    protocol VectorSearchEngine            Sources/WaxVectorSearch/VectorSearchEngine.swift:10-18
    enum VectorMetric                      Sources/WaxVectorSearch/VectorMetric.swift:5-54
    actor MetalVectorEngine (public API)   Sources/WaxVectorSearch/MetalVectorEngine.swift:144-146,153,330-446,682-828
    WaxVectorSearchSession.search          Sources/Wax/VectorSearchSession.swift:70-76
    VectorMath.normalizeL2/isNormalizedL2  Sources/Wax/Utilities/VectorMath.swift:15-33,123-127
    WaxError cases                         encodingError / capacityExceeded / invalidToc

Everything numeric happens in libwaxvs_cuda.so; nothing here computes a distance.
"""
from __future__ import annotations

import ctypes as C
import enum
import threading
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


# ---- WaxError (the cases the vector engines throw) ---------------------------------------------------------
class WaxError(Exception):
    pass


class EncodingError(WaxError):      # WaxError.encodingError(reason:)
    pass


class CapacityExceeded(WaxError):   # WaxError.capacityExceeded(limit:requested:)
    pass


class InvalidToc(WaxError):         # WaxError.invalidToc(reason:)
    pass


def _raise(rc: int) -> None:
    reason = L.last_error()
    if rc == L.ERR_DIMENSION:
        raise EncodingError(reason)
    if rc == L.ERR_CAPACITY:
        raise CapacityExceeded(reason)
    raise InvalidToc(f"{reason} (rc={rc})")


def _check(rc: int) -> None:
    if rc != L.OK:
        _raise(rc)


# ---- VectorMetric (VectorMetric.swift:5-54) -----------------------------------------------------------------
class VectorMetric(enum.Enum):
    cosine = 0
    dot = 1
    l2 = 2

    def to_vec_similarity(self) -> int:          # toVecSimilarity (:45-54); VecSimilarity raw value
        return self.value

    def score(self, from_distance: float) -> float:  # score(fromDistance:) (:32-43)
        d = np.float32(from_distance)
        if not np.isfinite(d):
            return 0.0
        return float(np.float32(1) - d) if self is VectorMetric.cosine else float(-d)


class VectorEnginePreference(enum.Enum):     # VectorSearchEngine.swift:4-8 (metalPreferred -> the GPU engine)
    auto = 0
    gpu_preferred = 1
    cpu_only = 2


# ---- VectorMath (host-side query preparation only) -------------------------------------------------------------
def normalize_l2(vector: Sequence[float]) -> np.ndarray:
    """VectorMath.normalizeL2 (VectorMath.swift:15-33): x * (1/|x|); empty or zero vectors unchanged."""
    v = np.ascontiguousarray(vector, dtype=np.float32)
    if v.size == 0:
        return v
    s = np.float32(0)
    for x in v:                       # vDSP_svesq: plain fp32 sum of squares
        s = np.float32(s + x * x)
    m = np.float32(np.sqrt(s))
    if not m > 0:
        return v
    return (v * np.float32(np.float32(1) / m)).astype(np.float32)


def is_normalized_l2(vector: Sequence[float], tolerance: float = 1e-3) -> bool:
    """VectorMath.isNormalizedL2 (VectorMath.swift:123-127)."""
    v = np.ascontiguousarray(vector, dtype=np.float32)
    if v.size == 0:
        return False
    s = np.float32(0)
    for x in v:
        s = np.float32(s + x * x)
    return bool(abs(np.float32(np.sqrt(s)) - np.float32(1)) <= np.float32(tolerance))


def _clamp_topk(top_k: int) -> int:
    """clampTopK (MetalVectorEngine.swift:842-846)."""
    return max(1, min(int(top_k), L.MAX_RESULTS))


def _as_rows(vectors, dims: int) -> np.ndarray:
    rows = [np.asarray(v, dtype=np.float32).reshape(-1) for v in vectors] if not isinstance(vectors, np.ndarray) \
        else None
    if rows is not None:
        for v in rows:
            if v.size != dims:  # MetalVectorEngine.swift:367-370
                raise EncodingError(f"vector dimension mismatch: expected {dims}, got {v.size}")
        return np.ascontiguousarray(np.stack(rows) if rows else np.zeros((0, dims), np.float32))
    arr = np.ascontiguousarray(vectors, dtype=np.float32)
    if arr.ndim != 2 or arr.shape[1] != dims:
        got = arr.shape[1] if arr.ndim == 2 else arr.size
        raise EncodingError(f"vector dimension mismatch: expected {dims}, got {got}")
    return arr


class CUDAVectorEngine:
    """Drop-in for MetalVectorEngine / USearchVectorEngine behind `VectorSearchEngine`.

    Thread-safety mirrors the actor + AsyncReadWriteLock: searches may run concurrently, mutators are
    exclusive (enforced inside the library).
    """

    @staticmethod
    def is_available() -> bool:                      # MetalVectorEngine.isAvailable (:144-146)
        n = C.c_int32(0)
        return L.lib().wax_vs_device_count(C.byref(n)) == L.OK and n.value > 0

    def __init__(self, metric: VectorMetric = VectorMetric.cosine, dimensions: int = 0,
                 device: Optional[int] = None):     # init(metric:dimensions:) (:153)
        if dimensions <= 0:
            raise InvalidToc("dimensions must be > 0")
        if dimensions > L.MAX_DIMENSIONS:
            raise CapacityExceeded(f"capacity exceeded: limit {L.MAX_DIMENSIONS}, requested {dimensions}")
        self.metric = metric
        self.dimensions = int(dimensions)
        self._dirty = False
        self._h = C.c_void_p()
        devs = (C.c_int32 * 1)(device) if device is not None else None
        _check(L.lib().wax_vs_create(self.dimensions, metric.to_vec_similarity(), devs,
                                     1 if device is not None else 0, C.byref(self._h)))
        self._closed = False
        self._lock = threading.Lock()

    @classmethod
    def load(cls, wax, metric: VectorMetric, dimensions: int, device: Optional[int] = None) -> "CUDAVectorEngine":
        """`static load(from:metric:dimensions:)` (MetalVectorEngine.swift:318-328): the committed vector-index blob,
        then the pending (uncommitted) embedding mutations replayed in order as upserts.  `wax` needs
        `read_committed_vec_index_bytes() -> bytes | None` and `pending_embedding_mutations() -> [(frameId, vector)]`
        (objects with `.frame_id` / `.vector` are accepted too); the store behind them is out of scope.
        The replay is ONE add_batch: the library resolves the rows sequentially, so a frameId that occurs twice keeps
        its last vector exactly as the reference's per-embedding loop does."""
        engine = cls(metric, dimensions, device)
        try:
            blob = wax.read_committed_vec_index_bytes()
            if blob is not None:
                engine.deserialize(blob)
            pending = list(wax.pending_embedding_mutations())
            if pending:
                ids = [int(getattr(m, "frame_id", m[0] if isinstance(m, (tuple, list)) else None)) for m in pending]
                vecs = [getattr(m, "vector", m[1] if isinstance(m, (tuple, list)) else None) for m in pending]
                engine.add_batch(ids, vecs)
        except Exception:
            engine.close()
            raise
        return engine

    # -- lifetime
    def close(self) -> None:
        if not getattr(self, "_closed", True):
            self._closed = True
            L.lib().wax_vs_destroy(self._h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def count(self) -> int:
        n = C.c_uint64(0)
        _check(L.lib().wax_vs_count(self._h, C.byref(n)))
        return n.value

    # -- VectorSearchEngine protocol
    def search(self, vector: Sequence[float], top_k: int) -> List[Tuple[int, float]]:
        """search(vector:topK:) -> [(frameId, score)] best first."""
        q = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        # clamp(topK) entries, as the Swift mirror allocates: sizing from a separate count() call would race with a
        # concurrent add (the library would then need more room than `cap` and report ERR_BUFFER on a valid search)
        cap = _clamp_topk(top_k)
        ids = np.empty(cap, np.uint64)
        scores = np.empty(cap, np.float32)
        n = C.c_uint32(0)
        _check(L.lib().wax_vs_search(self._h, q.ctypes.data_as(C.POINTER(C.c_float)), q.size, int(top_k),
                                     ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                     scores.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return [(int(ids[i]), float(scores[i])) for i in range(n.value)]

    def search_filtered(self, vector: Sequence[float], top_k: int, allow: Optional[Sequence[int]] = None,
                        deny: Optional[Sequence[int]] = None) -> List[Tuple[int, float]]:
        """Filter pushed below the top-k (API extension): the best `top_k` rows among the frames in
        `allow` (allow-list) or among all frames except those in `deny` (deny-list).  Replaces the reference's
        post-hoc frame filter + 3 x topK over-fetch (UnifiedSearch.swift:58,1241-1258)."""
        if (allow is None) == (deny is None):
            raise ValueError("pass exactly one of allow= / deny=")
        ids = np.ascontiguousarray(allow if allow is not None else deny, dtype=np.uint64).reshape(-1)
        q = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        cap = _clamp_topk(top_k)
        out_ids = np.empty(cap, np.uint64)
        scores = np.empty(cap, np.float32)
        n = C.c_uint32(0)
        idp = ids.ctypes.data_as(C.POINTER(C.c_uint64)) if ids.size else None
        _check(L.lib().wax_vs_search_filtered(self._h, q.ctypes.data_as(C.POINTER(C.c_float)), q.size, int(top_k), idp,
                                              ids.size, 0 if allow is not None else 1,
                                              out_ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                              scores.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return [(int(out_ids[i]), float(scores[i])) for i in range(n.value)]

    def search_batch_filtered(self, vectors, top_k: int, allow: Optional[Sequence[int]] = None,
                              deny: Optional[Sequence[int]] = None) -> List[List[Tuple[int, float]]]:
        """`search_filtered` for a batch of queries under ONE filter, in one pass over the corpus
        (wax_vs_search_batch_filtered); the same answers as calling search_filtered per query."""
        if (allow is None) == (deny is None):
            raise ValueError("pass exactly one of allow= / deny=")
        fids = np.ascontiguousarray(allow if allow is not None else deny, dtype=np.uint64).reshape(-1)
        qs = _as_rows(vectors, self.dimensions) if len(vectors) else np.zeros((0, self.dimensions), np.float32)
        b = qs.shape[0]
        if b == 0:
            return []
        cap = _clamp_topk(top_k)
        ids = np.zeros((b, cap), np.uint64)
        scores = np.zeros((b, cap), np.float32)
        ns = np.zeros(b, np.uint32)
        idp = fids.ctypes.data_as(C.POINTER(C.c_uint64)) if fids.size else None
        _check(L.lib().wax_vs_search_batch_filtered(self._h, qs.ctypes.data_as(C.POINTER(C.c_float)), b, qs.shape[1],
                                                    int(top_k), idp, fids.size, 0 if allow is not None else 1,
                                                    ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                    scores.ctypes.data_as(C.POINTER(C.c_float)), cap,
                                                    ns.ctypes.data_as(C.POINTER(C.c_uint32))))
        return [[(int(ids[i, j]), float(scores[i, j])) for j in range(int(ns[i]))] for i in range(b)]

    def search_batch(self, vectors, top_k: int) -> List[List[Tuple[int, float]]]:
        ids, scores, ns = self.search_batch_arrays(vectors, top_k)
        return [[(int(ids[i, j]), float(scores[i, j])) for j in range(int(ns[i]))] for i in range(ids.shape[0])]

    def search_batch_arrays(self, vectors, top_k: int):
        """Batched search returning arrays: (ids u64 [B, k_eff], scores f32 [B, k_eff], counts u32 [B]); row i holds
        counts[i] results, best first.  The list-of-tuples form above costs more host time than the GPU pass."""
        qs = _as_rows(vectors, self.dimensions) if len(vectors) else np.zeros((0, self.dimensions), np.float32)
        b = qs.shape[0]
        if b == 0:
            return np.zeros((0, 0), np.uint64), np.zeros((0, 0), np.float32), np.zeros(0, np.uint32)
        # b x min(clamp(topK), count) entries; a concurrent add can grow the count between the two calls, in which
        # case the library answers ERR_BUFFER and the buffers are re-sized (clamp(topK) always suffices)
        cap = max(1, min(_clamp_topk(top_k), max(self.count, 1)))
        for attempt in range(3):
            ids = np.zeros((b, cap), np.uint64)
            scores = np.zeros((b, cap), np.float32)
            ns = np.zeros(b, np.uint32)
            rc = L.lib().wax_vs_search_batch(self._h, qs.ctypes.data_as(C.POINTER(C.c_float)), b, qs.shape[1],
                                             int(top_k), ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                             scores.ctypes.data_as(C.POINTER(C.c_float)), cap,
                                             ns.ctypes.data_as(C.POINTER(C.c_uint32)))
            if rc != L.ERR_BUFFER:
                break
            cap = _clamp_topk(top_k) if attempt else max(1, min(_clamp_topk(top_k), max(self.count, 1)))
        _check(rc)
        return ids, scores, ns

    def add(self, frame_id: int, vector: Sequence[float]) -> None:
        v = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        _check(L.lib().wax_vs_add(self._h, int(frame_id), v.ctypes.data_as(C.POINTER(C.c_float)), v.size))
        self._dirty = True

    def add_batch(self, frame_ids: Sequence[int], vectors) -> None:
        ids = np.ascontiguousarray(frame_ids, dtype=np.uint64).reshape(-1)
        if ids.size == 0:                         # guard !frameIds.isEmpty (:360)
            return
        if ids.size != len(vectors):              # :361-363
            raise EncodingError("addBatch: frameIds.count != vectors.count")
        rows = _as_rows(vectors, self.dimensions)
        _check(L.lib().wax_vs_add_batch(self._h, ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        rows.ctypes.data_as(C.POINTER(C.c_float)), ids.size, rows.shape[1]))
        self._dirty = True

    def add_batch_streaming(self, frame_ids: Sequence[int], vectors, chunk_size: int = 256) -> None:
        """addBatchStreaming (:404-421)."""
        if len(frame_ids) == 0:
            return
        if len(frame_ids) != len(vectors):
            raise EncodingError("addBatchStreaming: frameIds.count != vectors.count")
        if len(frame_ids) <= chunk_size:
            return self.add_batch(frame_ids, vectors)
        for start in range(0, len(frame_ids), chunk_size):
            self.add_batch(frame_ids[start:start + chunk_size], vectors[start:start + chunk_size])

    def remove(self, frame_id: int) -> None:
        before = self.count
        _check(L.lib().wax_vs_remove(self._h, int(frame_id)))
        if self.count != before:
            self._dirty = True

    def remove_batch(self, frame_ids: Sequence[int]) -> int:
        """remove(frameId:) for many frames in one pass (one compaction, one hash rebuild); returns how many rows went.
        Same result as calling remove() for each id."""
        ids = np.ascontiguousarray(frame_ids, dtype=np.uint64).reshape(-1)
        if ids.size == 0:
            return 0
        gone = C.c_uint64(0)
        _check(L.lib().wax_vs_remove_batch(self._h, ids.ctypes.data_as(C.POINTER(C.c_uint64)), ids.size, C.byref(gone)))
        if gone.value:
            self._dirty = True
        return gone.value

    def reserve(self, rows: int) -> None:
        _check(L.lib().wax_vs_reserve(self._h, int(rows)))

    # -- row-sharded search (wax_vs_shard_*; no reference counterpart) ------------------------------
    def shard_open(self, rank: int, world: int, row_offset: int) -> bytes:
        """Become rank `rank` of `world`: allocates this engine's mailbox and returns the handle blob the other ranks
        need to reach it."""
        blob = (C.c_uint8 * L.SHARD_HANDLE_BYTES)()
        _check(L.lib().wax_vs_shard_open(self._h, int(rank), int(world), int(row_offset), blob))
        return bytes(blob)

    def shard_connect(self, blobs: Sequence[bytes]) -> None:
        """Map every rank's mailbox (`blobs` in rank order, this rank's own included)."""
        flat = b"".join(blobs)
        buf = (C.c_uint8 * len(flat)).from_buffer_copy(flat)
        _check(L.lib().wax_vs_shard_connect(self._h, buf, len(blobs)))

    def shard_close(self) -> None:
        _check(L.lib().wax_vs_shard_close(self._h))

    def shard_search(self, vector: Sequence[float], top_k: int) -> List[Tuple[int, float]]:
        """COLLECTIVE search over the whole sharded corpus (every rank calls it with the same query): scan + NVLink
        exchange + merge in one kernel launch, merged result delivered into host memory."""
        q = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        cap = _clamp_topk(top_k)
        ids = np.empty(cap, np.uint64)
        scores = np.empty(cap, np.float32)
        n = C.c_uint32(0)
        _check(L.lib().wax_vs_shard_search(self._h, q.ctypes.data_as(C.POINTER(C.c_float)), q.size, int(top_k),
                                           ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                           scores.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return [(int(ids[i]), float(scores[i])) for i in range(n.value)]

    def shard_search_filtered(self, vector: Sequence[float], top_k: int, allow: Optional[Sequence[int]] = None,
                              deny: Optional[Sequence[int]] = None) -> List[Tuple[int, float]]:
        """COLLECTIVE filtered search (every rank passes the same query and the same ids): each rank's fused scan
        applies the filter to the rows of its own shard, the in-kernel exchange merges -- one launch per rank."""
        if (allow is None) == (deny is None):
            raise ValueError("pass exactly one of allow= / deny=")
        fids = np.ascontiguousarray(allow if allow is not None else deny, dtype=np.uint64).reshape(-1)
        q = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        cap = _clamp_topk(top_k)
        ids = np.empty(cap, np.uint64)
        scores = np.empty(cap, np.float32)
        n = C.c_uint32(0)
        idp = fids.ctypes.data_as(C.POINTER(C.c_uint64)) if fids.size else None
        _check(L.lib().wax_vs_shard_search_filtered(self._h, q.ctypes.data_as(C.POINTER(C.c_float)), q.size, int(top_k),
                                                    idp, fids.size, 0 if allow is not None else 1,
                                                    ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                    scores.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return [(int(ids[i]), float(scores[i])) for i in range(n.value)]

    def time_shard_search(self, top_k: int, iters: int, warmup: int = 3, n_queries: int = 1, seed: int = 7):
        """Device-timed collective searches, strictly one at a time on one stream. Returns (ms_total, launches)."""
        ms, launches = C.c_float(0), C.c_uint64(0)
        _check(L.lib().wax_vs_debug_time_shard_search(self._h, n_queries, int(top_k), seed, warmup, iters,
                                                      C.byref(ms), C.byref(launches)))
        return ms.value, launches.value

    # -- persistence (MV2V encoding = 2)
    def serialize(self) -> bytes:
        n = C.c_uint64(0)
        _check(L.lib().wax_vs_serialized_length(self._h, C.byref(n)))
        buf = bytearray(n.value)          # calloc'ed, written once by the library: no zero-fill pass, no trailing copy
        out = C.c_uint64(0)
        ptr = (C.c_uint8 * len(buf)).from_buffer(buf) if buf else (C.c_uint8 * 1)()
        _check(L.lib().wax_vs_serialize(self._h, ptr, len(buf), C.byref(out)))
        del ptr
        if out.value != len(buf):
            del buf[out.value:]
        return buf                          # bytes-like (compares equal to bytes; pass to deserialize() as is)

    def deserialize(self, data: bytes) -> None:
        buf = np.frombuffer(data, np.uint8)     # zero-copy view of bytes / bytearray / memoryview
        ptr = buf.ctypes.data_as(C.POINTER(C.c_uint8)) if buf.size else C.cast(C.c_char_p(b""), C.POINTER(C.c_uint8))
        _check(L.lib().wax_vs_deserialize(self._h, ptr, buf.size))
        self._dirty = False

    def stage_for_commit(self, into) -> None:
        """stageForCommit(into:) (:818-828): `into` needs stage_vec_index_for_next_commit(bytes, vector_count,
        dimension, similarity) -- the Wax store itself is out of scope."""
        if not self._dirty:
            return
        into.stage_vec_index_for_next_commit(bytes=self.serialize(), vector_count=self.count,
                                             dimension=self.dimensions,
                                             similarity=self.metric.to_vec_similarity())
        self._dirty = False

    # -- instrumentation
    def debug_buffer_pool_stats(self) -> Tuple[int, int]:
        a, r = C.c_uint64(0), C.c_uint64(0)
        _check(L.lib().wax_vs_debug_pool_stats(self._h, C.byref(a), C.byref(r)))
        return a.value, r.value

    def fill_synthetic(self, seed: int, rows: int, first_row: int = 0, id_base: int = 0,
                       normalize: bool = True) -> None:
        _check(L.lib().wax_vs_debug_fill_synthetic(self._h, seed, first_row, rows, id_base, int(normalize)))
        self._dirty = True

    def read_rows(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dimensions), np.float32)
        _check(L.lib().wax_vs_debug_read_rows(self._h, first, n, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def time_search(self, top_k: int, iters: int, warmup: int = 3, n_queries: int = 1, seed: int = 7):
        """Kernel-only timing (CUDA events on the launching stream). Returns (ms_total, launches)."""
        ms, launches = C.c_float(0), C.c_uint64(0)
        _check(L.lib().wax_vs_debug_time_search(self._h, n_queries, int(top_k), seed, warmup, iters,
                                                C.byref(ms), C.byref(launches)))
        return ms.value, launches.value

    def batch_stats(self) -> Tuple[int, int]:
        """(queries answered by the tensor-core path with a completed proof, queries re-run exactly)."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(L.lib().wax_vs_debug_batch_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def counter(self, name: str) -> int:
        """Named instrumentation counter (wax_vs_debug_counter): batch_bf16_queries, batch_retry_queries, shadow_bytes, ..."""
        v = C.c_uint64(0)
        _check(L.lib().wax_vs_debug_counter(self._h, name.encode(), C.byref(v)))
        return v.value

    def time_search_batch(self, n_queries: int, top_k: int, iters: int, warmup: int = 2, seed: int = 7):
        """Device-only timing of the batched path. Returns (ms_total, launches, unproven_in_last_step)."""
        ms, launches, bad = C.c_float(0), C.c_uint64(0), C.c_uint32(0)
        _check(L.lib().wax_vs_debug_time_search_batch(self._h, n_queries, int(top_k), seed, warmup, iters,
                                                      C.byref(ms), C.byref(launches), C.byref(bad)))
        return ms.value, launches.value, bad.value

    def stream_read_gbs(self, iters: int = 5) -> float:
        """Plain coalesced read of the corpus bytes: the box's streaming-read ceiling in GB/s."""
        ms, nbytes = C.c_float(0), C.c_uint64(0)
        _check(L.lib().wax_vs_debug_stream_read(self._h, iters, C.byref(ms), C.byref(nbytes)))
        return nbytes.value / (ms.value * 1e6) if ms.value > 0 else 0.0

    def set_option(self, key: str, value: int) -> None:
        _check(L.lib().wax_vs_debug_set_option(self._h, key.encode(), int(value)))


class VectorSearchSession:
    """The score-preserving entry `WaxVectorSearchSession.search` (VectorSearchSession.swift:70-76):
    cosine queries that are not unit length (tolerance 1e-3) are normalised on the host first."""

    def __init__(self, engine: CUDAVectorEngine):
        self.engine = engine
        self.metric = engine.metric

    def search(self, vector: Sequence[float], top_k: int):
        q = np.ascontiguousarray(vector, dtype=np.float32)
        if self.metric is VectorMetric.cosine and q.size and not is_normalized_l2(q):
            q = normalize_l2(q)
        return self.engine.search(q, top_k)
