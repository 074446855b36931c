"""ctypes binding of the CPU ORACLE (oracle/wax_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from wax_b200/ (the product path has no CPU fallback).
See oracle/wax_oracle.h for what each function restates (reference file:line) and for the
"parity unpinned" statement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libwax_oracle.so"

COSINE, DOT, L2 = 0, 1, 2
ACC_F32_SEQ, ACC_F64, ACC_F32_TREE = 0, 1, 2
MAX_RESULTS = 10_000


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src_m = max((_HERE / n).stat().st_mtime for n in ("wax_oracle.c", "wax_oracle.h", "Makefile"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(_LIB_PATH))
        f32p, u64p, u32p, u8p = (C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint8))
        L = _lib
        L.wax_oracle_score_from_distance.restype = C.c_float
        L.wax_oracle_score_from_distance.argtypes = [C.c_int, C.c_float]
        L.wax_oracle_clamp_topk.restype = C.c_int64
        L.wax_oracle_clamp_topk.argtypes = [C.c_int64]
        L.wax_oracle_distance.restype = C.c_float
        L.wax_oracle_distance.argtypes = [C.c_int, C.c_int, f32p, f32p, C.c_uint32]
        L.wax_oracle_metal_cosine_distance.restype = C.c_float
        L.wax_oracle_metal_cosine_distance.argtypes = [f32p, f32p, C.c_uint32]
        L.wax_oracle_normalize_l2.restype = None
        L.wax_oracle_normalize_l2.argtypes = [f32p, f32p, C.c_uint32]
        L.wax_oracle_is_normalized_l2.restype = C.c_int
        L.wax_oracle_is_normalized_l2.argtypes = [f32p, C.c_uint32, C.c_float]
        L.wax_oracle_search.restype = C.c_int
        L.wax_oracle_search.argtypes = [C.c_int, C.c_int, f32p, C.c_uint64, C.c_uint32, f32p, C.c_int64,
                                        C.c_uint64, C.c_int, u64p, f32p, f32p, u32p]
        L.wax_oracle_search_synth.restype = C.c_int
        L.wax_oracle_search_synth.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                              C.c_uint32, C.c_int, f32p, C.c_int64, C.c_int, u64p, f32p,
                                              f32p, u32p]
        L.wax_oracle_search_multi.restype = C.c_int
        L.wax_oracle_search_multi.argtypes = [C.c_int, C.c_int, f32p, C.c_uint64, C.c_uint32, f32p, C.c_uint32,
                                              C.c_int64, C.c_uint64, C.c_int, u64p, f32p, f32p, u32p]
        L.wax_oracle_search_synth_multi.restype = C.c_int
        L.wax_oracle_search_synth_multi.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                                    C.c_uint32, C.c_int, f32p, C.c_uint32, C.c_int64, C.c_int,
                                                    u64p, f32p, f32p, u32p]
        L.wax_oracle_metal_cpu_topk.restype = C.c_uint32
        L.wax_oracle_metal_cpu_topk.argtypes = [f32p, C.c_uint64, C.c_uint32, u64p, f32p]
        L.wax_oracle_synth_row.restype = None
        L.wax_oracle_synth_row.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, f32p]
        L.wax_oracle_synth_rows.restype = None
        L.wax_oracle_synth_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int,
                                            C.c_int, f32p]
        L.wax_oracle_mv2v_length.restype = C.c_uint64
        L.wax_oracle_mv2v_length.argtypes = [C.c_uint32, C.c_uint64]
        L.wax_oracle_mv2v_encode.restype = C.c_int
        L.wax_oracle_mv2v_encode.argtypes = [C.c_uint8, C.c_uint32, C.c_uint64, f32p, u64p, u8p,
                                             C.c_uint64, u64p]
        L.wax_oracle_mv2v_decode.restype = C.c_int
        L.wax_oracle_mv2v_decode.argtypes = [u8p, C.c_uint64, C.c_uint8, C.c_uint32, u64p,
                                             C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def host_threads() -> int:
    """Host cores this process can really use: the affinity mask, capped by the cgroup CPU quota when there is one
    (a container may see 128 CPUs in its mask and still be limited to a few cores' worth of time)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        try:
            q = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            per = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0 and per > 0:
                n = max(1, min(n, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def score_from_distance(metric: int, d: float) -> float:
    return float(lib().wax_oracle_score_from_distance(metric, C.c_float(d)))


def clamp_topk(k: int) -> int:
    return int(lib().wax_oracle_clamp_topk(int(k)))


def distance(metric: int, mode: int, a, b) -> float:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape and a.ndim == 1
    return float(lib().wax_oracle_distance(metric, mode, _p(a, C.c_float), _p(b, C.c_float), a.size))


def metal_cosine_distance(q, v) -> float:
    q, v = _f32(q), _f32(v)
    return float(lib().wax_oracle_metal_cosine_distance(_p(q, C.c_float), _p(v, C.c_float), q.size))


def normalize_l2(v) -> np.ndarray:
    v = _f32(v)
    out = np.empty_like(v)
    lib().wax_oracle_normalize_l2(_p(v, C.c_float), _p(out, C.c_float), v.size)
    return out


def is_normalized_l2(v, tol: float = 1e-3) -> bool:
    v = _f32(v)
    return bool(lib().wax_oracle_is_normalized_l2(_p(v, C.c_float), v.size, C.c_float(tol)))


def search(metric: int, corpus, query, top_k: int, mode: int = ACC_F32_SEQ, row_base: int = 0,
           threads: int = 1):
    """Exact scan.  Returns (rows u64[n], distances f32[n], scores f32[n])."""
    corpus, query = _f32(corpus), _f32(query)
    n_rows, dims = (corpus.shape if corpus.ndim == 2 else (0, query.size))
    if n_rows and query.size != dims:
        raise ValueError(f"vector dimension mismatch: expected {dims}, got {query.size}")
    cap = max(1, min(clamp_topk(top_k), max(n_rows, 1)))
    rows = np.zeros(cap, np.uint64); d = np.zeros(cap, np.float32); s = np.zeros(cap, np.float32)
    n = C.c_uint32(0)
    rc = lib().wax_oracle_search(metric, mode, _p(corpus, C.c_float), n_rows, dims, _p(query, C.c_float),
                                 int(top_k), row_base, threads, _p(rows, C.c_uint64), _p(d, C.c_float),
                                 _p(s, C.c_float), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"wax_oracle_search rc={rc}")
    return rows[: n.value].copy(), d[: n.value].copy(), s[: n.value].copy()


def search_synth(metric: int, seed: int, first_row: int, n_rows: int, dims: int, normalize: bool, query,
                 top_k: int, mode: int = ACC_F32_SEQ, threads: int = 1):
    query = _f32(query)
    cap = max(1, min(clamp_topk(top_k), max(n_rows, 1)))
    rows = np.zeros(cap, np.uint64); d = np.zeros(cap, np.float32); s = np.zeros(cap, np.float32)
    n = C.c_uint32(0)
    rc = lib().wax_oracle_search_synth(metric, mode, seed, first_row, n_rows, dims, int(normalize),
                                       _p(query, C.c_float), int(top_k), threads, _p(rows, C.c_uint64),
                                       _p(d, C.c_float), _p(s, C.c_float), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"wax_oracle_search_synth rc={rc}")
    return rows[: n.value].copy(), d[: n.value].copy(), s[: n.value].copy()


def search_multi(metric: int, corpus, queries, top_k: int, mode: int = ACC_F32_SEQ, row_base: int = 0,
                 threads: int = 1):
    """Exact scan for several queries in one pass.  Returns (rows u64[B,k], distances f32[B,k], scores f32[B,k],
    counts u32[B])."""
    corpus, queries = _f32(corpus), _f32(queries)
    n_rows, dims = corpus.shape
    queries = queries.reshape(-1, dims)
    b = queries.shape[0]
    cap = max(1, min(clamp_topk(top_k), max(n_rows, 1)))
    rows = np.zeros((b, cap), np.uint64); d = np.zeros((b, cap), np.float32); s = np.zeros((b, cap), np.float32)
    n = np.zeros(b, np.uint32)
    rc = lib().wax_oracle_search_multi(metric, mode, _p(corpus, C.c_float), n_rows, dims, _p(queries, C.c_float), b,
                                       int(top_k), row_base, threads, _p(rows, C.c_uint64), _p(d, C.c_float),
                                       _p(s, C.c_float), _p(n, C.c_uint32))
    if rc != 0:
        raise RuntimeError(f"wax_oracle_search_multi rc={rc}")
    return rows, d, s, n


def search_synth_multi(metric: int, seed: int, first_row: int, n_rows: int, dims: int, normalize: bool, queries,
                       top_k: int, mode: int = ACC_F32_SEQ, threads: int = 1):
    """search_synth for several queries: each synthetic row is generated once and scored against all of them."""
    queries = _f32(queries).reshape(-1, dims)
    b = queries.shape[0]
    cap = max(1, min(clamp_topk(top_k), max(n_rows, 1)))
    rows = np.zeros((b, cap), np.uint64); d = np.zeros((b, cap), np.float32); s = np.zeros((b, cap), np.float32)
    n = np.zeros(b, np.uint32)
    rc = lib().wax_oracle_search_synth_multi(metric, mode, seed, first_row, n_rows, dims, int(normalize),
                                             _p(queries, C.c_float), b, int(top_k), threads, _p(rows, C.c_uint64),
                                             _p(d, C.c_float), _p(s, C.c_float), _p(n, C.c_uint32))
    if rc != 0:
        raise RuntimeError(f"wax_oracle_search_synth_multi rc={rc}")
    return rows, d, s, n


def metal_cpu_topk(distances, k: int):
    distances = _f32(distances)
    cap = max(1, min(k, distances.size))
    rows = np.zeros(cap, np.uint64); d = np.zeros(cap, np.float32)
    n = lib().wax_oracle_metal_cpu_topk(_p(distances, C.c_float), distances.size, k, _p(rows, C.c_uint64),
                                        _p(d, C.c_float))
    return rows[:n].copy(), d[:n].copy()


def synth_row(seed: int, row: int, dims: int, normalize: bool = True) -> np.ndarray:
    out = np.empty(dims, np.float32)
    lib().wax_oracle_synth_row(seed, row, dims, int(normalize), _p(out, C.c_float))
    return out


def synth_rows(seed: int, first_row: int, n_rows: int, dims: int, normalize: bool = True,
               threads: int = 0) -> np.ndarray:
    out = np.empty((n_rows, dims), np.float32)
    lib().wax_oracle_synth_rows(seed, first_row, n_rows, dims, int(normalize),
                                threads or host_threads(), _p(out, C.c_float))
    return out


def mv2v_encode(similarity: int, vectors, frame_ids) -> bytes:
    vectors = _f32(vectors)
    ids = np.ascontiguousarray(frame_ids, dtype=np.uint64)
    count, dims = vectors.shape
    assert ids.size == count
    need = int(lib().wax_oracle_mv2v_length(dims, count))
    buf = np.zeros(need, np.uint8)
    out_len = C.c_uint64(0)
    rc = lib().wax_oracle_mv2v_encode(similarity, dims, count, _p(vectors, C.c_float), _p(ids, C.c_uint64),
                                      _p(buf, C.c_uint8), need, C.byref(out_len))
    assert rc == 0 and out_len.value == need
    return buf.tobytes()


def mv2v_decode(blob: bytes, similarity: int, dims: int):
    """Returns (rc, vectors[count,dims] or None, ids[count] or None)."""
    buf = np.frombuffer(blob, np.uint8)
    count = C.c_uint64(0); pv = C.c_void_p(); pi = C.c_void_p()
    rc = lib().wax_oracle_mv2v_decode(_p(buf, C.c_uint8), buf.size, similarity, dims, C.byref(count),
                                      C.byref(pv), C.byref(pi))
    if rc != 0:
        return rc, None, None
    n = count.value
    vec = np.frombuffer(blob, np.float32, n * dims, 36).reshape(n, dims).copy()
    ids = np.frombuffer(blob, np.uint64, n, 36 + n * dims * 4 + 8).copy() if n else np.zeros(0, np.uint64)
    return 0, vec, ids
