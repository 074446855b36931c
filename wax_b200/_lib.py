"""ctypes loader for libwaxvs_cuda.so (the C-ABI in include/wax_vs_cuda.h).

The library is the product: there is no Python/CPU fallback.  If the shared object is missing this module
raises at import of the symbol table; if no CUDA device is present every engine call returns
WAX_VS_ERR_CUDA, surfaced as WaxError.invalidToc by engine.py (as MetalVectorEngine does for a missing
Metal device, MetalVectorEngine.swift:167-169).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libwaxvs_cuda.so"

# return codes (include/wax_vs_cuda.h)
OK, ERR_NULL, ERR_DIMENSION, ERR_CAPACITY, ERR_CUDA, ERR_FORMAT, ERR_ARGUMENT, ERR_BUFFER, ERR_UNSUPPORTED = (
    0, -1, -2, -3, -4, -5, -6, -7, -8)
MAX_RESULTS = 10_000
MAX_DIMENSIONS = 1_000_000
SHARD_HANDLE_BYTES, SHARD_MAX_RANKS, SHARD_MAX_K = 128, 16, 128


class Candidate(C.Structure):
    """wax_vs_candidate (24 bytes)."""
    _fields_ = [("distance", C.c_float), ("valid", C.c_uint32), ("row", C.c_uint64), ("frame_id", C.c_uint64)]


# Every symbol include/wax_vs_cuda.h declares, with its signature.  tests/test_abi.py checks this table
# against the header and against the built library.
_f32p, _u64p, _u32p, _u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
_eng = C.c_void_p
SIGNATURES = {
    "wax_vs_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "wax_vs_create": (C.c_int32, [C.c_uint32, C.c_uint8, C.POINTER(C.c_int32), C.c_int32, C.POINTER(_eng)]),
    "wax_vs_destroy": (None, [_eng]),
    "wax_vs_dimensions": (C.c_int32, [_eng, _u32p]),
    "wax_vs_similarity": (C.c_int32, [_eng, _u8p]),
    "wax_vs_count": (C.c_int32, [_eng, _u64p]),
    "wax_vs_reserve": (C.c_int32, [_eng, C.c_uint64]),
    "wax_vs_add": (C.c_int32, [_eng, C.c_uint64, _f32p, C.c_uint32]),
    "wax_vs_add_batch": (C.c_int32, [_eng, _u64p, _f32p, C.c_uint64, C.c_uint32]),
    "wax_vs_remove": (C.c_int32, [_eng, C.c_uint64]),
    "wax_vs_remove_batch": (C.c_int32, [_eng, _u64p, C.c_uint64, _u64p]),
    "wax_vs_search": (C.c_int32, [_eng, _f32p, C.c_uint32, C.c_int64, _u64p, _f32p, C.c_uint32, _u32p]),
    "wax_vs_search_filtered": (C.c_int32, [_eng, _f32p, C.c_uint32, C.c_int64, _u64p, C.c_uint64, C.c_int32, _u64p, _f32p,
                                           C.c_uint32, _u32p]),
    "wax_vs_search_batch_filtered": (C.c_int32, [_eng, _f32p, C.c_uint32, C.c_uint32, C.c_int64, _u64p, C.c_uint64, C.c_int32,
                                                 _u64p, _f32p, C.c_uint32, _u32p]),
    "wax_vs_shard_search_filtered": (C.c_int32, [_eng, _f32p, C.c_uint32, C.c_int64, _u64p, C.c_uint64, C.c_int32, _u64p,
                                                 _f32p, C.c_uint32, _u32p]),
    "wax_vs_search_batch": (C.c_int32, [_eng, _f32p, C.c_uint32, C.c_uint32, C.c_int64, _u64p, _f32p,
                                        C.c_uint32, _u32p]),
    "wax_vs_search_device": (C.c_int32, [_eng, C.c_void_p, C.c_uint32, C.c_int64, C.c_uint64, C.c_void_p,
                                         C.c_void_p]),
    "wax_vs_search_batch_device": (C.c_int32, [_eng, C.c_void_p, C.c_uint32, C.c_int64, C.c_uint64, C.c_void_p,
                                               C.c_void_p]),
    "wax_vs_merge_candidates_device": (C.c_int32, [_eng, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.c_void_p, C.c_void_p]),
    "wax_vs_shard_open": (C.c_int32, [_eng, C.c_int32, C.c_int32, C.c_uint64, _u8p]),
    "wax_vs_shard_connect": (C.c_int32, [_eng, _u8p, C.c_int32]),
    "wax_vs_shard_close": (C.c_int32, [_eng]),
    "wax_vs_shard_search": (C.c_int32, [_eng, _f32p, C.c_uint32, C.c_int64, _u64p, _f32p, C.c_uint32, _u32p]),
    "wax_vs_shard_search_device": (C.c_int32, [_eng, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "wax_vs_serialized_length": (C.c_int32, [_eng, _u64p]),
    "wax_vs_serialize": (C.c_int32, [_eng, _u8p, C.c_uint64, _u64p]),
    "wax_vs_deserialize": (C.c_int32, [_eng, _u8p, C.c_uint64]),
    "wax_vs_last_error": (C.c_char_p, []),
    "wax_vs_debug_pool_stats": (C.c_int32, [_eng, _u64p, _u64p]),
    "wax_vs_debug_fill_synthetic": (C.c_int32, [_eng, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32]),
    "wax_vs_debug_read_rows": (C.c_int32, [_eng, C.c_uint64, C.c_uint64, _f32p]),
    "wax_vs_debug_time_search": (C.c_int32, [_eng, C.c_uint32, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32,
                                             _f32p, _u64p]),
    "wax_vs_debug_time_shard_search": (C.c_int32, [_eng, C.c_uint32, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32,
                                                   _f32p, _u64p]),
    "wax_vs_debug_transfer_probe": (C.c_int32, [_eng, C.c_uint64, _f32p]),
    "wax_vs_debug_phase_trace": (C.c_int32, [_eng, C.c_int64, C.c_uint32, _f32p]),
    "wax_vs_debug_batch_stats": (C.c_int32, [_eng, _u64p, _u64p]),
    "wax_vs_debug_counter": (C.c_int32, [_eng, C.c_char_p, _u64p]),
    "wax_vs_debug_time_search_batch": (C.c_int32, [_eng, C.c_uint32, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32,
                                                   _f32p, _u64p, _u32p]),
    "wax_vs_debug_stream_read": (C.c_int32, [_eng, C.c_uint32, _f32p, _u64p]),
    "wax_vs_debug_set_option": (C.c_int32, [_eng, C.c_char_p, C.c_int64]),
    "wax_vs_version": (C.c_char_p, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load libwaxvs_cuda.so; raise loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension is the only implementation of the vector scan. "
                "Build it with `python -m wax_b200.build` (or __graft_entry__.build()).")
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if a declared symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().wax_vs_last_error().decode("utf-8", "replace")
