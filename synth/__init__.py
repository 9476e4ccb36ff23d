"""Deterministic synthetic columns shaped like ClickBench `hits` / TPC-H `lineitem`.

Input data only: this package contains no liquid-cache logic. One call yields one 8192-row Arrow batch, so the
100 M-row workloads are streamed entry by entry and never exist on the host at once.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

URL_POOL = 65536       # per-entry unique pool; Zipf(1.2) draws leave ~1 800 distinct values per 8192 rows
URL_INJECT_P = 4e-5    # fraction of rows carrying the token "google" (paper Table 1: Q20 selectivity < 0.01 %)
SEED_URL, SEED_INT, SEED_TPCH = 20, 30, 40


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liblc_synth.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", _HERE], check=True)
        l = C.CDLL(so)
        l.lcs_url_entry.restype = C.c_uint64
        l.lcs_url_entry.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p,
                                    C.c_uint64]
        l.lcs_int_entry.restype = None
        l.lcs_int_entry.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        l.lcs_init.restype = None
        l.lcs_init.argtypes = [C.c_uint32]
        _LIB = l
    return _LIB


def url_entry(entry_idx: int, rows: int = 8192, seed: int = SEED_URL, pool: int = URL_POOL,
              inject_p: float = URL_INJECT_P) -> pa.Array:
    """One batch of the URL column as a Utf8 array (config 2)."""
    l = lib()
    l.lcs_init(pool)
    off = np.empty(rows + 1, dtype=np.int32)
    data = np.empty(rows * 512, dtype=np.uint8)
    n = l.lcs_url_entry(seed, entry_idx, rows, pool, inject_p, off.ctypes.data, data.ctypes.data, len(data))
    return pa.Array.from_buffers(pa.string(), rows, [None, pa.py_buffer(off), pa.py_buffer(data[:n].copy())])


_INT_KINDS = {"EventTime": (0, np.int64, pa.int64()), "UserID": (1, np.int64, pa.int64()),
              "l_shipdate": (2, np.int32, pa.date32()), "AdvEngineID": (3, np.int16, pa.int16())}


def int_entry(column: str, entry_idx: int, rows: int = 8192, seed: int = SEED_INT) -> pa.Array:
    """One batch of an integer-like column (configs 3 and 4)."""
    kind, np_t, pa_t = _INT_KINDS[column]
    out = np.empty(rows, dtype=np_t)
    lib().lcs_int_entry(seed, entry_idx, rows, kind, out.ctypes.data)
    arr = pa.array(out)
    return arr.cast(pa_t) if arr.type != pa_t else arr
