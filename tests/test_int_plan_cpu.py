"""k_int_scan's predicate planner (liquid_cache_b200/csrc/int_plan.cuh) on the CPU: the (comparison kind, threshold) it
derives from an entry header and `col <op> literal`, applied to packed values exactly as the scan loops apply it, must
give the plain comparison for entries of every integer type, literal kind and position of the literal relative to the
entry's value window (below it, inside, above, outside the type, across the signed / unsigned boundary)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["=", "!=", "<", "<=", ">", ">="]
PY = {"=": lambda a, k: a == k, "!=": lambda a, k: a != k, "<": lambda a, k: a < k, "<=": lambda a, k: a <= k, ">": lambda a, k: a > k, ">=": lambda a, k: a >= k}


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "tests", "libint_plan_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-Wno-unused-function", "-shared", "-fPIC", f"-I{ROOT}",
                        "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "int_plan_host.cc"), "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    l = C.CDLL(out)
    l.ip_eval.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32, C.c_int32, C.c_int64, C.c_uint64,
                          C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64)]
    return l


def plan_eval(lib, tbits, width, signed, reference, op, k, packed, squeeze_kind=0, bucket_width=0, lit_kind=None):
    out = np.zeros(len(packed), dtype=np.uint8)
    thr = C.c_uint64(0)
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    if lit_kind is None:
        lit_kind = 0 if -(2**63) <= k < 2**63 else 1
    lit_i = k if lit_kind == 0 else 0
    lit_u = k if lit_kind == 1 else 0
    lib.ip_eval(tbits, width, int(signed), reference & ((1 << tbits) - 1), squeeze_kind, bucket_width, OPS.index(op), lit_kind, lit_i, lit_u,
                packed.ctypes.data, len(packed), out.ctypes.data, C.byref(thr))
    return out.astype(bool)


@pytest.mark.parametrize("np_dt", [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64], ids=lambda d: np.dtype(d).name)
def test_full_entries_compare_like_the_values(lib, np_dt):
    info = np.iinfo(np_dt)
    tbits, signed = info.bits, info.min < 0
    rng = np.random.default_rng(tbits + signed)
    for width in sorted({1, 3, tbits // 2, tbits - 1, tbits}):
        span = (1 << width) - 1
        for reference in {info.min, 0 if not signed else -5, max(info.min, info.max - span - 7), info.max - span}:
            if reference + span > info.max or reference < info.min:
                continue
            offs = np.unique(np.concatenate([rng.integers(0, span, size=200, endpoint=True, dtype=np.uint64), np.array([0, span, span // 2], dtype=np.uint64)]))
            values = [reference + int(o) for o in offs]
            lits = {reference - 1, reference, reference + 1, reference + span - 1, reference + span, reference + span + 1, info.min, info.max, 0,
                    reference + span // 2, -1, 2**63, 2**64 - 1, -(2**63)}
            for k in lits:
                if not (-(2**63) <= k < 2**64):
                    continue
                for op in OPS:
                    got = plan_eval(lib, tbits, width, signed, reference, op, k, offs)
                    want = np.array([PY[op](v, k) for v in values])
                    assert np.array_equal(got, want), (np.dtype(np_dt).name, width, reference, op, k)
