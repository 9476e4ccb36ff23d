"""k_int_scan's predicate planner (liquid_cache_b200/csrc/int_plan.cuh) on the CPU: the (comparison kind, threshold) it
derives from an entry header and `col <op> literal`, applied to packed values exactly as the scan loops apply it, must
give (1) the plain comparison for full entries of every integer type, literal kind and position of the literal relative
to the entry's value window, and (2) for squeezed entries the answer of the restated reference arrays
(oracle/liquid_oracle.py OracleClampedArray / OracleQuantizedArray) whenever those decide from the codes — which is all the
host lets through (squeeze_host.cc) — plus the two probes that find the rows they cannot decide."""
import ctypes as C
import os
import subprocess

import numpy as np
import pyarrow as pa
import pytest

from oracle import liquid_oracle as O

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


@pytest.mark.parametrize("typ,base,span", [(pa.int32(), -1_000_000, 1 << 16), (pa.uint32(), 1_000_000, 1 << 16), (pa.int64(), -(2**40), 1 << 20),
                                           (pa.uint16(), 100, 1 << 12), (pa.int64(), -(2**62), 2**62), (pa.uint64(), 2**63, 1 << 20), (pa.int8(), -128, 255)], ids=str)
@pytest.mark.parametrize("policy", ["clamp", "quantize"])
def test_squeezed_entries_answer_like_the_reference_arrays(lib, typ, base, span, policy):
    rng = np.random.default_rng(span % 1000 + typ.bit_width)
    vals = [base + int(d) for d in rng.integers(0, span, size=3000, endpoint=True)]
    arr = pa.array(vals, typ)
    full = O.OracleIntArray.from_arrow(arr)
    sq, _image = O.squeeze_int(full, O.OracleSqueezeIo(), "PredicateColumn", policy)
    codes = sq._codes()
    valid = np.ones(len(codes), dtype=bool)
    info = np.iinfo(typ.to_pandas_dtype())
    signed = info.min < 0
    kind = 1 if policy == "clamp" else 2
    bw = getattr(sq, "bucket_width", 0)
    mn, mx = min(vals), max(vals)
    last = (1 << sq.bit_width) - 1
    lits = {mn - 1, mn, mn + 1, mx, mx + 1, mn + last - 1, mn + last, mn + last + 1, info.min, info.max, mn + (bw or 1), mn + (bw or 1) - 1,
            mn + 5 * (bw or 1), mn + 5 * (bw or 1) + 1, mn + 6 * (bw or 1) - 1} | {int(v) for v in rng.choice(vals, 8)}
    decided = doubted = 0
    for k in sorted(x for x in lits if info.min <= x <= info.max):
        for op in OPS:
            got = plan_eval(lib, typ.bit_width, sq.bit_width, signed, sq.reference, op, k, codes, squeeze_kind=kind, bucket_width=bw)
            try:
                want = sq._eval_inner(op, k, codes, valid)
            except O.NeedsBacking:
                # the host never lets such a call through; what it runs first is the probe, which must find the rows in doubt
                doubted += 1
                if policy == "clamp":
                    probe = plan_eval(lib, typ.bit_width, sq.bit_width, signed, sq.reference, "=", 0, codes, squeeze_kind=kind, bucket_width=bw, lit_kind=8)
                    assert np.array_equal(probe, codes == last) and probe.any()
                else:
                    probe = plan_eval(lib, typ.bit_width, sq.bit_width, signed, sq.reference, "=", k, codes, squeeze_kind=kind, bucket_width=bw)
                    q = (k - sq.reference) // bw
                    assert np.array_equal(probe, codes.astype(np.uint64) == np.uint64(q)) and probe.any()
                continue
            decided += 1
            assert np.array_equal(got, np.asarray(want.to_numpy(zero_copy_only=False), dtype=bool)), (policy, str(typ), op, k)
    assert decided > 0 and doubted > 0
