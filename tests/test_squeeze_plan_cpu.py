"""The decision taken on the host before a predicate runs on a squeezed entry (csrc/squeeze_plan.h), together with the
kernel's planner (csrc/int_plan.cuh), on the CPU: for every operator and literal the pair must reproduce the restated
reference arrays (oracle/liquid_oracle.py OracleClampedArray / OracleQuantizedArray) exactly —
  * the oracle raises NeedsBacking  <=>  the host is in doubt AND its probe (run through the planner) finds a valid row;
  * otherwise the predicate run through the planner on the codes is the oracle's mask.
This is the whole correctness argument of squeezed predicates short of the device's thread indexing."""
import ctypes as C
import os
import subprocess

import numpy as np
import pyarrow as pa
import pytest

from oracle import liquid_oracle as O
from tests.test_int_plan_cpu import OPS, plan_eval
from tests.test_int_plan_cpu import lib as plan_lib  # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def doubt_lib():
    out = os.path.join(ROOT, "build", "tests", "libsqueeze_plan_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-Wno-unused-function", "-shared", "-fPIC", f"-I{ROOT}",
                        "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "squeeze_plan_host.cc"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    l = C.CDLL(out)
    l.sp_doubt.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.c_int64, C.c_uint64]
    return l


@pytest.mark.parametrize("typ,base,span", [(pa.int32(), -1_000_000, 1 << 16), (pa.uint32(), 1_000_000, 1 << 16), (pa.int64(), -(2**40), 1 << 20),
                                           (pa.uint16(), 100, 1 << 12), (pa.int64(), -(2**62), 2**62), (pa.uint64(), 2**63, 1 << 20), (pa.int8(), -128, 255),
                                           (pa.int16(), -20000, 1 << 15)], ids=str)
@pytest.mark.parametrize("policy", ["clamp", "quantize"])
def test_host_decision_plus_kernel_planner_equal_the_reference_arrays(plan_lib, doubt_lib, typ, base, span, policy):
    rng = np.random.default_rng(span % 997 + typ.bit_width)
    n = 3000
    vals = [base + int(d) for d in rng.integers(0, span, size=n, endpoint=True)]
    nulls = rng.random(n) < 0.1
    mn = min(v for v, m in zip(vals, nulls) if not m)
    arr = pa.array(np.array([mn if m else v for v, m in zip(vals, nulls)], dtype=typ.to_pandas_dtype()), typ, mask=nulls)
    sq, _image = O.squeeze_int(O.OracleIntArray.from_arrow(arr), O.OracleSqueezeIo(), "PredicateColumn", policy)
    codes_all, valid_all = sq._codes(), ~nulls
    info = np.iinfo(typ.to_pandas_dtype())
    signed = info.min < 0
    kind = 1 if policy == "clamp" else 2
    bw = getattr(sq, "bucket_width", 0)
    ref_raw = sq.reference & ((1 << typ.bit_width) - 1)
    last = (1 << sq.bit_width) - 1
    mx = max(v for v, m in zip(vals, nulls) if not m)
    lits = {mn - 1, mn, mn + 1, mx, mx + 1, mn + last - 1, mn + last, mn + last + 1, info.min, info.max, mn + (bw or 1), mn + (bw or 1) - 1, mn + 5 * (bw or 1),
            mn + 5 * (bw or 1) + 1, mn + 6 * (bw or 1) - 1, info.max + 1, info.min - 1} | {int(v) for v in rng.choice(vals, 10)}
    seen = {"codes": 0, "backing": 0, "doubt cleared by the probe": 0}
    for sel_p in (1.0, 0.5, 0.02):
        sel = rng.random(n) < sel_p
        codes, valid = codes_all[sel], valid_all[sel]
        for k in sorted(x for x in lits if -(2**63) <= x < 2**64):
            lit_kind = 0 if k < 2**63 else 1
            for op in OPS:
                d = doubt_lib.sp_doubt(typ.bit_width, sq.bit_width, int(signed), ref_raw, kind, bw, OPS.index(op), lit_kind, k if lit_kind == 0 else 0,
                                       k if lit_kind == 1 else 0)
                try:
                    want = sq._eval_inner(op, k, codes, valid)
                    oracle_backing = False
                except O.NeedsBacking:
                    want, oracle_backing = None, True
                if want is None and not oracle_backing:  # Ok(None): the literal is outside the column's type
                    assert d == 3, (policy, str(typ), op, k)
                    continue
                assert d != 3, (policy, str(typ), op, k)
                found = False
                if d == 1:
                    found = bool((plan_eval(plan_lib, typ.bit_width, sq.bit_width, signed, sq.reference, "=", 0, codes, squeeze_kind=kind, bucket_width=bw, lit_kind=8) & valid).any())
                elif d == 2:
                    found = bool((plan_eval(plan_lib, typ.bit_width, sq.bit_width, signed, sq.reference, "=", k, codes, squeeze_kind=kind, bucket_width=bw) & valid).any())
                assert found == oracle_backing, (policy, str(typ), op, k, sel_p, d)
                if oracle_backing:
                    seen["backing"] += 1
                    continue
                seen["codes"] += 1
                seen["doubt cleared by the probe"] += d in (1, 2)
                got = plan_eval(plan_lib, typ.bit_width, sq.bit_width, signed, sq.reference, op, k, codes, squeeze_kind=kind, bucket_width=bw)
                w = np.asarray(want.to_numpy(zero_copy_only=False))
                assert np.array_equal(got[valid], np.array([bool(x) for x in w[valid]])), (policy, str(typ), op, k, sel_p)
    assert seen["codes"] and seen["backing"], seen
