"""Pins the float (ALP) and decimal parts of the CPU oracle against the reference's own tests, and the device's
power-of-ten tables against the oracle's. No GPU needed.

Known answers transcribed from (paths relative to /root/reference/src/core/src):
  liquid_array/float_array.rs:1058-1124   round trips: basic / with nulls / all nulls / empty, Float32 and Float64
  liquid_array/float_array.rs:1126-1181   filter: basic, all nulls, empty result
  liquid_array/float_array.rs:1183-1210   2000 x `i as f32/f64` must come out smaller than the Arrow array
  cache/transcode.rs:330-349              Float32 / Float64 `0..8192` round trip through transcode
  liquid_array/decimal_array.rs:643-654   Decimal128(10,2) [100, NULL, 250] round trip
  liquid_array/decimal_array.rs:670-695   Decimal128(10,2) [100, 200, NULL, 300] >= 100 -> [T, T, NULL, T]
"""
import decimal
import os
import re

import numpy as np
import pyarrow as pa
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import liquid_oracle as O
from oracle.liquid_oracle import OracleDecimalArray, OracleFloatArray, float_total_order_compare, transcode
from tests.util import assert_float_bits_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOAT_TYPES = [pa.float32(), pa.float64()]


@pytest.mark.parametrize("typ", FLOAT_TYPES, ids=str)
@pytest.mark.parametrize("values", [[-1.0, 1.0, 0.0], [-1.0, 1.0, 0.0, None], [None, None, None, None], []],
                         ids=["basic", "with_nones", "all_nones", "empty"])
def test_float_roundtrip_known_answers(typ, values):
    arr = pa.array(values, typ)
    o = OracleFloatArray.from_arrow(arr)
    assert o.to_arrow().equals(arr)
    if values and all(v is None for v in values):
        assert o.bit_width is None and (o.e, o.f) == (0, 0)  # float_array.rs:620-630


def test_float_filter_known_answers():
    o = OracleFloatArray.from_arrow(pa.array([1.0, 2.1, 3.2, None, 5.5], pa.float32()))
    got = o.filter(pa.array([True, False, True, False, True]))
    assert got.equals(pa.array([1.0, 3.2, 5.5], pa.float32()))
    o = OracleFloatArray.from_arrow(pa.array([None] * 4, pa.float32()))
    assert o.filter(pa.array([True, False, False, True])).equals(pa.array([None, None], pa.float32()))
    o = OracleFloatArray.from_arrow(pa.array([1.0, 2.1, 3.3], pa.float32()))
    assert len(o.filter(pa.array([False] * 3))) == 0


@pytest.mark.parametrize("typ", FLOAT_TYPES, ids=str)
def test_float_transcode_and_compression(typ):
    np_dt = typ.to_pandas_dtype()
    arr = pa.array(np.arange(8192).astype(np_dt), typ)  # transcode.rs:330-349
    o = transcode(arr)
    assert isinstance(o, OracleFloatArray) and o.to_arrow().equals(arr)
    assert len(o.patch_indices) == 0
    small = OracleFloatArray.from_arrow(pa.array(np.arange(2000).astype(np_dt), typ))  # float_array.rs:1183-1210
    packed_bytes = ((2000 + 1023) // 1024) * 128 * small.bit_width
    assert packed_bytes + 8 * len(small.patch_indices) < 2000 * np.dtype(np_dt).itemsize


def test_alp_arithmetic_constants():
    # SWEET = 2^(FRACTIONAL_BITS) + 2^(FRACTIONAL_BITS-1); MAX_EXPONENT; table ends (float_array.rs:127-222)
    assert float(O._ALP_SWEET[32]) == 12582912.0 and float(O._ALP_SWEET[64]) == 6755399441055744.0
    f10, if10 = O._ALP[64]
    assert len(f10) == 24 and f10[23] == 1e23 and if10[23] == 1e-23 and f10[0] == 1.0
    f10, if10 = O._ALP[32]
    assert len(f10) == 11 and float(f10[10]) == 1e10 and if10[1] == np.float32(0.1)
    # fast_round: round-half-even through the sweet spot, saturating cast
    enc = O.alp_encode_values(np.array([0.5, 1.5, 2.5, -0.5, np.nan, np.inf, -np.inf, 1e300]), 1, 0)
    assert enc.tolist()[:4] == [5, 15, 25, -5] and enc[4] == 0 and enc[5] == np.iinfo(np.int64).max and enc[6] == np.iinfo(np.int64).min


def test_device_tables_match_the_oracle():
    """liquid_cache_b200/csrc/alp_math.cuh spells the power-of-ten tables as hex floats; they must be the very
    numbers the oracle computes from exact rationals."""
    text = open(os.path.join(ROOT, "liquid_cache_b200", "csrc", "alp_math.cuh")).read()

    def table(name):
        body = re.search(name + r"\[\d+\] = \{(.*?)\};", text, re.S).group(1)
        return [float.fromhex(x.strip().rstrip("f")) for x in body.split(",")]

    assert table("kAlpF10d") == O._ALP[64][0] and table("kAlpIF10d") == O._ALP[64][1]
    assert table("kAlpF10f") == [float(x) for x in O._ALP[32][0]] and table("kAlpIF10f") == [float(x) for x in O._ALP[32][1]]


def test_alp_quirks_follow_the_reference():
    """`decoded.eq(&v)` is IEEE equality: NaN and the infinities are patched and come back bit for bit; -0.0 equals the
    decoded +0.0, is NOT patched and therefore comes back as +0.0 (float_array.rs:633-640)."""
    nan_payload = np.array([0x7FF8000000000123], dtype=np.uint64).view(np.float64)[0]
    x = np.array([0.0, -0.0, nan_payload, np.inf, -np.inf, 2.5, 1e300])
    o = OracleFloatArray.from_arrow(pa.array(x))
    got = np.asarray(o.to_arrow().to_numpy(zero_copy_only=False))
    want = x.copy()
    want[1] = 0.0
    assert got.view(np.uint64).tolist() == want.view(np.uint64).tolist()
    assert {2, 3, 4}.issubset(set(o.patch_indices.tolist())) and 1 not in o.patch_indices.tolist()


def test_float_compare_is_total_order():
    a = pa.array([0.0, -0.0, float("nan"), float("inf"), None, -1.5], pa.float64())
    assert float_total_order_compare(a, "=", float("nan")).to_pylist() == [False, False, True, False, None, False]
    assert float_total_order_compare(a, "<", 0.0).to_pylist() == [False, True, False, False, None, True]
    assert float_total_order_compare(a, ">", float("inf")).to_pylist() == [False, False, True, False, None, False]
    assert float_total_order_compare(a, "!=", -0.0).to_pylist() == [True, False, True, True, None, True]
    b = pa.array([np.float32(0.1), np.float32(0.3)], pa.float32())
    assert float_total_order_compare(b, "=", np.float32(0.1)).to_pylist() == [True, False]


@settings(max_examples=150, deadline=None)
@given(st.lists(st.one_of(st.none(), st.floats(allow_nan=True, allow_infinity=True, width=64)), min_size=0, max_size=300),
       st.sampled_from([32, 64]))
def test_float_roundtrip_property(values, bits):
    typ, np_dt = (pa.float32(), np.float32) if bits == 32 else (pa.float64(), np.float64)
    with np.errstate(over="ignore"):
        vals = [None if v is None else float(np_dt(v)) for v in values]
    arr = pa.array(vals, typ)
    got = OracleFloatArray.from_arrow(arr).to_arrow()
    # lossless except for the reference's -0.0 -> +0.0
    want = pa.array([None if v is None else (0.0 if v == 0 else v) for v in vals], typ)
    assert_float_bits_equal(got, want, "alp round trip")


def test_best_exponents_sampling_rules():
    # > 1024 rows: every (n // 1024)-th slot, nulls dropped (float_array.rs:719-727)
    n = 5000
    x = np.round(np.linspace(0, 999, n), 1)
    mask = np.zeros(n, dtype=bool)
    mask[::4] = True  # every sampled slot (step 4) is null -> empty sample -> first pair (1, 0)
    assert OracleFloatArray.best_exponents(pa.array(x, mask=mask)) == (1, 0)
    e, f = OracleFloatArray.best_exponents(pa.array(x))
    assert 0 <= f < e < 18
    o = OracleFloatArray.from_arrow(pa.array(x))
    assert o.to_arrow().equals(pa.array(x))


def test_decimal_known_answers():
    d = pa.array([decimal.Decimal("1.00"), None, decimal.Decimal("2.50")], pa.decimal128(10, 2))  # 100, NULL, 250
    o = transcode(d)
    assert isinstance(o, OracleDecimalArray) and o.to_arrow().equals(d)
    assert o.ints.reference == 100 and o.ints.bit_width == 8  # 250 - 100 = 150 -> 8 bits
    d = pa.array([decimal.Decimal("1.00"), decimal.Decimal("2.00"), None, decimal.Decimal("3.00")], pa.decimal128(10, 2))
    lit = pa.scalar(decimal.Decimal("1.00"), pa.decimal128(10, 2))
    assert transcode(d).try_eval_predicate(">=", lit, pa.array([True] * 4)).to_pylist() == [True, True, None, True]
    # values outside u64 (negative, or >= 2^64) are not LiquidDecimalArray material (decimal_array.rs:127-132): they take
    # the LiquidFixedLenByteArray form (transcode.rs:118-153, tests/test_oracle_fixed_len.py)
    from oracle.liquid_oracle import OracleFixedLenByteArray
    for outside in (pa.array([decimal.Decimal("-0.01")], pa.decimal128(10, 2)), pa.array([decimal.Decimal(2**64)], pa.decimal128(38, 0))):
        assert isinstance(transcode(outside), OracleFixedLenByteArray) and transcode(outside).to_arrow().equals(outside)
    big = pa.array([decimal.Decimal(2**64 - 1), None], pa.decimal256(50, 0))
    assert transcode(big).to_arrow().equals(big)


# ---- LQDA (the reference's serialized form) -------------------------------------------------------------------
def test_lqda_known_answers_and_round_trips():
    """ipc.rs:308-418: header fields of a serialized Int32 array, round trips incl. all-null / no-null / single / empty /
    sparse-null arrays; the same for floats (float_array.rs:1058-1124 run every case through to_bytes/from_bytes) and
    decimals (decimal_array.rs:656-668)."""
    arr = pa.array([10, 20, 30, None, 50], pa.int32())
    b = O.to_bytes(O.OracleIntArray.from_arrow(arr))
    assert b[0:4] == (0x4C514441).to_bytes(4, "little") and int.from_bytes(b[4:6], "little") == 1
    assert int.from_bytes(b[6:8], "little") == 1 and int.from_bytes(b[8:10], "little") == 2  # Integer, Int32
    assert len(b) > 100 and len(b) == 24 + 24 + 768  # header+ref pad 8 | bit-pack header + 1 null byte pad 8 | 1024 x 6 bits
    assert O.read_from_bytes(b).to_arrow().equals(arr)
    cases = [pa.array([None] * 1000, pa.int32()), pa.array(list(range(1000)), pa.int32()), pa.array([42], pa.int32()),
             pa.array([], pa.int32()), pa.array([None if i in (1000, 5000, 9000) else i for i in range(10000)], pa.int32()),
             pa.array([1, 2, 3], pa.timestamp("us")), pa.array([-5, 2**62], pa.int64()), pa.array([1, None, 2**64 - 1], pa.uint64()),
             pa.array([8036, 10556, None], pa.date32())]
    for a in cases:
        assert O.read_from_bytes(O.to_bytes(O.OracleIntArray.from_arrow(a))).to_arrow().equals(a), a.type
    for typ in FLOAT_TYPES:
        for values in ([-1.0, 1.0, 0.0], [-1.0, 1.0, 0.0, None], [None] * 4, [], [0.1, float("nan"), 1e30, 2.5]):
            a = pa.array(values, typ)
            o = OracleFloatArray.from_arrow(a)
            r = O.read_from_bytes(O.to_bytes(o))
            assert (r.e, r.f, r.bit_width) == (o.e, o.f, o.bit_width)
            assert_float_bits_equal(r.to_arrow(), o.to_arrow(), f"{typ} {values}")
    d = pa.array([decimal.Decimal("12.345"), decimal.Decimal("67.890")], pa.decimal128(12, 3))
    b = O.to_bytes(OracleDecimalArray.from_arrow(d))
    assert int.from_bytes(b[6:8], "little") == 6 and int.from_bytes(b[8:10], "little") == 7 and list(b[16:19]) == [0, 12, 3]
    assert O.read_from_bytes(b).to_arrow().equals(d)


@settings(max_examples=120, deadline=None)
@given(st.sampled_from([pa.int8(), pa.int16(), pa.int32(), pa.int64(), pa.uint8(), pa.uint16(), pa.uint32(), pa.uint64(), pa.date32(),
                        pa.timestamp("ms")]),
       st.lists(st.one_of(st.none(), st.integers(min_value=-(2**63), max_value=2**64 - 1)), min_size=0, max_size=400), st.data())
def test_integer_oracle_properties(typ, raw, data):
    """For any integer-like column: the LQDA image reads back as the array (ipc.rs round-trip tests generalised), and
    the oracle's predicate on a random selection equals Arrow's compare on the filtered array."""
    import pyarrow.compute as pc

    store = pa.int32() if pa.types.is_date32(typ) else (pa.int64() if pa.types.is_timestamp(typ) else typ)
    info = np.iinfo(store.to_pandas_dtype())
    span = int(info.max) - int(info.min) + 1
    vals = [None if v is None else (int(v) - int(info.min)) % span + int(info.min) for v in raw]
    arr = pa.array(vals, store).cast(typ)
    o = O.OracleIntArray.from_arrow(arr)
    assert o.to_arrow().equals(arr)
    assert O.read_from_bytes(O.to_bytes(o)).to_arrow().equals(arr)
    if len(arr):
        sel = pa.array(data.draw(st.lists(st.booleans(), min_size=len(arr), max_size=len(arr))))
        lit_i = data.draw(st.sampled_from([v for v in vals if v is not None] or [0]))
        lit = pa.scalar(lit_i, store).cast(typ)
        for op, fn in (("=", pc.equal), ("<", pc.less), (">=", pc.greater_equal), ("!=", pc.not_equal)):
            want = fn(arr.filter(sel), lit)
            assert o.try_eval_predicate(op, lit, sel).to_pylist() == want.to_pylist()


@settings(max_examples=60, deadline=None)
@given(st.lists(st.one_of(st.none(), st.integers(min_value=0, max_value=2**64 - 1)), min_size=0, max_size=200),
       st.sampled_from([pa.decimal128(38, 0), pa.decimal128(20, 3), pa.decimal256(50, 4)]))
def test_decimal_oracle_properties(raw, typ):
    with decimal.localcontext() as cx:
        cx.prec = 100
        vals = [None if v is None else decimal.Decimal(v).scaleb(-typ.scale) for v in raw]
        fits = all(v is None or abs(v.scaleb(typ.scale)) < 10**typ.precision for v in vals)
    if not fits:
        return
    arr = pa.array(vals, typ)
    o = OracleDecimalArray.from_arrow(arr)
    assert o is not None and o.to_arrow().equals(arr)
    assert O.read_from_bytes(O.to_bytes(o)).to_arrow().equals(arr)
