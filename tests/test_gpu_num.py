"""GPU parity: ALP floats and u64 decimals through the C ABI vs. the CPU oracle.

Known answers transcribed from the reference (paths relative to /root/reference/src/core/src):
  liquid_array/float_array.rs:1058-1181   round trips and filters, Float32 / Float64
  cache/transcode.rs:330-349              Float32 / Float64 `0..8192` through transcode
  liquid_array/decimal_array.rs:643-695   Decimal128 round trip, `>= 1.00` with a null
Bars: decoded floats BIT-exact against the oracle (which reproduces the reference's one lossy case, -0.0 -> +0.0);
masks and validity exact; the ALP layout the device chose (exponents, bit width, reference, patch list) equal to the
oracle's restatement of get_best_exponents / encode_arrow_array.
"""
import decimal
import struct
import zlib

import numpy as np
import pyarrow as pa
import pytest

from oracle.liquid_oracle import OracleDecimalArray, OracleFloatArray
from tests.util import assert_arrays_equal, assert_float_bits_equal, assert_masks_equal, random_selection

pytestmark = pytest.mark.gpu

INT_HDR = struct.Struct("<I4B I I Q 5I 4I")  # csrc/entry_layout.h IntHeader up to patch_val_off


def _expr(op, value):
    from liquid_cache_b200 import BinaryExpr, Column, LiquidExpr, Literal

    return LiquidExpr.new_unchecked(BinaryExpr(Column("liquid_predicate_col", 0), op, Literal(value)))


def parse_float_image(img: bytes):
    (magic, phys, tbits, bit_width, has_nulls, n, n_chunks, reference, validity_off, packed_off, blob_bytes, null_count,
     is_signed, alp_ef, n_patches, patch_idx_off, patch_val_off) = INT_HDR.unpack_from(img, 0)
    assert magic == 0x3149514C and blob_bytes == len(img)
    it, ft = (np.int32, np.float32) if tbits == 32 else (np.int64, np.float64)
    ref = int(np.array([reference & ((1 << tbits) - 1)], dtype=np.uint64).astype(np.uint32 if tbits == 32 else np.uint64).view(it)[0])
    return dict(n=n, bit_width=bit_width, reference=ref, e=alp_ef & 0xFF, f=(alp_ef >> 8) & 0xFF, null_count=null_count,
                patch_idx=np.frombuffer(img, dtype=np.uint32, count=n_patches, offset=patch_idx_off) if n_patches else np.zeros(0, np.uint32),
                patch_val=np.frombuffer(img, dtype=ft, count=n_patches, offset=patch_val_off) if n_patches else np.zeros(0, ft))


def make_floats(rng, kind: str, n: int, np_dt):
    if kind == "integers":
        return rng.integers(-5000, 5000, size=n).astype(np_dt)
    if kind == "prices":  # two decimals: the case ALP is made for
        return np.round(rng.uniform(0, 1000, size=n), 2).astype(np_dt)
    if kind == "normal":  # nothing is representable: every row is a patch
        return rng.standard_normal(n).astype(np_dt)
    if kind == "mixed":   # mostly prices, a few rows that need patches, the special values
        x = np.round(rng.uniform(-100, 100, size=n), 1).astype(np_dt)
        k = max(1, n // 50)
        x[rng.integers(0, n, size=k)] = rng.standard_normal(k).astype(np_dt)
        specials = np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, np.finfo(np_dt).max, np.finfo(np_dt).tiny], dtype=np_dt)
        idx = rng.integers(0, n, size=min(n, len(specials)))
        x[idx] = specials[: len(idx)]
        return x
    raise ValueError(kind)


@pytest.mark.parametrize("typ", [pa.float32(), pa.float64()], ids=str)
@pytest.mark.parametrize("kind", ["integers", "prices", "normal", "mixed"])
@pytest.mark.parametrize("n", [1, 500, 2048, 8192, 10000])
@pytest.mark.parametrize("null_p", [0.0, 0.2])
def test_float_round_trip_filter_predicate(cache, typ, kind, n, null_p):
    from liquid_cache_b200 import _native as N

    np_dt = np.dtype(typ.to_pandas_dtype())
    rng = np.random.default_rng(zlib.crc32(repr((str(typ), kind, n, null_p)).encode()))
    vals = make_floats(rng, kind, n, np_dt)
    mask = rng.random(n) < null_p if null_p else None
    arr = pa.array(vals, type=typ, mask=mask)
    oracle = OracleFloatArray.from_arrow(arr)
    liquid = cache.transcode(arr)
    assert liquid.len() == n and liquid.data_type() == N.LIQUID_FLOAT
    assert liquid.original_arrow_data_type() == typ
    assert_float_bits_equal(liquid.to_arrow_array(), oracle.to_arrow(), "to_arrow")
    # the layout the device chose = the oracle's restatement of get_best_exponents / encode_arrow_array
    img = parse_float_image(liquid.entry_image())
    if oracle.bit_width is None:
        assert img["bit_width"] == 0 and img["null_count"] == n
    else:
        assert (img["e"], img["f"]) == (oracle.e, oracle.f), "exponents"
        assert img["bit_width"] == oracle.bit_width and img["reference"] == oracle.reference
        assert img["patch_idx"].tolist() == oracle.patch_indices.tolist()
        it = np.uint32 if np_dt.itemsize == 4 else np.uint64
        assert img["patch_val"].view(it).tolist() == oracle.patch_values.view(it).tolist()
    lits = [float(vals[rng.integers(0, n)]), 0.0, -0.0, float("nan"), float("inf"), 12.5]
    ops = ("=", "!=", "<", "<=", ">", ">=")
    for pi, p in enumerate((1.0, 0.5, 0.03, 0.0)):
        sel = random_selection(rng, n, p)
        assert_float_bits_equal(liquid.filter(sel), oracle.filter(sel), f"filter p={p}")
        for li, lit in enumerate(lits if p == 1.0 else lits[:2]):
            lit = float(np_dt.type(lit))
            for op in (ops if li == 0 else ops[(li + pi) % 3::3]):  # every op on a value of the column, two per special
                got = liquid.try_eval_predicate(_expr(op, lit), sel)
                want = oracle.try_eval_predicate(op, lit, sel)
                assert_masks_equal(got, want, f"{typ} {kind} n={n} {op} {lit} p={p}")


def test_float_reference_known_answers(cache):
    """float_array.rs:1058-1181 and transcode.rs:330-349 through the CUDA path."""
    for typ in (pa.float32(), pa.float64()):
        for values in ([-1.0, 1.0, 0.0], [-1.0, 1.0, 0.0, None], [None, None, None, None], []):
            arr = pa.array(values, typ)
            assert_arrays_equal(cache.transcode(arr).to_arrow_array(), arr, f"{typ} {values}")
        arr = pa.array(np.arange(8192).astype(typ.to_pandas_dtype()), typ)
        liquid = cache.transcode(arr)
        assert_float_bits_equal(liquid.to_arrow_array(), arr, "0..8192")
        assert liquid.get_array_memory_size() < arr.nbytes  # float_array.rs:1183-1210: ALP must pay off here
    liquid = cache.transcode(pa.array([1.0, 2.1, 3.2, None, 5.5], pa.float32()))
    got = liquid.filter(pa.array([True, False, True, False, True]))
    assert_arrays_equal(got, pa.array([1.0, 3.2, 5.5], pa.float32()), "filter basic")
    liquid = cache.transcode(pa.array([None] * 4, pa.float32()))
    assert_arrays_equal(liquid.filter(pa.array([True, False, False, True])), pa.array([None, None], pa.float32()), "all nulls")
    liquid = cache.transcode(pa.array([1.0, 2.1, 3.3], pa.float32()))
    assert len(liquid.filter(pa.array([False] * 3))) == 0


def test_float_minus_zero_and_nan_follow_the_reference(cache):
    nan_payload = np.array([0x7FF8000000000123], dtype=np.uint64).view(np.float64)[0]
    x = np.array([0.0, -0.0, nan_payload, np.inf, -np.inf, 2.5, 1e300, np.nan])
    arr = pa.array(x)
    liquid = cache.transcode(arr)
    want = x.copy()
    want[1] = 0.0  # the reference's ALP round trip turns -0.0 into +0.0 (oracle: test_alp_quirks_follow_the_reference)
    got = np.asarray(liquid.to_arrow_array().to_numpy(zero_copy_only=False))
    assert got.view(np.uint64).tolist() == want.view(np.uint64).tolist()
    sel = pa.array([True] * len(x))
    # total order: NaN equals NaN only bit for bit (the payload NaN is a different, larger value); the stored -0.0 is +0.0
    assert liquid.try_eval_predicate(_expr("=", float("nan")), sel).to_pylist() == [False, False, False, False, False, False, False, True]
    assert liquid.try_eval_predicate(_expr(">", float("nan")), sel).to_pylist() == [False, False, True, False, False, False, False, False]
    assert liquid.try_eval_predicate(_expr("<", 0.0), sel).to_pylist() == [False, False, False, False, True, False, False, False]
    assert liquid.try_eval_predicate(_expr(">=", float("inf")), sel).to_pylist() == [False, False, True, True, False, False, False, True]
    oracle = OracleFloatArray.from_arrow(arr)
    for op in ("=", "!=", "<", "<=", ">", ">="):
        for lit in (float("nan"), -0.0, 0.0, 2.5, float("-inf")):
            assert_masks_equal(liquid.try_eval_predicate(_expr(op, lit), sel), oracle.try_eval_predicate(op, lit, sel), f"{op} {lit}")


def test_float_sliced_input_with_offset(cache):
    base = pa.array([None if i % 3 == 0 else i * 0.25 for i in range(5000)], pa.float64())
    arr = base.slice(13, 3001)
    assert_float_bits_equal(cache.transcode(arr).to_arrow_array(), OracleFloatArray.from_arrow(arr).to_arrow(), "sliced")


@pytest.mark.parametrize("typ", [pa.float32(), pa.float64()], ids=str)
def test_float_batched_calls_and_scan_pipeline(cache, typ):
    """lc_eval_predicate_many / lc_to_arrow_many over several float entries with selections, and the device-resident
    conjunct pipeline (lc_scan_filter = decode + compare + AND into the running selection)."""
    np_dt = np.dtype(typ.to_pandas_dtype())
    rng = np.random.default_rng(5 + np_dt.itemsize)
    rows_n, n_entries = 8192, 6
    arrays = []
    for i in range(n_entries):
        kind = ["prices", "mixed", "integers"][i % 3]
        vals = make_floats(rng, kind, rows_n, np_dt)
        arrays.append(pa.array(vals, typ, mask=(rng.random(rows_n) < 0.1) if i % 2 else None))
    liquids = [cache.transcode(a) for a in arrays]
    oracles = [OracleFloatArray.from_arrow(a) for a in arrays]
    handles = np.array([l.handle for l in liquids], dtype=np.uint64)
    rows = np.full(n_entries, rows_n, dtype=np.uint64)
    bools = [rng.random(rows_n) < p for p in (0.3, 0.001, 1.0, 0.0, 0.5, 0.9)]
    sels = [np.concatenate([np.packbits(b, bitorder="little"), np.zeros(8, np.uint8)]) for b in bools]
    sels[2] = None
    lit = float(np_dt.type(12.5))
    vals, valid, offs, out_len, out_nulls, true_counts = cache.eval_predicate_many(handles, rows, _expr(">", lit), typ, sels)
    for i, o in enumerate(oracles):
        want = o.try_eval_predicate(">", lit, pa.array(bools[i]))
        k = int(out_len[i])
        assert k == int(bools[i].sum()) and int(out_nulls[i]) == want.null_count
        got = np.unpackbits(vals[int(offs[i]):int(offs[i]) + (k + 7) // 8], bitorder="little")[:k].astype(bool)
        assert got.tolist() == [bool(x) if x is not None else False for x in want.to_pylist()], f"entry {i}"
        gv = np.unpackbits(valid[int(offs[i]):int(offs[i]) + (k + 7) // 8], bitorder="little")[:k].astype(bool)
        assert gv.tolist() == [x is not None for x in want.to_pylist()], f"validity {i}"
        assert int(true_counts[i]) == sum(1 for x in want.to_pylist() if x)
    concat = cache.to_arrow_many(handles, sels)
    want = pa.concat_arrays([o.filter(pa.array(b)) for o, b in zip(oracles, bools)])
    assert_float_bits_equal(concat, want, "to_arrow_many")
    # device pipeline: two float conjuncts, then get-with-selection of the survivors
    with cache.scan([rows_n] * n_entries) as scan:
        scan.filter(handles, _expr(">=", float(np_dt.type(-20.0))), typ)
        scan.filter(handles, _expr("<", float(np_dt.type(40.0))), typ)
        counts, total = scan.counts()
        got = scan.read(handles)
        want_parts = []
        for b, o in enumerate(oracles):
            all_rows = pa.array(np.ones(rows_n, dtype=bool))
            m1 = o.try_eval_predicate(">=", float(np_dt.type(-20.0)), all_rows).fill_null(False)
            m2 = o.try_eval_predicate("<", float(np_dt.type(40.0)), all_rows).fill_null(False)
            sel = pa.array(np.asarray(m1.to_numpy(zero_copy_only=False), dtype=bool) & np.asarray(m2.to_numpy(zero_copy_only=False), dtype=bool))
            assert scan.selection(b).to_pylist() == sel.to_pylist(), f"selection of batch {b}"
            assert int(counts[b]) == sum(sel.to_pylist())
            want_parts.append(o.filter(sel))
        assert total == sum(len(p) for p in want_parts)
        assert_float_bits_equal(got, pa.concat_arrays(want_parts), "scan read")


# ---- decimals ------------------------------------------------------------------------------------------------
def _dec_array(ints, typ, mask=None):
    scale = typ.scale
    with decimal.localcontext() as cx:
        cx.prec = 100
        vals = [None if (mask is not None and mask[i]) else decimal.Decimal(int(v)).scaleb(-scale) for i, v in enumerate(ints)]
    return pa.array(vals, type=typ)


def test_decimal_reference_known_answers(cache):
    """decimal_array.rs:643-695."""
    from liquid_cache_b200 import _native as N

    d = _dec_array([100, 0, 250], pa.decimal128(10, 2), mask=[False, True, False])
    liquid = cache.transcode(d)
    assert liquid.data_type() == N.LIQUID_DECIMAL and liquid.original_arrow_data_type() == d.type
    assert_arrays_equal(liquid.to_arrow_array(), d, "decimal_u64_roundtrip")
    d = _dec_array([100, 200, 0, 300], pa.decimal128(10, 2), mask=[False, False, True, False])
    got = cache.transcode(d).try_eval_predicate(_expr(">=", decimal.Decimal("1.00")), pa.array([True] * 4))
    assert got.to_pylist() == [True, True, None, True]


@pytest.mark.parametrize("typ", [pa.decimal128(15, 2), pa.decimal128(38, 0), pa.decimal256(50, 4)], ids=str)
@pytest.mark.parametrize("n", [1, 1000, 8192, 10000])
@pytest.mark.parametrize("null_p", [0.0, 0.2])
def test_decimal_round_trip_filter_predicate(cache, typ, n, null_p):
    rng = np.random.default_rng(zlib.crc32(repr((str(typ), n, null_p)).encode()))
    hi = 2**64 - 1 if typ.precision >= 20 else 10**typ.precision - 1
    lo = hi - 100000 if typ.precision >= 38 else 0  # exercise the top of u64 too
    ints = rng.integers(lo, hi, size=n, dtype=np.uint64, endpoint=True)
    mask = (rng.random(n) < null_p) if null_p else None
    arr = _dec_array(ints, typ, mask)
    oracle = OracleDecimalArray.from_arrow(arr)
    liquid = cache.transcode(arr)
    assert_arrays_equal(liquid.to_arrow_array(), oracle.to_arrow(), "to_arrow")
    scale = typ.scale
    lits = [int(ints[rng.integers(0, n)]), lo, hi, -5, 2**64 + 7, 0]
    ops = ("=", "!=", "<", "<=", ">", ">=")
    for pi, p in enumerate((1.0, 0.4, 0.0)):
        sel = random_selection(rng, n, p)
        assert_arrays_equal(liquid.filter(sel), oracle.filter(sel), f"filter p={p}")
        for li, u in enumerate(lits if p == 1.0 else lits[:2]):
            if abs(u) >= 10**typ.precision:
                continue  # not a value of this type: DataFusion could not have produced the literal
            lit = decimal.Decimal(u).scaleb(-scale)
            for op in (ops if li == 0 else ops[(li + pi) % 3::3]):
                got = liquid.try_eval_predicate(_expr(op, lit), sel)
                want = oracle.try_eval_predicate(op, pa.scalar(lit, typ), sel)
                assert_masks_equal(got, want, f"{typ} n={n} {op} {u} p={p}")


def test_decimal_outside_u64_takes_the_fixed_length_form(cache):
    """transcode.rs:118-131: such arrays become LiquidFixedLenByteArray (tests/test_gpu_zy_fixed_len.py has the rest).
    Null slots do not count (decimal_array.rs:127-132)."""
    for arr in (_dec_array([5, -1, 7], pa.decimal128(10, 2)), _dec_array([2**64], pa.decimal128(38, 0)),
                _dec_array([1, 2**70], pa.decimal256(60, 0))):
        liquid = cache.transcode(arr)
        assert liquid.data_type() == 3
        assert_arrays_equal(liquid.to_arrow_array(), arr, "decimal outside u64")
    # a null slot whose payload is negative is fine
    arr = pa.array([decimal.Decimal("-1.00"), decimal.Decimal("2.00")], pa.decimal128(10, 2))
    arr = pa.Array.from_buffers(arr.type, 2, [pa.py_buffer(bytes([0b10])), arr.buffers()[1]], null_count=1)
    assert_arrays_equal(cache.transcode(arr).to_arrow_array(), arr, "negative payload under a null")


def test_decimal_scan_pipeline_and_batched(cache):
    typ = pa.decimal128(15, 2)  # TPC-H l_discount / l_quantity shapes (q6: l_discount between 0.05 and 0.07, l_quantity < 24)
    rng = np.random.default_rng(77)
    rows_n, n_entries = 8192, 5
    disc = [_dec_array(rng.integers(0, 11, size=rows_n), typ, (rng.random(rows_n) < 0.05) if i == 2 else None) for i in range(n_entries)]
    qty = [_dec_array(rng.integers(100, 5001, size=rows_n), typ) for _ in range(n_entries)]
    ld, lq = [cache.transcode(a) for a in disc], [cache.transcode(a) for a in qty]
    od, oq = [OracleDecimalArray.from_arrow(a) for a in disc], [OracleDecimalArray.from_arrow(a) for a in qty]
    hd = np.array([l.handle for l in ld], dtype=np.uint64)
    hq = np.array([l.handle for l in lq], dtype=np.uint64)
    d = decimal.Decimal
    with cache.scan([rows_n] * n_entries) as scan:
        scan.filter(hd, _expr(">=", d("0.05")), typ)
        scan.filter(hd, _expr("<=", d("0.07")), typ)
        scan.filter(hq, _expr("<", d("24.00")), typ)
        counts, total = scan.counts()
        got = scan.read(hq)
        parts = []
        for b in range(n_entries):
            all_rows = pa.array(np.ones(rows_n, dtype=bool))
            m = [od[b].try_eval_predicate(">=", pa.scalar(d("0.05"), typ), all_rows), od[b].try_eval_predicate("<=", pa.scalar(d("0.07"), typ), all_rows),
                 oq[b].try_eval_predicate("<", pa.scalar(d("24.00"), typ), all_rows)]
            sel = np.ones(rows_n, dtype=bool)
            for x in m:
                sel &= np.asarray(x.fill_null(False).to_numpy(zero_copy_only=False), dtype=bool)
            assert scan.selection(b).to_pylist() == sel.tolist(), f"batch {b}"
            assert int(counts[b]) == int(sel.sum())
            parts.append(oq[b].filter(pa.array(sel)))
        assert_arrays_equal(got, pa.concat_arrays(parts), "scan read of decimals")
    sels = [np.concatenate([np.packbits(rng.random(rows_n) < 0.2, bitorder="little"), np.zeros(8, np.uint8)]) for _ in range(n_entries)]
    concat = cache.to_arrow_many(hd, sels)
    want = pa.concat_arrays([o.filter(pa.array(np.unpackbits(s, bitorder="little")[:rows_n].astype(bool))) for o, s in zip(od, sels)])
    assert_arrays_equal(concat, want, "to_arrow_many decimals")
