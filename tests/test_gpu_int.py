"""GPU parity: integer path (FoR + FastLanes bit-packing) through the C ABI vs. the CPU oracle.

Known answers transcribed from the reference:
  README.md:43-88 / src/core/README.md:17-104          quick-start (BASELINE config 1)
  src/core/src/liquid_array/primitive_array.rs:751-982  round trips, nulls, extreme ranges, filters
  src/core/src/liquid_array/raw/bit_pack_array.rs:355-436  1024/500/2048 values, widths 8..32
  src/core/src/cache/transcode.rs:301-436               dtype matrix at 8192 rows
"""
import datetime as dt
import zlib

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle.liquid_oracle import OracleIntArray
from tests.util import assert_arrays_equal, assert_masks_equal, random_selection

pytestmark = pytest.mark.gpu


def _expr(op, value):
    from liquid_cache_b200 import BinaryExpr, Column, LiquidExpr, Literal

    return LiquidExpr.new_unchecked(BinaryExpr(Column("liquid_predicate_col", 0), op, Literal(value)))


def test_quick_start_config1(cache):
    """README quick-start: UInt64 [10..15]; selection T F T F T F -> [10,12,14]; col > 12 -> [F,F,F,T,T,T]."""
    from liquid_cache_b200 import EntryID

    arr = pa.array([10, 11, 12, 13, 14, 15], pa.uint64())
    eid = EntryID(424242)
    cache.insert(eid, arr).run()
    assert cache.is_cached(eid)
    assert cache.get(eid).read().equals(arr)
    sel = pa.array([True, False, True, False, True, False])
    assert cache.get(eid).with_selection(sel).read().to_pylist() == [10, 12, 14]
    mask = cache.eval_predicate(eid, _expr(">", 12)).read()
    assert mask.to_pylist() == [False, False, False, True, True, True]
    mask = cache.eval_predicate(eid, _expr(">", 12)).with_selection(sel).read()
    assert mask.to_pylist() == [False, False, True]
    assert cache.get(EntryID(999999999)).read() is None
    assert cache.eval_predicate(EntryID(999999999), _expr(">", 12)).read() is None


TYPES = [
    (pa.int8(), -128, 127), (pa.int16(), -2000, 31000), (pa.int32(), -(2**31), 2**31 - 1),
    (pa.int64(), -(2**63), 2**63 - 1), (pa.uint8(), 0, 255), (pa.uint16(), 100, 160),
    (pa.uint32(), 0, 2**32 - 1), (pa.uint64(), 0, 2**64 - 1), (pa.int64(), 1373832014, 1373832014 + 86400),
    (pa.int32(), 7, 7),
]


@pytest.mark.parametrize("typ,lo,hi", TYPES)
@pytest.mark.parametrize("n", [1, 500, 1024, 2048, 8192, 10000])
@pytest.mark.parametrize("null_p", [0.0, 0.2])
def test_round_trip_filter_predicate(cache, typ, lo, hi, n, null_p):
    rng = np.random.default_rng(zlib.crc32(repr((str(typ), lo, hi, n, null_p)).encode()))
    if hi < 2**63:
        vals = rng.integers(lo, hi, size=n, endpoint=True, dtype=np.int64)
    else:
        vals = rng.integers(lo, hi, size=n, endpoint=True, dtype=np.uint64)
    mask = rng.random(n) < null_p if null_p else None
    arr = pa.array(vals, type=typ, mask=mask)
    oracle = OracleIntArray.from_arrow(arr)
    liquid = cache.transcode(arr)
    assert liquid.len() == n
    assert_arrays_equal(liquid.to_arrow_array(), oracle.to_arrow(), "to_arrow")
    for p in (1.0, 0.5, 0.03, 0.0):
        sel = random_selection(rng, n, p)
        assert_arrays_equal(liquid.filter(sel), oracle.filter(sel), f"filter p={p}")
        lit = int(vals[rng.integers(0, n)])
        for op in ("=", "!=", "<", "<=", ">", ">="):
            got = liquid.try_eval_predicate(_expr(op, lit), sel)
            want = oracle.try_eval_predicate(op, lit, sel)
            assert_masks_equal(got, want, f"{typ} n={n} {op} {lit} p={p}")
    # literals outside the value window / type range fold to constants
    sel = random_selection(rng, n, 0.7)
    for lit in (lo - 1 if lo > -(2**63) else lo, hi + 1 if hi < 2**64 - 1 else hi):
        if not (-(2**63) <= lit < 2**64):
            continue
        np_t = typ.to_pandas_dtype()
        if lit < np.iinfo(np_t).min or lit > np.iinfo(np_t).max:
            continue  # pyarrow cannot build the scalar; the packed-domain planner handles it (see below)
        for op in ("=", "!=", "<", ">="):
            assert_masks_equal(liquid.try_eval_predicate(_expr(op, lit), sel), oracle.try_eval_predicate(op, lit, sel),
                               f"edge {op} {lit}")


def test_literal_outside_type_range(cache):
    arr = pa.array([1, 2, None, 4], pa.int16())
    liquid = cache.transcode(arr)
    sel = pa.array([True] * 4)
    assert liquid.try_eval_predicate(_expr("<", 100000), sel).to_pylist() == [True, True, None, True]
    assert liquid.try_eval_predicate(_expr("=", 100000), sel).to_pylist() == [False, False, None, False]
    assert liquid.try_eval_predicate(_expr(">", -100000), sel).to_pylist() == [True, True, None, True]
    assert liquid.try_eval_predicate(_expr("!=", -100000), sel).to_pylist() == [True, True, None, True]


def test_all_null_and_single_value_and_zero_reference(cache):
    """primitive_array.rs:804-925: all-null arrays, single value, zero reference value."""
    for arr in (
        pa.array([None, None, None], pa.int32()),
        pa.array([42], pa.int32()),
        pa.array([0, 1, 2, 3, 4], pa.int32()),
        pa.array([-128, 127], pa.int8()),
        pa.array([-(2**63), 2**63 - 1, None, 0], pa.int64()),
        pa.array([], pa.int32()),
    ):
        liquid = cache.transcode(arr)
        assert_arrays_equal(liquid.to_arrow_array(), arr, str(arr.type))
        if len(arr):
            sel = pa.array([i % 2 == 0 for i in range(len(arr))])
            assert_arrays_equal(liquid.filter(sel), arr.filter(sel), "filter")
            got = liquid.try_eval_predicate(_expr(">=", 0), sel)
            want = OracleIntArray.from_arrow(arr).try_eval_predicate(">=", 0, sel)
            assert_masks_equal(got, want, "all-null / single")


def test_dates_and_timestamps(cache):
    d32 = pa.array([dt.date(1992, 1, 2), None, dt.date(1998, 12, 1), dt.date(1994, 6, 15)], pa.date32())
    liquid = cache.transcode(d32)
    assert_arrays_equal(liquid.to_arrow_array(), d32, "date32")
    sel = pa.array([True, True, True, True])
    got = liquid.try_eval_predicate(_expr(">=", dt.date(1994, 1, 1)), sel)
    assert got.to_pylist() == [False, None, True, True]
    for unit in ("s", "ms", "us", "ns"):
        ts = pa.array([i * 1000 for i in range(3000)], pa.timestamp(unit))
        liquid = cache.transcode(ts)
        assert_arrays_equal(liquid.to_arrow_array(), ts, unit)
    d64 = pa.array([86400000 * i for i in range(100)], pa.date64())
    assert_arrays_equal(cache.transcode(d64).to_arrow_array(), d64, "date64")


def test_date64_predicate_compares_milliseconds(cache):
    """A `datetime.date` literal against a Date64 column is compared in milliseconds (the column's physical unit), checked
    against Arrow's own comparison."""
    days = [dt.date(2019, 12, 30) + dt.timedelta(days=i) for i in range(400)]
    d64 = pa.array([None if i % 11 == 0 else d for i, d in enumerate(days)], pa.date64())
    liquid = cache.transcode(d64)
    sel = pa.array([i % 3 != 0 for i in range(len(days))])
    lit = dt.date(2020, 1, 1)
    for op, fn in ((">=", pc.greater_equal), ("<", pc.less), ("=", pc.equal), ("!=", pc.not_equal)):
        got = liquid.try_eval_predicate(_expr(op, lit), sel)
        want = fn(d64.filter(sel), pa.scalar(lit, pa.date64()))
        assert_masks_equal(got, want, f"date64 {op}")


def test_unsupported_types_are_declined(cache):
    """transcode.rs:420-436: Boolean (and tz timestamps) stay Arrow -> Err(array)."""
    from liquid_cache_b200 import _native as N

    for arr in (pa.array([True, False]), pa.array([1, 2], pa.timestamp("us", tz="UTC")), pa.array([1.5, 2.5], pa.float16()),
                pa.array([[1], [2]], pa.list_(pa.int32())), pa.array(["a"], pa.large_string())):
        with pytest.raises(N.UnsupportedType):
            cache.transcode(arr)


def test_sliced_input_with_offset(cache):
    """bit_pack_array.rs:512-531: a sliced null bitmap (non-zero offset) must be re-aligned."""
    base = pa.array([None if i % 3 == 0 else i for i in range(5000)], pa.int32())
    arr = base.slice(13, 3001)
    liquid = cache.transcode(arr)
    assert_arrays_equal(liquid.to_arrow_array(), arr, "sliced")


def test_batched_many(cache):
    rng = np.random.default_rng(7)
    arrays = [pa.array(rng.integers(0, 1 << w, size=8192, dtype=np.int64), pa.int64()) for w in (1, 5, 17, 33, 40)]
    liquids = [cache.transcode(a) for a in arrays]
    handles = np.array([l.handle for l in liquids], dtype=np.uint64)
    rows = np.array([8192] * len(arrays), dtype=np.uint64)
    sels = [np.packbits(rng.random(8192) < 0.3, bitorder="little") for _ in arrays]
    sels = [np.concatenate([s, np.zeros(8, np.uint8)]) for s in sels]
    expr = _expr(">", 3)
    vals, valid, offs, out_len, out_nulls, true_counts = cache.eval_predicate_many(handles, rows, expr, pa.int64(), sels)
    for i, a in enumerate(arrays):
        sel = pa.array(np.unpackbits(sels[i], bitorder="little")[:8192].astype(bool))
        want = OracleIntArray.from_arrow(a).try_eval_predicate(">", 3, sel)
        k = int(out_len[i])
        got = np.unpackbits(vals[int(offs[i]):int(offs[i]) + (k + 7) // 8], bitorder="little")[:k].astype(bool)
        assert got.tolist() == want.to_pylist()
        assert int(true_counts[i]) == sum(1 for x in want.to_pylist() if x)
    concat = cache.to_arrow_many(handles, sels)
    want = pa.concat_arrays([a.filter(pa.array(np.unpackbits(s, bitorder="little")[:8192].astype(bool))) for a, s in zip(arrays, sels)])
    assert_arrays_equal(concat, want, "to_arrow_many")


@pytest.mark.parametrize("density", [0.0, 0.0003, 0.02, 0.5])
def test_batched_many_sparse_and_dense_selections(cache, density):
    """Batched calls over enough entries to take the sparse selection upload ({word index, word} pairs, zero fill and
    scatter on the device) when the selections are nearly empty, and the dense copy otherwise; a few entries are fully
    selected or get no selection at all."""
    rng = np.random.default_rng(int(density * 1e4) + 3)
    n_entries, rows_n = 24, 8192
    arrays = [pa.array(rng.integers(-1000, 1000, size=rows_n), pa.int32(), mask=rng.random(rows_n) < 0.05) for _ in range(n_entries)]
    liquids = [cache.transcode(a) for a in arrays]
    handles = np.array([l.handle for l in liquids], dtype=np.uint64)
    rows = np.array([rows_n] * n_entries, dtype=np.uint64)
    bools = [rng.random(rows_n) < density for _ in range(n_entries)]
    bools[3] = np.ones(rows_n, dtype=bool)
    sels = [np.concatenate([np.packbits(b, bitorder="little"), np.zeros(8, np.uint8)]) for b in bools]
    sels[7] = None  # no selection = all rows
    bools[7] = np.ones(rows_n, dtype=bool)
    vals, valid, offs, out_len, out_nulls, true_counts = cache.eval_predicate_many(handles, rows, _expr("<", 0), pa.int32(), sels)
    for i, a in enumerate(arrays):
        want = OracleIntArray.from_arrow(a).try_eval_predicate("<", 0, pa.array(bools[i]))
        k = int(out_len[i])
        assert k == int(bools[i].sum()) and int(out_nulls[i]) == want.null_count
        got = np.unpackbits(vals[int(offs[i]):int(offs[i]) + (k + 7) // 8], bitorder="little")[:k].astype(bool)
        assert got.tolist() == [bool(x) if x is not None else False for x in want.to_pylist()]  # null -> false bit
        assert int(true_counts[i]) == sum(1 for x in want.to_pylist() if x)
    concat = cache.to_arrow_many(handles, sels)
    want = pa.concat_arrays([a.filter(pa.array(b)) for a, b in zip(arrays, bools)])
    assert_arrays_equal(concat, want, "to_arrow_many")


@pytest.mark.parametrize("with_nulls", [False, True])
def test_batched_many_into_pinned_buffers(cache, with_nulls):
    """Page-locked result buffers take the direct paths: nearly-empty masks come back as {word index, word} pairs
    (zero fill + scatter on the host), dense ones through the chunked copy; the sequence sparse, dense, sparse, dense
    also walks the per-list hint both ways."""
    import torch

    rng = np.random.default_rng(99)
    n_entries, rows_n = 72, 8192
    arrays = [pa.array(rng.integers(-1000, 1000, size=rows_n), pa.int32(),
                       mask=(rng.random(rows_n) < 0.05) if (with_nulls and i % 3 == 0) else None) for i in range(n_entries)]
    liquids = [cache.transcode(a) for a in arrays]
    handles = np.array([l.handle for l in liquids], dtype=np.uint64)
    rows = np.array([rows_n] * n_entries, dtype=np.uint64)
    sizes = (((rows + 7) // 8 + 15) // 16) * 16
    offs = np.zeros(n_entries, dtype=np.uint64)
    np.cumsum(sizes[:-1], out=offs[1:])
    total = int(sizes.sum())
    pin = lambda nb: torch.full((nb,), 0xA5, dtype=torch.uint8).pin_memory().numpy()  # noqa: E731  (garbage to overwrite)
    for lit in (-995, 0, -995, 5000):  # sparse, dense, sparse, dense (all true)
        bufs = (pin(total), pin(total), offs, np.zeros(n_entries, np.uint64), np.zeros(n_entries, np.uint64), np.zeros(n_entries, np.uint64))
        pred = _expr("<", lit).to_native(pa.int32())
        vals, valid, _, out_len, out_nulls, true_counts = cache._eval_many_native(handles, rows, pred, None, bufs)
        for i, a in enumerate(arrays):
            want = pc.less(a, pa.scalar(lit, pa.int32())).to_pylist()  # == the oracle's answer for an all-rows selection
            assert int(out_len[i]) == rows_n and int(out_nulls[i]) == a.null_count
            got = np.unpackbits(vals[int(offs[i]):int(offs[i]) + rows_n // 8], bitorder="little").astype(bool)
            assert got.tolist() == [bool(x) if x is not None else False for x in want], f"lit {lit} entry {i}"
            assert int(true_counts[i]) == sum(1 for x in want if x)
            if a.null_count:
                gv = np.unpackbits(valid[int(offs[i]):int(offs[i]) + rows_n // 8], bitorder="little").astype(bool)
                assert gv.tolist() == [x is not None for x in want]


def test_and_then(cache):
    """reader/utils/boolean_selection.rs:233-256 known answer + random equivalence (datafusion/src/utils.rs:317-408)."""
    from oracle.liquid_oracle import boolean_buffer_and_then

    left = pa.array([bool(int(c)) for c in "001011010101"])
    right = pa.array([bool(int(c)) for c in "001101"])
    assert "".join("1" if x else "0" for x in cache.and_then(left, right).to_pylist()) == "000001010001"
    rng = np.random.default_rng(11)
    for n in (1, 63, 64, 65, 1000, 8192, 20000):
        l = rng.random(n) < 0.4
        r = rng.random(int(l.sum())) < 0.5
        got = cache.and_then(pa.array(l), pa.array(r))
        assert got.to_pylist() == boolean_buffer_and_then(pa.array(l), pa.array(r)).to_pylist()
