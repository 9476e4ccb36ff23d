"""The squeezed integer arrays of the CPU oracle against the reference's own tests
(liquid_array/hybrid_primitive_array.rs:870-1291, all under /root/reference/src/core/src): a column narrower than 8 bits
does not squeeze; the full bytes hydrate back to the original; rows below the clamp boundary materialize without IO;
for every comparison operator the literals the reference lists as resolvable are answered from the half-width codes with
zero reads, the ones it lists as unresolvable read the backing bytes — and both give the answer of the plain comparison.
The reference draws its arrays from rand's StdRng; here the same shapes (length, base, range 2^16, null share) come from
NumPy seeds, so the checks are the reference's assertions on different draws."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import liquid_oracle as O

HINT = "PredicateColumn"
OPS = ["=", "!=", "<", "<=", ">", ">="]


def make_array(typ, n, base_min, rng_range, null_prob, seed):
    rng = np.random.default_rng(seed)
    vals = [base_min + int(d) for d in rng.integers(0, rng_range, size=n, endpoint=True)]
    nulls = rng.random(n) < null_prob
    return pa.array([None if m else int(v) for v, m in zip(vals, nulls)], typ)  # null slots hold 0 in the buffer


def boundary_of(arr):
    """compute_boundary_i32 / _u32 (:848-868): min + ((1 << (bit_width(range) / 2)) - 1)"""
    mn, mx = pc.min_max(arr)["min"].as_py(), pc.min_max(arr)["max"].as_py()
    half = O.get_bit_width(mx - mn) // 2
    return mn + ((1 << half) - 1 if half else 0)


def expected_for(arr, sel, op, k):
    return O._PC_CMP[op](pc.filter(arr, sel), pa.scalar(k, arr.type))


def squeeze(arr, policy):
    io = O.OracleSqueezeIo()
    got = O.squeeze_int(O.OracleIntArray.from_arrow(arr), io, HINT, policy)
    assert got is not None, "squeezable"
    hybrid, full = got
    io.set_bytes(full)
    return hybrid, full, io


def test_clamp_unsqueezable_small_range():
    arr = make_array(pa.int32(), 64, 10_000, 100, 0.1, 0x5171)  # range < 128 -> bit width < 8
    assert O.squeeze_int(O.OracleIntArray.from_arrow(arr), O.OracleSqueezeIo(), HINT, "clamp") is None
    wide = make_array(pa.int32(), 64, 10_000, 1 << 12, 0.1, 0x5171)
    assert O.squeeze_int(O.OracleIntArray.from_arrow(wide), O.OracleSqueezeIo(), None, "clamp") is None  # no hint, no squeeze
    dates = pa.array(list(range(8036, 10556)), pa.int32()).cast(pa.date32())
    assert O.squeeze_int(O.OracleIntArray.from_arrow(dates), O.OracleSqueezeIo(), HINT, "clamp") is None  # wants a date field hint


def test_clamp_squeeze_full_read_roundtrip_i32():
    arr = make_array(pa.int32(), 128, -50_000, 1 << 16, 0.1, 0x5172)
    liq = O.OracleIntArray.from_arrow(arr)
    hybrid, full, io = squeeze(arr, "clamp")
    assert full == O.to_bytes(liq)
    recovered = O.read_from_bytes(full)
    assert recovered.to_arrow().equals(arr) and O.to_bytes(recovered) == full
    assert hybrid.bit_width == liq.bit_width // 2 and len(hybrid) == len(arr)
    boundary = boundary_of(arr)
    keep = pa.array([True if v is None else v < boundary for v in arr.to_pylist()])
    io.reset_reads()
    got = hybrid.filter(keep)
    assert io.reads == 0
    assert got.equals(pc.filter(arr, keep))
    io.reset_reads()
    everything = hybrid.filter(pa.array([True] * len(arr)))  # rows at the sentinel: the full bytes are read
    assert io.reads > 0 and everything.equals(arr)
    assert len(hybrid.filter(pa.array([False] * len(arr)))) == 0


@pytest.mark.parametrize("typ,n,base,null_p,seed", [(pa.int32(), 200, -1_000_000, 0.2, 0x5173), (pa.uint32(), 180, 1_000_000, 0.15, 0x5174),
                                                   (pa.int64(), 3000, -(2**40), 0.1, 7), (pa.uint16(), 2500, 100, 0.0, 8),
                                                   (pa.int8(), 300, -128, 0.3, 9), (pa.uint64(), 1500, 2**63, 0.05, 10)])
def test_clamp_predicate_eval_resolvable_and_unresolvable(typ, n, base, null_p, seed):
    span = {8: 255, 16: 1 << 12}.get(typ.bit_width, 1 << 16)
    arr = make_array(typ, n, base, span, null_p, seed)
    hybrid, _full, io = squeeze(arr, "clamp")
    boundary = boundary_of(arr)
    sel = pa.array(np.random.default_rng(seed + 1).random(n) < 0.5)
    resolvable = [("=", boundary - 1), ("!=", boundary - 1), ("<", boundary), ("<=", boundary - 1), (">", boundary - 1), (">=", boundary)]
    for op, k in resolvable:
        io.reset_reads()
        got = hybrid.try_eval_predicate(op, k, sel)
        assert io.reads == 0, (op, k)
        assert got.equals(expected_for(arr, sel, op, k)), (op, k)
    unresolvable = [("=", boundary), ("!=", boundary), ("<", boundary + 1), ("<=", boundary), (">", boundary + 1), (">=", boundary + 1)]
    for op, k in unresolvable:
        io.reset_reads()
        got = hybrid.try_eval_predicate(op, k, sel)
        assert io.reads > 0, (op, k)  # the draw holds selected rows at or past the boundary
        assert got.equals(expected_for(arr, sel, op, k)), (op, k)


@pytest.mark.parametrize("typ,n,base,seed", [(pa.uint32(), 200, 1_000_000, 0x5184), (pa.int32(), 220, -1_000_000, 0x5185),
                                             (pa.int64(), 2100, -(2**50), 11), (pa.uint16(), 1200, 7, 12)])
def test_quantize_predicate_eval_resolvable_and_unresolvable(typ, n, base, seed):
    span = (1 << 12) if typ.bit_width == 16 else (1 << 16)
    arr = make_array(typ, n, base, span, 0.2, seed)
    hybrid, _full, io = squeeze(arr, "quantize")
    mn = pc.min_max(arr)["min"].as_py()
    lo = max(mn - 1, 0) if pa.types.is_unsigned_integer(typ) else mn - 1  # min.saturating_sub(1)
    sel = pa.array([True] * n)
    for op, k, const in [("=", lo, False), ("!=", lo, True), ("<", mn, False), ("<=", lo, False), (">", lo, True), (">=", mn, True)]:
        io.reset_reads()
        got = hybrid.try_eval_predicate(op, k, sel)
        want = pa.array([None if v is None else const for v in arr.to_pylist()], pa.bool_())
        assert io.reads == 0, (op, k)
        assert got.equals(want) and got.equals(expected_for(arr, sel, op, k)), (op, k)
    k_present = next(v for v in arr.to_pylist() if v is not None)
    io.reset_reads()
    got = hybrid.try_eval_predicate("=", k_present, sel)
    assert io.reads > 0
    assert got.equals(expected_for(arr, sel, "=", k_present))


def test_quantize_to_arrow_reads_the_backing():
    arr = make_array(pa.uint32(), 64, 1000, 1 << 12, 0.0, 0x5186)
    hybrid, _full, io = squeeze(arr, "quantize")
    io.reset_reads()
    assert hybrid.to_arrow().equals(arr) and io.reads > 0


@pytest.mark.parametrize("policy", ["clamp", "quantize"])
@pytest.mark.parametrize("typ", [pa.int16(), pa.int32(), pa.uint32(), pa.int64(), pa.uint64()], ids=str)
def test_squeezed_answers_always_equal_the_plain_comparison(policy, typ):
    """What the reference's two code paths guarantee together: whichever way a call goes, the mask is the one of the full
    array — over literals below, inside and above the value range, bucket edges, and literals that do not fit the type."""
    rng = np.random.default_rng(typ.bit_width * 4 + (policy == "clamp") + 2 * pa.types.is_unsigned_integer(typ))
    info = np.iinfo(typ.to_pandas_dtype())
    base = int(rng.integers(info.min // 2, info.max // 2, dtype=np.int64)) if typ.bit_width < 64 else (info.min // 2 + 12345)
    span = (1 << 13) if typ.bit_width == 16 else (1 << 20)
    arr = make_array(typ, 4000, base, span, 0.1, 99)
    hybrid, _full, io = squeeze(arr, policy)
    mn, mx = pc.min_max(arr)["min"].as_py(), pc.min_max(arr)["max"].as_py()
    bw = getattr(hybrid, "bucket_width", 1)
    lits = {mn - 1, mn, mn + 1, mx - 1, mx, mx + 1, boundary_of(arr), boundary_of(arr) - 1, boundary_of(arr) + 1, mn + bw, mn + bw - 1,
            mn + 5 * bw, mn + 5 * bw + bw - 1, mn + 5 * bw + 1, info.min, info.max}
    lits |= {int(v) for v in rng.choice([v for v in arr.to_pylist() if v is not None], 6)}
    sel = pa.array(rng.random(len(arr)) < 0.6)
    saved = reads = 0
    for k in sorted(x for x in lits if info.min <= x <= info.max):
        for op in OPS:
            io.reset_reads()
            got = hybrid.try_eval_predicate(op, k, sel)
            assert got.equals(expected_for(arr, sel, op, k)), (op, k)
            reads += io.reads > 0
            saved += io.reads == 0
    assert saved > 0 and reads > 0  # both paths were taken
    # a literal outside the native type: Ok(None) -> the filtered array is materialized and compared as 64-bit
    if typ.bit_width < 64:
        got = hybrid.try_eval_predicate("<", pa.scalar(info.max + 10, pa.int64()), sel)
        assert got.equals(pc.less(pc.filter(arr, sel).cast(pa.int64()), pa.scalar(info.max + 10, pa.int64())))


def test_all_null_and_narrow_arrays_do_not_squeeze():
    io = O.OracleSqueezeIo()
    assert O.squeeze_int(O.OracleIntArray.from_arrow(pa.array([None] * 50, pa.int32())), io, HINT) is None  # no bit width
    assert O.squeeze_int(O.OracleIntArray.from_arrow(pa.array(list(range(100)), pa.int32())), io, HINT) is None  # width 7
    assert O.squeeze_int(O.OracleIntArray.from_arrow(pa.array(list(range(200)), pa.int32())), io, HINT) is not None  # width 8


# ---- Date32 / Timestamp columns: one date component (liquid_array/squeezed_date32_array.rs:489-747) ----
D = O.ymd_to_epoch_days


def _dates(vals):
    return pa.array(vals, pa.int32()).cast(pa.date32())


def _squeezed(field, vals):
    return O.OracleSqueezedDate32Array.from_liquid(O.OracleIntArray.from_arrow(_dates(vals)), field)


def test_date_arithmetic_matches_the_calendar():
    import datetime

    epoch = datetime.date(1970, 1, 1)
    for y, m, d in [(1970, 1, 1), (1969, 12, 31), (2000, 2, 29), (1900, 3, 1), (2024, 2, 29), (1, 1, 1), (9999, 12, 31), (1600, 2, 29)]:
        days = (datetime.date(y, m, d) - epoch).days
        assert D(y, m, d) == days
        yy, mm, dd = O.ymd_from_epoch_days(np.array([days]))
        assert (int(yy[0]), int(mm[0]), int(dd[0])) == (y, m, d)
        assert int(O.component_from_days("DayOfWeek", np.array([days]))[0]) == (datetime.date(y, m, d).weekday() + 1) % 7
    rng = np.random.default_rng(5)
    days = rng.integers(-800_000, 2_900_000, size=20_000)  # years -220 .. 9900, both sides of the civil epoch
    y, m, d = O.ymd_from_epoch_days(days)
    back = np.array([D(int(a), int(b), int(c)) for a, b, c in zip(y[:2000], m[:2000], d[:2000])])
    assert np.array_equal(back, days[:2000])
    got = pc.year(pa.array(days[(days > -719_000)].astype(np.int32), pa.int32()).cast(pa.date32()))  # arrow's calendar from year 1 on
    assert np.array_equal(np.asarray(got), y[days > -719_000])


def test_extraction_correctness():
    """:510-562"""
    assert _squeezed("Year", [-1, 0, D(1971, 7, 15), None]).to_component_date32().equals(_dates([1969, 1970, 1971, None]))
    assert _squeezed("Month", [D(1970, 1, 31), D(1970, 2, 1), D(1970, 12, 31), None]).to_component_date32().equals(_dates([1, 2, 12, None]))
    assert _squeezed("Day", [D(1970, 1, 1), D(1970, 1, 31), D(1970, 2, 1), None]).to_component_date32().equals(_dates([1, 31, 1, None]))
    assert _squeezed("DayOfWeek", [D(1970, 1, 4), D(1970, 1, 5), D(1970, 1, 10), None]).to_component_date32().equals(_dates([0, 1, 6, None]))


def test_lossy_reconstruction_mapping():
    """:565-616"""
    assert _squeezed("Year", [D(1999, 12, 31), D(2000, 6, 1), None]).to_arrow_date32_lossy().equals(_dates([D(1999, 1, 1), D(2000, 1, 1), None]))
    assert _squeezed("Month", [D(1980, 3, 14), D(1977, 12, 5), None]).to_arrow_date32_lossy().equals(_dates([D(1970, 3, 1), D(1970, 12, 1), None]))
    assert _squeezed("Day", [D(1980, 3, 14), D(1977, 12, 5), None]).to_arrow_date32_lossy().equals(_dates([D(1970, 1, 14), D(1970, 1, 5), None]))
    assert _squeezed("DayOfWeek", [D(2020, 5, 17), D(2020, 5, 18), None]).to_arrow_date32_lossy().equals(_dates([D(1970, 1, 4), D(1970, 1, 5), None]))


def test_roundtrip_idempotence_and_all_nulls():
    """:619-662"""
    vals = [D(1969, 12, 31), D(1970, 1, 1), D(1970, 1, 31), D(1970, 2, 1), D(1971, 7, 15), None]
    for field in O.DATE32_FIELDS:
        comp1 = _squeezed(field, vals).to_component_date32()
        lossy = _squeezed(field, vals).to_arrow_date32_lossy()
        comp2 = O.OracleSqueezedDate32Array.from_liquid(O.OracleIntArray.from_arrow(lossy), field).to_component_date32()
        assert comp1.equals(comp2), field
        nulls = _squeezed(field, [None, None, None])
        assert nulls.bit_width is None
        assert nulls.to_component_date32().equals(_dates([None] * 3)) and nulls.to_arrow_date32_lossy().equals(_dates([None] * 3))


def test_to_component_array_round_trips_through_extract():
    """:672-709, :712-746"""
    vals = [D(1970, 1, 1), D(1971, 7, 15), D(1999, 12, 31), D(2024, 2, 29), D(4709, 11, 24), None]
    comp = _squeezed("Year", vals).to_component_array()
    assert comp.type == pa.date32()
    want = [None if v is None else int(O.component_from_days("Year", np.array([v]))[0]) for v in vals]
    got = [None if v is None else int(O.component_from_days("Year", np.array([v]))[0]) for v in comp.cast(pa.int32()).to_pylist()]
    assert got == want == [1970, 1971, 1999, 2024, 4709, None]
    stamps = pa.array([1_609_459_200_000_000, 1_640_995_200_000_000, None], pa.int64()).cast(pa.timestamp("us"))
    sq = O.OracleSqueezedDate32Array.from_liquid(O.OracleIntArray.from_arrow(stamps), "Year")
    out = sq.to_component_array()
    assert out.type == pa.timestamp("us")
    assert out.cast(pa.int64()).to_pylist() == [D(2021, 1, 1) * 86_400_000_000, D(2022, 1, 1) * 86_400_000_000, None]
    assert sq.to_component_date32().cast(pa.int32()).to_pylist() == [2021, 2022, None]


@pytest.mark.parametrize("field", O.DATE32_FIELDS)
def test_components_agree_with_arrow_on_random_columns(field):
    rng = np.random.default_rng(len(field))
    days = rng.integers(-30_000, 60_000, size=5000).astype(np.int32)  # 1887 .. 2134
    arr = pa.array(days, pa.int32(), mask=rng.random(5000) < 0.1).cast(pa.date32())
    f = {"Year": pc.year, "Month": pc.month, "Day": pc.day, "DayOfWeek": lambda a: pc.day_of_week(a, count_from_zero=True, week_start=7)}[field]
    sq = O.OracleSqueezedDate32Array.from_liquid(O.OracleIntArray.from_arrow(arr), field)
    assert sq.to_component_date32().cast(pa.int32()).cast(pa.int64()).equals(f(arr))
    assert f(sq.to_component_array()).equals(f(arr))  # date_part over the lossy array gives the component back
    for unit in ("s", "ms", "us", "ns"):
        ts = pa.array(days.astype(np.int64) * O._TICKS_PER_DAY[unit] + rng.integers(0, O._TICKS_PER_DAY[unit], size=5000), pa.int64(),
                      mask=rng.random(5000) < 0.1).cast(pa.timestamp(unit))
        sqt = O.OracleSqueezedDate32Array.from_liquid(O.OracleIntArray.from_arrow(ts), field)
        assert sqt.to_component_date32().cast(pa.int32()).cast(pa.int64()).equals(f(ts)), unit
        assert f(sqt.to_component_array()).equals(f(ts)), unit


def test_date_columns_squeeze_only_under_a_date_field_hint():
    arr = _dates(list(range(8036, 10556)) + [None])
    liq = O.OracleIntArray.from_arrow(arr)
    io = O.OracleSqueezeIo()
    assert O.squeeze_int(liq, io, HINT) is None and O.squeeze_int(liq, io, None) is None
    sq, full = O.squeeze_int(liq, io, ("ExtractDate32", "Month"))
    io.set_bytes(full)
    assert full == O.to_bytes(liq) and sq.bit_width == 4 and sq.reference == 1 and len(sq) == len(arr)
    assert io.reads == 0 and sq.to_component_array().type == pa.date32() and io.reads == 0
    assert sq.to_arrow().equals(arr) and io.reads == 1
    sel = pa.array([i % 7 == 0 for i in range(len(arr))])
    assert sq.filter(sel).equals(pc.filter(arr, sel)) and io.reads == 2
    assert len(sq.filter(pa.array([False] * len(arr)))) == 0 and io.reads == 2
    got = sq.try_eval_predicate(">=", pa.scalar(9000, pa.int32()).cast(pa.date32()), sel)
    assert got.equals(pc.greater_equal(pc.filter(arr, sel), pa.scalar(9000, pa.int32()).cast(pa.date32()))) and io.reads == 3
