"""GPU parity of squeezed integer entries: LiquidArray::squeeze under IntegerSqueezePolicy::{Clamp, Quantize} and the
LiquidSqueezedArray calls on the result.

Reference: liquid_array/primitive_array.rs:389-499 (squeeze), liquid_array/hybrid_primitive_array.rs:72-790 (the two
arrays) and its tests :870-1291, liquid_array/mod.rs:209-263 (trait), cache/io_context.rs:144-180 (TestSqueezeIo) — all
under /root/reference/src/core/src.
Checked against the CPU restatement (oracle/liquid_oracle.py, pinned on the same reference tests in
tests/test_oracle_squeeze.py): which columns squeeze at all; the full bytes handed back; the half-width codes word for word
(FastLanes order), their bit width and the bucket width; every mask and every materialized array; and WHEN the backing
bytes are read — zero reads for the literals the reference lists as resolvable, a read for the unresolvable ones, the same
decision as the restatement everywhere else.
Columns with nulls are built with the column minimum in their null slots: the reference clamps / quantizes whatever the
Arrow buffer holds there and takes the quantize maximum over those slots too, this build over the valid rows only.
"""
import struct

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from liquid_cache_b200 import BinaryExpr, CacheExpression, Column, LiquidExpr, Literal
from liquid_cache_b200 import _native as N
from oracle import liquid_oracle as O
from tests.util import assert_arrays_equal, assert_masks_equal

pytestmark = pytest.mark.gpu

INT_HDR = struct.Struct("<I4B I I Q 6I")
HINT = CacheExpression.PredicateColumn
OPS = ["=", "!=", "<", "<=", ">", ">="]


class CountingIo:
    """TestSqueezeIo (cache/io_context.rs:144-180)"""

    def __init__(self):
        self.bytes = None
        self.reads = 0

    def set_bytes(self, b):
        self.bytes = b

    def reset_reads(self):
        self.reads = 0

    def read(self, rng):
        self.reads += 1
        return self.bytes[rng[0]:rng[1]]


def make_array(typ, n, base_min, span, null_prob, seed):
    rng = np.random.default_rng(seed)
    vals = [base_min + int(d) for d in rng.integers(0, span, size=n, endpoint=True)]
    nulls = rng.random(n) < null_prob
    nulls[0] = False
    mn = min(v for v, m in zip(vals, nulls) if not m)
    np_vals = np.array([mn if m else v for v, m in zip(vals, nulls)], dtype=typ.to_pandas_dtype())
    return pa.array(np_vals, type=typ, mask=nulls if nulls.any() else None)


def boundary_of(arr):
    mn, mx = pc.min_max(arr)["min"].as_py(), pc.min_max(arr)["max"].as_py()
    half = O.get_bit_width(mx - mn) // 2
    return mn + ((1 << half) - 1 if half else 0)


def expr_of(op, k):
    return LiquidExpr.new_unchecked(BinaryExpr(Column("col", 0), op, Literal(k)))


def squeeze_both(cache, arr, policy):
    io, oio = CountingIo(), O.OracleSqueezeIo()
    full = cache.transcode(arr)
    got = full.squeeze(io, HINT, policy)
    want = O.squeeze_int(O.OracleIntArray.from_arrow(arr), oio, "PredicateColumn", policy)
    assert (got is None) == (want is None)
    if got is None:
        return None
    (sq, image), (osq, oimage) = got, want
    assert image == oimage == full.to_bytes()
    io.set_bytes(image)
    oio.set_bytes(oimage)
    return sq, io, osq, oio, full


CASES = [(pa.int32(), 200, -1_000_000, 1 << 16, 0.2, 0x5173), (pa.uint32(), 180, 1_000_000, 1 << 16, 0.15, 0x5174),
         (pa.int64(), 8192, -(2**40), 1 << 20, 0.1, 7), (pa.uint16(), 2500, 100, 1 << 12, 0.0, 8), (pa.int8(), 300, -128, 255, 0.3, 9),
         (pa.uint64(), 1500, 2**63, 1 << 16, 0.05, 10), (pa.int16(), 5000, -20000, 1 << 15, 0.0, 11), (pa.int64(), 3000, -(2**62), 2**62, 0.1, 12)]


def test_columns_the_reference_does_not_squeeze(cache):
    io = CountingIo()
    narrow = cache.transcode(make_array(pa.int32(), 64, 10_000, 100, 0.1, 0x5171))       # clamp_unsqueezable_small_range
    assert narrow.squeeze(io, HINT, "clamp") is None and narrow.squeeze(io, HINT, "quantize") is None
    wide = cache.transcode(make_array(pa.int32(), 64, 10_000, 1 << 12, 0.1, 0x5171))
    assert wide.squeeze(io, None, "clamp") is None                                         # no hint
    assert wide.squeeze(io, HINT, "clamp") is not None
    assert cache.transcode(pa.array([None] * 50, pa.int32())).squeeze(io, HINT) is None    # no bit width
    assert cache.transcode(pa.array(list(range(100)), pa.int32())).squeeze(io, HINT) is None   # width 7
    assert cache.transcode(pa.array(list(range(200)), pa.int32())).squeeze(io, HINT) is not None   # width 8
    dates = pa.array(list(range(8036, 10556)), pa.int32()).cast(pa.date32())
    assert cache.transcode(dates).squeeze(io, HINT) is None                                # wants a date-field hint
    stamps = pa.array([i * 1_000_000 for i in range(3000)], pa.int64()).cast(pa.timestamp("us"))
    assert cache.transcode(stamps).squeeze(io, HINT) is None
    assert cache.transcode(pa.array([0.5 * i for i in range(3000)])).squeeze(io, HINT) is None      # floats: not built
    assert cache.transcode(pa.array([f"s{i}" for i in range(300)])).squeeze(io, HINT) is None      # byte views: not built


@pytest.mark.parametrize("policy", ["clamp", "quantize"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[1]}")
def test_squeezed_codes_match_the_restatement(cache, policy, case):
    typ, n, base, span, null_p, seed = case
    arr = make_array(typ, n, base, span, null_p, seed)
    sq, io, osq, oio, full = squeeze_both(cache, arr, policy)
    assert sq.policy() == policy and sq.len() == len(arr) and sq.original_arrow_data_type() == arr.type
    assert sq.bit_width() == osq.bit_width == O.OracleIntArray.from_arrow(arr).bit_width // 2
    assert sq.disk_backing() == len(io.bytes)
    if policy == "quantize":
        assert sq.bucket_width() == osq.bucket_width
    assert sq.get_array_memory_size() < full.get_array_memory_size()
    with pytest.raises(N.NativeError):
        sq.to_bytes()
    with pytest.raises(N.NativeError):  # batched reads and the scan calls take full entries only
        cache.to_arrow_many(np.array([sq.handle], dtype=np.uint64), None)
    img = sq.entry_image()  # last: a dry run without a device stops here
    magic, phys, tbits, bit_width, has_nulls, nn, n_chunks, reference, validity_off, packed_off, blob_bytes, null_count, \
        is_signed, _ = INT_HDR.unpack_from(img, 0)
    assert (nn, bit_width, null_count) == (len(arr), osq.bit_width, arr.null_count)
    assert reference == osq.reference & ((1 << tbits) - 1)
    words = np.frombuffer(img, dtype=osq.packed.dtype, count=len(osq.packed), offset=packed_off)
    assert np.array_equal(words, osq.packed), "half-width codes differ from the restatement's"


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[1]}")
def test_clamp_predicates_resolvable_and_unresolvable(cache, case):
    """clamp_predicate_eval_{i32,u32}_resolvable_and_unresolvable (:935-1110)"""
    typ, n, base, span, null_p, seed = case
    arr = make_array(typ, n, base, span, null_p, seed)
    sq, io, osq, oio, _full = squeeze_both(cache, arr, "clamp")
    boundary = boundary_of(arr)
    sel = pa.array(np.random.default_rng(seed + 1).random(n) < 0.5)
    resolvable = [("=", boundary - 1), ("!=", boundary - 1), ("<", boundary), ("<=", boundary - 1), (">", boundary - 1), (">=", boundary)]
    unresolvable = [("=", boundary), ("!=", boundary), ("<", boundary + 1), ("<=", boundary), (">", boundary + 1), (">=", boundary + 1)]
    for cases, reads in ((resolvable, False), (unresolvable, True)):
        for op, k in cases:
            io.reset_reads()
            oio.reset_reads()
            got = sq.try_eval_predicate(expr_of(op, k), sel)
            want = osq.try_eval_predicate(op, k, sel)
            assert_masks_equal(got, want, f"clamp {typ} {op} {k}")
            assert_masks_equal(got, O._PC_CMP[op](pc.filter(arr, sel), pa.scalar(k, arr.type)), f"clamp {typ} {op} {k} vs arrow")
            assert (io.reads > 0) == reads == (oio.reads > 0), (op, k, io.reads, oio.reads)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[1]}")
def test_quantize_predicates_resolvable_and_unresolvable(cache, case):
    """quantize_predicate_eval_{u32,i32}_resolvable_and_unresolvable (:1112-1276)"""
    typ, n, base, span, null_p, seed = case
    arr = make_array(typ, n, base, span, null_p, seed)
    sq, io, osq, oio, _full = squeeze_both(cache, arr, "quantize")
    mn = pc.min_max(arr)["min"].as_py()
    info = np.iinfo(typ.to_pandas_dtype())
    lo = max(mn - 1, info.min)  # min.saturating_sub(1)
    sel = pa.array([True] * n)
    consts = [("=", lo, False), ("!=", lo, True), ("<", mn, False), ("<=", lo, False), (">", lo, True), (">=", mn, True)]
    for op, k, const in consts:
        if k == mn and op in ("=", "!=", "<=", ">"):
            continue  # the column starts at the type's minimum: min - 1 saturates onto a present value
        io.reset_reads()
        got = sq.try_eval_predicate(expr_of(op, k), sel)
        want = pa.array([None if v is None else const for v in arr.to_pylist()], pa.bool_())
        assert_masks_equal(got, want, f"quantize {typ} {op} {k}")
        assert io.reads == 0, (op, k)
    k_present = next(v for v in arr.to_pylist() if v is not None)
    io.reset_reads()
    got = sq.try_eval_predicate(expr_of("=", k_present), sel)
    assert_masks_equal(got, pc.equal(arr, pa.scalar(k_present, arr.type)), "quantize = present value")
    assert io.reads > 0


@pytest.mark.parametrize("policy", ["clamp", "quantize"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[1]}")
def test_every_answer_and_every_read_decision_matches(cache, policy, case):
    typ, n, base, span, null_p, seed = case
    arr = make_array(typ, n, base, span, null_p, seed)
    sq, io, osq, oio, _full = squeeze_both(cache, arr, policy)
    rng = np.random.default_rng(seed + 2)
    info = np.iinfo(typ.to_pandas_dtype())
    mn, mx = pc.min_max(arr)["min"].as_py(), pc.min_max(arr)["max"].as_py()
    bw, b = getattr(osq, "bucket_width", 1), boundary_of(arr)
    lits = {mn - 1, mn, mn + 1, mx - 1, mx, mx + 1, b - 1, b, b + 1, mn + bw - 1, mn + bw, mn + 5 * bw, mn + 5 * bw + 1, mn + 6 * bw - 1,
            info.min, info.max}
    lits |= {int(v) for v in rng.choice([v for v in arr.to_pylist() if v is not None], 5)}
    for sel in (pa.array(rng.random(n) < 0.6), pa.array([True] * n), pa.array(rng.random(n) < 0.01)):
        for k in sorted(x for x in lits if info.min <= x <= info.max):
            for op in OPS:
                io.reset_reads()
                oio.reset_reads()
                got = sq.try_eval_predicate(expr_of(op, k), sel)
                want = osq.try_eval_predicate(op, k, sel)
                assert_masks_equal(got, want, f"{policy} {typ} {op} {k}")
                assert io.reads == oio.reads, (policy, op, k, io.reads, oio.reads)


@pytest.mark.parametrize("case", CASES[:4], ids=lambda c: f"{c[0]}-{c[1]}")
def test_materializing_a_squeezed_entry(cache, case):
    """clamp_squeeze_full_read_roundtrip_i32 (:886-932), quantize_to_arrow_is_err (:1278-1291)"""
    typ, n, base, span, null_p, seed = case
    arr = make_array(typ, n, base, span, null_p, seed)
    sq, io, osq, oio, _full = squeeze_both(cache, arr, "clamp")
    boundary = boundary_of(arr)
    known = pa.array([True if v is None else v < boundary for v in arr.to_pylist()])
    io.reset_reads()
    assert_arrays_equal(sq.filter(known), pc.filter(arr, known), "rows below the boundary")
    assert io.reads == 0
    io.reset_reads()
    assert_arrays_equal(sq.to_arrow_array(), arr, "everything")
    assert io.reads > 0
    io.reset_reads()
    assert len(sq.filter(pa.array([False] * n))) == 0 and io.reads == 0
    sel = pa.array(np.random.default_rng(seed).random(n) < 0.3)
    oio.reset_reads()
    io.reset_reads()
    assert_arrays_equal(sq.filter(sel), osq.filter(sel), "random selection")
    assert io.reads == oio.reads
    q, qio, oq, oqio, _ = squeeze_both(cache, arr, "quantize")
    qio.reset_reads()
    assert_arrays_equal(q.to_arrow_array(), arr, "quantize: to_arrow")
    assert qio.reads > 0
    qio.reset_reads()
    assert_arrays_equal(q.filter(sel), pc.filter(arr, sel), "quantize: filter")
    assert qio.reads > 0


def test_a_failing_or_wrong_backing_is_reported(cache):
    arr = make_array(pa.int32(), 4000, -5000, 1 << 16, 0.1, 3)
    full = cache.transcode(arr)

    class Broken(CountingIo):
        def read(self, rng):
            self.reads += 1
            raise OSError("gone")

    sq, _bytes = full.squeeze(Broken(), HINT, "quantize")
    with pytest.raises(N.NativeError):
        sq.to_arrow_array()
    other = cache.transcode(make_array(pa.int32(), 3000, 9, 1 << 16, 0.0, 4)).to_bytes()
    io = CountingIo()
    sq2, image = full.squeeze(io, HINT, "quantize")
    io.set_bytes(other.ljust(len(image), b"\0"))  # a well-formed image of another column (3000 rows, not 4000)
    with pytest.raises(N.NativeError):
        sq2.to_arrow_array()
    io.set_bytes(image)
    assert_arrays_equal(sq2.to_arrow_array(), arr, "with the right bytes again")


# ---- Date32 / Timestamp columns: one date component (liquid_array/squeezed_date32_array.rs) ----
FIELDS = ["Year", "Month", "Day", "DayOfWeek"]
D = O.ymd_to_epoch_days


def _date_cases():
    rng = np.random.default_rng(404)
    days = rng.integers(-30_000, 60_000, size=8192).astype(np.int32)
    out = [pa.array(days, pa.int32(), mask=rng.random(8192) < 0.1).cast(pa.date32()),
           pa.array([D(1969, 12, 31), D(1970, 1, 1), D(1970, 1, 31), D(1970, 2, 1), D(1971, 7, 15), None], pa.int32()).cast(pa.date32()),
           pa.array([D(1970, 1, 1), D(1971, 7, 15), D(1999, 12, 31), D(2024, 2, 29), D(4709, 11, 24), None], pa.int32()).cast(pa.date32()),
           pa.array([-1, 0, D(1971, 7, 15), None, -719_468, -800_000], pa.int32()).cast(pa.date32()),   # both sides of the civil epoch
           pa.array([None, None, None], pa.int32()).cast(pa.date32()), pa.array([D(2000, 2, 29)] * 1500, pa.int32()).cast(pa.date32()),
           pa.array([1_609_459_200_000_000, 1_640_995_200_000_000, None, -1, 0], pa.int64()).cast(pa.timestamp("us"))]
    for unit in ("s", "ms", "us", "ns"):
        t = O._TICKS_PER_DAY[unit]
        ticks = days[:5000].astype(np.int64) * t + rng.integers(0, t, size=5000)
        out.append(pa.array(ticks, pa.int64(), mask=rng.random(5000) < 0.05).cast(pa.timestamp(unit)))
    return out


@pytest.mark.parametrize("field", FIELDS)
def test_date_component_squeeze_matches_the_restatement(cache, field):
    images = []
    for ci, arr in enumerate(_date_cases()):
        io, oio = CountingIo(), O.OracleSqueezeIo()
        full = cache.transcode(arr)
        assert full.squeeze(io, HINT) is None and full.squeeze(io, None) is None  # only a date-field hint squeezes these
        hint = CacheExpression.extract_date32(field)
        sq, image = full.squeeze(io, hint)
        osq, oimage = O.squeeze_int(O.OracleIntArray.from_arrow(arr), oio, hint)
        assert len(image) == len(oimage)
        io.set_bytes(image)
        what = f"{arr.type} case {ci} {field}"
        assert sq.policy() == "date32" and sq.field() == field and sq.len() == len(arr) and sq.original_arrow_data_type() == arr.type
        assert sq.bit_width() == (osq.bit_width or 0) and sq.disk_backing() == len(image)
        images.append((sq, osq, what))
        io.reset_reads()
        assert_arrays_equal(sq.to_component_date32(), osq.to_component_date32(), what + ": to_component_date32")
        comp = sq.to_component_array()
        assert comp.type == arr.type
        assert_arrays_equal(comp, osq.to_component_array(), what + ": to_component_array")
        assert io.reads == 0
        raw = arr.cast(pa.int32() if pa.types.is_date32(arr.type) else pa.int64())  # raw day / tick counts
        per_day = O._TICKS_PER_DAY[arr.type.unit] if pa.types.is_timestamp(arr.type) else 1
        if arr.null_count < len(arr) and pc.min(raw).as_py() > -700_000 * per_day:  # arrow's calendar starts at year 1
            f = {"Year": pc.year, "Month": pc.month, "Day": pc.day,
                 "DayOfWeek": lambda a: pc.day_of_week(a, count_from_zero=True, week_start=7)}[field]
            assert f(comp).equals(f(arr)), what + ": date_part over the component array"
        # everything else reads the backing bytes
        assert_arrays_equal(sq.to_arrow_array(), arr, what + ": to_arrow")
        assert io.reads == 1
        sel = pa.array(np.random.default_rng(ci).random(len(arr)) < 0.4)
        assert_arrays_equal(sq.filter(sel), pc.filter(arr, sel), what + ": filter")
        before = io.reads
        empty = sq.filter(pa.array([False] * len(arr)))
        assert len(empty) == 0 and empty.type == arr.type and io.reads == before
        lit = next((v for v in raw.to_pylist() if v is not None), None)
        if lit is not None and pc.any(sel).as_py():
            got = sq.try_eval_predicate(expr_of(">=", lit), sel)
            assert_masks_equal(got, pc.greater_equal(pc.filter(raw, sel), pa.scalar(lit, raw.type)), what + ": >=")
            assert io.reads == before + 1
    for sq, osq, what in images:  # last: a dry run without a device stops here
        img = sq.entry_image()
        magic, phys, tbits, bit_width, has_nulls, nn, n_chunks, reference, *_ = INT_HDR.unpack_from(img, 0)
        assert tbits == 32 and reference == osq.reference & 0xFFFFFFFF, what
        if osq.bit_width is not None:
            packed_off = INT_HDR.unpack_from(img, 0)[9]
            words = np.frombuffer(img, dtype=np.uint32, count=len(osq.packed), offset=packed_off)
            assert np.array_equal(words, osq.packed), what + ": packed component offsets"


# ---- one call over a list of entries: full, clamped and quantized batches of one column ----
def _mask_of(vals, valid, off, length, nulls):
    bits = np.unpackbits(vals[off:off + (length + 7) // 8], bitorder="little")[:length].astype(bool)
    if nulls == 0:
        return pa.array(bits, pa.bool_())
    ok = np.unpackbits(valid[off:off + (length + 7) // 8], bitorder="little")[:length].astype(bool)
    return pa.array(bits & ok, pa.bool_(), mask=~ok)


@pytest.mark.parametrize("typ,base,span", [(pa.int64(), -(2**40), 1 << 20), (pa.uint32(), 1_000_000, 1 << 16), (pa.int16(), -20000, 1 << 15)], ids=str)
def test_batched_predicates_over_full_and_squeezed_entries(cache, typ, base, span):
    rng = np.random.default_rng(span)
    arrays, handles, ios, oracles, keep = [], [], [], [], []
    for b in range(12):
        arr = make_array(typ, int(rng.integers(900, 8193)), base + int(rng.integers(0, span // 4)), span, 0.1 if b % 3 else 0.0, 100 + b)
        full = cache.transcode(arr)
        form = ("full", "clamp", "quantize")[b % 3]
        if form == "full":
            entry, io, osq = full, None, None
        else:
            io, oio = CountingIo(), O.OracleSqueezeIo()
            entry, image = full.squeeze(io, HINT, form)
            osq, oimage = O.squeeze_int(O.OracleIntArray.from_arrow(arr), oio, "PredicateColumn", form)
            io.set_bytes(image)
            oio.set_bytes(oimage)
            osq._io = oio
        arrays.append(arr)
        handles.append(entry.handle)
        ios.append(io)
        oracles.append(osq)
        keep.append((full, entry))
    rows = np.array([len(a) for a in arrays], dtype=np.uint64)
    hs = np.array(handles, dtype=np.uint64)
    all_vals = np.concatenate([np.asarray(a.drop_null().cast(pa.int64() if typ != pa.uint64() else pa.uint64())) for a in arrays])
    lits = sorted({int(all_vals.min()) - 1, int(all_vals.min()), int(np.median(all_vals)), int(all_vals.max()), int(all_vals.max()) + 1,
                   boundary_of(arrays[1]), boundary_of(arrays[1]) - 1, int(all_vals[7]), int(all_vals[-3])})
    info = np.iinfo(typ.to_pandas_dtype())
    for sel_p in (None, 0.5, 0.01):
        sels = None if sel_p is None else [np.packbits(rng.random(len(a)) < sel_p, bitorder="little") for a in arrays]
        for k in (x for x in lits if info.min <= x <= info.max):
            for op in OPS:
                for io in ios:
                    if io is not None:
                        io.reset_reads()
                vals, valid, offs, out_len, out_nulls, out_true = cache.eval_predicate_many(hs, rows, expr_of(op, k), typ, sels)
                for i, arr in enumerate(arrays):
                    sel = pa.array([True] * len(arr)) if sels is None else pa.array(np.unpackbits(sels[i], bitorder="little")[:len(arr)].astype(bool))
                    want = O._PC_CMP[op](pc.filter(arr, sel), pa.scalar(k, typ))
                    got = _mask_of(vals, valid, int(offs[i]), int(out_len[i]), int(out_nulls[i]))
                    assert_masks_equal(got, want, f"{typ} batch {i} {op} {k} sel={sel_p}")
                    assert int(out_true[i]) == pc.sum(pc.fill_null(want, False)).as_py() or (int(out_true[i]) == 0 and not pc.any(pc.fill_null(want, False)).as_py())
                    if oracles[i] is not None:  # the backing is read exactly when the restatement reads it
                        oracles[i]._io.reset_reads()
                        oracles[i].try_eval_predicate(op, k, sel)
                        assert ios[i].reads == oracles[i]._io.reads, (i, op, k, sel_p, ios[i].reads, oracles[i]._io.reads)


# ---- the device-resident scan pipeline over squeezed entries (lc_scan_filter) ----
@pytest.mark.parametrize("typ,base,span", [(pa.int64(), -(2**40), 1 << 20), (pa.uint32(), 1_000_000, 1 << 16)], ids=str)
def test_scan_filter_over_full_and_squeezed_entries(cache, typ, base, span):
    rng = np.random.default_rng(span + 1)
    arrays, mixed, fulls, ios, keep = [], [], [], [], []
    for b in range(9):
        arr = make_array(typ, int(rng.integers(900, 8193)), base + int(rng.integers(0, span // 4)), span, 0.1 if b % 3 else 0.0, 300 + b)
        full = cache.transcode(arr)
        form = ("full", "clamp", "quantize")[b % 3]
        io = None
        entry = full
        if form != "full":
            io = CountingIo()
            entry, image = full.squeeze(io, HINT, form)
            io.set_bytes(image)
        arrays.append(arr)
        fulls.append(full.handle)
        mixed.append(entry.handle)
        ios.append(io)
        keep.append((full, entry))
    rows = [len(a) for a in arrays]
    h_full, h_mixed = np.array(fulls, dtype=np.uint64), np.array(mixed, dtype=np.uint64)
    all_vals = np.concatenate([np.asarray(a.drop_null().cast(pa.int64())) for a in arrays])
    lo, hi, present = int(np.quantile(all_vals, 0.3)), int(np.quantile(all_vals, 0.8)), int(all_vals[11])
    conjunct_sets = [[(">=", lo), ("<", hi)], [("=", present)], [("!=", present), ("<=", hi), (">", lo)], [("<", int(all_vals.min()))],
                     [(">=", boundary_of(arrays[1])), ("<", boundary_of(arrays[1]) + 3)]]
    for seeded in (False, True):
        for conjuncts in conjunct_sets:
            with cache.scan(rows) as want_scan, cache.scan(rows) as got_scan:
                if seeded:  # a selection seeded from the host, then refined on the device
                    for b, a in enumerate(arrays):
                        sel = pa.array(np.random.default_rng(b).random(len(a)) < 0.5)
                        want_scan.set_selection(b, sel)
                        got_scan.set_selection(b, sel)
                for op, k in conjuncts:
                    want_scan.filter(h_full, expr_of(op, k), typ)
                    got_scan.filter(h_mixed, expr_of(op, k), typ)
                    wc, wt = want_scan.counts()
                    gc, gt = got_scan.counts()
                    assert gt == wt and np.array_equal(gc, wc), (seeded, conjuncts, op, k)
                for b in range(len(arrays)):
                    assert_masks_equal(got_scan.selection(b), want_scan.selection(b), f"batch {b} after {conjuncts} seeded={seeded}")
    assert any(io is not None and io.reads for io in ios)  # some conjunct did go back to the backing bytes
