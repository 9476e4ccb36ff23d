"""GPU parity of LiquidFixedLenByteArray: Decimal128 / Decimal256 arrays with a value outside u64.

Reference: cache/transcode.rs:118-153 (dispatch: fits_u64 -> LiquidDecimalArray, else this type under the column chunk's
FSST compressor), liquid_array/fix_len_byte_array.rs:26-420 (u16 dictionary over the 16 / 32-byte values, FSST, keyed or
full decompression) and its tests :452-598 (round trips and filters over generated decimals); the type has no predicate
of its own (LiquidArray default, liquid_array/mod.rs:116-130).
Checked: the dispatch; bit-exact round trips and filters against the input / the oracle for both widths, with nulls,
negatives and many duplicates; the dictionary the device built (keys in first-occurrence order, values decompressing to
the oracle's, stored in order-preserving byte form); the six comparisons against Arrow's decimal kernels — the reference
evaluates them on the decoded rows, here they run on the dictionary; batched reads and predicates over several entries; what
is declined (serialization, > 65536 distinct values).
"""
import decimal

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from liquid_cache_b200 import BinaryExpr, Column, EntryID, LiquidExpr, Literal
from liquid_cache_b200 import _native as N
from oracle import liquid_oracle as O
from tests.test_gpu_insert_layout import fsst_decompress, parse_str_image
from tests.test_oracle_fixed_len import gen_decimals
from tests.util import assert_arrays_equal, assert_masks_equal

pytestmark = pytest.mark.gpu

TYPES = [pa.decimal128(38, 6), pa.decimal128(20, 0), pa.decimal256(60, 10), pa.decimal256(76, 0)]
OPS = ["=", "!=", "<", "<=", ">", ">="]


def ordered(le: bytes) -> bytes:
    """the order-preserving form the entry stores: big-endian, sign bit flipped (csrc/k_bits.cu k_fixed_to_ordered)"""
    be = le[::-1]
    return bytes([be[0] ^ 0x80]) + be[1:]


def expr_of(op, value):
    return LiquidExpr.new_unchecked(BinaryExpr(Column("c", 0), op, Literal(value)))


@pytest.mark.parametrize("typ", TYPES, ids=str)
@pytest.mark.parametrize("n,n_distinct,null_p", [(8192, 1500, 0.1), (3000, 3000, 0.0), (5, 3, 0.4), (8192, 7, 0.02)])
def test_round_trip_filter_and_dictionary(cache, typ, n, n_distinct, null_p):
    arr = gen_decimals(typ, n, n_distinct, null_p, n + typ.precision)
    if O.OracleDecimalArray.fits_u64(arr):
        pytest.skip("this draw fits u64")
    scope = (77 << 32) | (typ.precision << 16)
    liquid = cache.transcode(arr, compressor_scope=scope)
    oracle = O.transcode(arr)
    assert isinstance(oracle, O.OracleFixedLenByteArray)
    assert liquid.data_type() == 3 and liquid.len() == len(arr) and liquid.original_arrow_data_type() == typ
    assert_arrays_equal(liquid.to_arrow_array(), arr, f"{typ} round trip")
    rng = np.random.default_rng(n)
    for p in (0.5, 0.01, 1.0):
        sel = pa.array(rng.random(len(arr)) < p)
        assert_arrays_equal(liquid.filter(sel), oracle.filter(sel), f"{typ} filter p={p}")
    assert len(liquid.filter(pa.array([False] * len(arr)))) == 0
    # a second batch of the chunk reuses the symbol table and stays readable
    more = gen_decimals(typ, 1000, 200, 0.05, 5 + typ.precision)
    assert_arrays_equal(cache.transcode(more, compressor_scope=scope).to_arrow_array(), more, f"{typ} second batch")
    # (last: a dry run without a device stops here) the dictionary: first-occurrence keys, every value 16 / 32 bytes and equal to the oracle's after decompression
    h = parse_str_image(liquid.entry_image())
    want_keys = np.array([0 if k is None else k for k in oracle.keys], dtype=np.uint16)
    valid = np.array([k is not None for k in oracle.keys], dtype=bool)
    assert h["n_unique"] == len(oracle.uniques) and np.array_equal(h["valid"], valid)
    assert np.array_equal(h["keys"][valid], want_keys[valid])
    offs = [(h["slope"] * i + h["intercept"] + int(h["resid"][i])) & 0xFFFFFFFF for i in range(h["n_unique"] + 1)]
    table = liquid.fsst_table()
    for u in range(0, h["n_unique"], max(1, h["n_unique"] // 64)):
        assert fsst_decompress(table, h["comp"][offs[u]:offs[u + 1]]) == ordered(oracle.uniques[u]), f"dictionary value {u}"


@pytest.mark.parametrize("typ", TYPES, ids=str)
def test_comparisons_match_arrow(cache, typ):
    arr = gen_decimals(typ, 6000, 400, 0.1, 17 + typ.precision)
    liquid = cache.transcode(arr)
    assert liquid.data_type() == 3
    rng = np.random.default_rng(typ.precision)
    present = [v for v in arr.to_pylist() if v is not None]
    with decimal.localcontext() as cx:
        cx.prec = 100
        unit = decimal.Decimal(1).scaleb(-typ.scale)
        lits = [present[0], present[1] + unit, present[2] - unit, min(present), max(present), min(present) - unit, max(present) + unit,
                decimal.Decimal(0), -unit, unit, decimal.Decimal(2**64).scaleb(-typ.scale), decimal.Decimal(-(2**100)).scaleb(-typ.scale)]
    lits = [v for v in lits if abs(int(v.scaleb(typ.scale))) < 10 ** typ.precision]  # representable in the column's type
    for sel in (pa.array([True] * len(arr)), pa.array(rng.random(len(arr)) < 0.3)):
        for lit in lits:
            for op in OPS:
                got = liquid.try_eval_predicate(expr_of(op, lit), sel)
                want = O._PC_CMP[op](pc.filter(arr, sel), pa.scalar(lit, typ))
                assert_masks_equal(got, want, f"{typ} {op} {lit}")


def test_dispatch_follows_fits_u64(cache):
    fits = pa.array([decimal.Decimal("12.50"), None, decimal.Decimal("0.01")], pa.decimal128(15, 2))
    assert cache.transcode(fits).data_type() == 6                      # LiquidDecimalArray
    neg = pa.array([decimal.Decimal("-0.01"), decimal.Decimal("1.00"), None], pa.decimal128(15, 2))
    liquid = cache.transcode(neg)
    assert liquid.data_type() == 3                                     # one negative value: LiquidFixedLenByteArray
    assert_arrays_equal(liquid.to_arrow_array(), neg, "negative decimal")
    big = pa.array([decimal.Decimal(2**64), decimal.Decimal(1)], pa.decimal256(40, 0))
    assert cache.transcode(big).data_type() == 3
    assert_arrays_equal(cache.transcode(big).to_arrow_array(), big, "2^64")
    sliced = gen_decimals(pa.decimal128(38, 6), 4000, 900, 0.2, 3).slice(1000, 2500)  # offset into values and validity
    assert_arrays_equal(cache.transcode(sliced).to_arrow_array(), sliced, "sliced input")


def test_cache_level_calls_and_batched_reads(cache):
    typ = pa.decimal128(38, 6)
    arrays = [gen_decimals(typ, n, 300, 0.1, 40 + i) for i, n in enumerate((8192, 100, 8192, 1))]
    ids = [EntryID((91 << 48) | (2 << 32) | (4 << 16) | i) for i in range(len(arrays))]
    for eid, arr in zip(ids, arrays):
        cache.insert(eid, arr).run()
    cache.insert_many([EntryID(int(i) + 100) for i in ids], arrays)   # the list form takes them batch by batch
    for eid, arr in zip(ids, arrays):
        assert_arrays_equal(cache.get(eid).read(), arr, "get")
        assert_arrays_equal(cache.get(EntryID(int(eid) + 100)).read(), arr, "get after insert_many")
    handles = cache.handles(ids)
    sels = [np.packbits(np.random.default_rng(i).random(len(a)) < 0.4, bitorder="little") for i, a in enumerate(arrays)]
    got = cache.to_arrow_many(handles, sels)
    want = pa.concat_arrays([pc.filter(a, pa.array(np.unpackbits(s, bitorder="little")[:len(a)].astype(bool))) for a, s in zip(arrays, sels)])
    assert_arrays_equal(got, want, "to_arrow_many over four fixed-length entries")
    lit = next(v for v in arrays[0].to_pylist() if v is not None)
    for eid, arr in zip(ids, arrays):
        sel = pa.array(np.random.default_rng(int(eid) & 0xFF).random(len(arr)) < 0.5)
        got = cache.eval_predicate(eid, expr_of("<", lit)).with_selection(sel).read()
        assert_masks_equal(got, pc.less(pc.filter(arr, sel), pa.scalar(lit, typ)), "eval_predicate < through the cache")


def test_what_is_declined(cache):
    typ = pa.decimal128(38, 6)
    arr = gen_decimals(typ, 2000, 300, 0.1, 9)
    liquid = cache.transcode(arr)
    with pytest.raises(N.UnsupportedType):
        liquid.to_bytes()
    with decimal.localcontext() as cx:
        cx.prec = 60
        many = pa.array([decimal.Decimal(-i - 1) for i in range(70_000)], pa.decimal128(38, 0))
    with pytest.raises(N.UnsupportedType):   # the reference's u16 dictionary builder overflows here
        cache.transcode(many)


def test_literals_wider_than_128_bits(cache):
    """A Decimal256 literal beyond i128 crosses the ABI as its 32 little-endian bytes: on a LiquidFixedLenByteArray entry it
    is compared like any other, on a u64-shaped LiquidDecimalArray entry it folds to a constant side."""
    typ = pa.decimal256(60, 0)
    with decimal.localcontext() as cx:
        cx.prec = 100
        small = pa.array([decimal.Decimal(5), None, decimal.Decimal(2**63), decimal.Decimal(0)], typ)         # fits u64
        wide = pa.array([decimal.Decimal(10**50), decimal.Decimal(-(10**45)), None, decimal.Decimal(7)], typ)  # does not
        lits = [decimal.Decimal(10**50), decimal.Decimal(10**50 + 1), decimal.Decimal(-(10**45)), decimal.Decimal(-(10**55)), decimal.Decimal(2**127)]
    for arr, kind in ((small, 6), (wide, 3)):
        liquid = cache.transcode(arr)
        assert liquid.data_type() == kind
        sel = pa.array([True] * len(arr))
        for lit in lits:
            for op in OPS:
                got = liquid.try_eval_predicate(expr_of(op, lit), sel)
                assert_masks_equal(got, O._PC_CMP[op](arr, pa.scalar(lit, typ)), f"kind {kind}: {op} {lit}")
