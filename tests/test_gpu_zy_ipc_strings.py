"""GPU parity of the byte-view half of LQDA: LiquidByteViewArray::to_bytes / from_bytes and the symbol table that travels
beside it.

Reference: liquid_array/byte_view_array/serialization.rs:87-325 (sections: FSST values, bit-packed u16 keys, CompactOffsets,
prefix keys, shared prefix, fingerprints), liquid_array/raw/fsst_buffer.rs:854-932 (save/load_symbol_table),
liquid_array/ipc.rs:238-283 (the compressor comes from LiquidIPCContext, not from the image).
Checked: an image the device writes parses in the CPU restatement to the same dictionary, keys, prefix keys, offsets and
fingerprints the restatement builds from the Arrow input (the FSST table differs, so the compressed bytes are compared
after decompression); an image the restatement wrote — with ITS symbol table loaded into a fresh scope — becomes an entry
that reads back, filters and evaluates predicates like the original and serializes to the same bytes; to_bytes -> from_bytes
is the identity; damaged images and images without a symbol table are refused.
"""
import numpy as np
import pyarrow as pa
import pytest

from liquid_cache_b200 import BinaryExpr, CacheExpression, Column, LikeExpr, LiquidExpr, Literal
from liquid_cache_b200 import _native as N
from oracle import liquid_oracle as O
from tests.util import assert_arrays_equal, assert_masks_equal

pytestmark = pytest.mark.gpu

TYPES = [pa.string(), pa.binary(), pa.string_view(), pa.binary_view(), pa.dictionary(pa.uint16(), pa.string()),
         pa.dictionary(pa.uint16(), pa.binary())]


def _build(vals, typ):
    if pa.types.is_dictionary(typ):
        text = pa.types.is_string(typ.value_type)
        plain = pa.array([None if v is None else (v if text else v.encode()) for v in vals], typ.value_type)
        return plain.dictionary_encode().cast(typ)
    text = pa.types.is_string(typ) or pa.types.is_string_view(typ)
    return pa.array([None if v is None else (v if text else v.encode()) for v in vals], typ)


def _cases():
    rng = np.random.default_rng(77)
    urls = [f"http://host{int(i)}.example.com/{'google' if i % 13 == 0 else 'page'}/{int(i) * 7919 % 1000}" for i in rng.integers(0, 900, 6000)]
    with_nulls = [None if rng.random() < 0.1 else u for u in urls[:3000]]
    return [["hello_world", "hello_rust", None, "hello_test", "hello_world"],  # byte_view_array/tests.rs
            [], [None, None, None], ["", "", None, ""], ["only"], urls, with_nulls,
            [f"{i:05d}" for i in range(2500)],                      # every row its own value, more than one key chunk
            ["x" * 300 + str(i % 40) for i in range(1500)]]         # long shared prefix, residual width > 1


def _oracle_parts(o):
    return dict(keys=list(o.keys), pkeys=o.prefix_keys, sp=o.shared_prefix, fps=(o.fingerprints or None),
                resid=o.offsets.residuals, slope=o.offsets.slope, intercept=o.offsets.intercept, uniques=list(o.uniques))


@pytest.mark.parametrize("typ", TYPES, ids=str)
@pytest.mark.parametrize("fp", [False, True], ids=["plain", "fingerprints"])
def test_device_image_parses_like_the_restatement(cache, typ, fp):
    for ci, vals in enumerate(_cases()):
        arr = _build(vals, typ)
        scope = 1000 + 100 * TYPES.index(typ) + 20 * int(fp) + ci
        liquid = cache.transcode(arr, hint=CacheExpression.SubstringSearch if fp else None, compressor_scope=scope)
        image = liquid.to_bytes()
        assert image[0:4] == O.LQDA_MAGIC and int.from_bytes(image[4:6], "little") == 1
        assert int.from_bytes(image[6:8], "little") == 4 and int.from_bytes(image[8:10], "little") == O._arrow_byte_type_id(arr.type)
        table = O.load_symbol_table(cache.save_symbol_table(scope))
        got = O.byte_view_from_bytes(image, table, arr.type)
        want = O.OracleByteViewArray.from_arrow(arr, build_fingerprints=fp)
        g, w = _oracle_parts(got), _oracle_parts(want)
        for k in ("keys", "pkeys", "sp", "fps", "uniques"):
            assert g[k] == w[k], f"{typ} case {ci}: section {k} differs"
        # CompactOffsets are a fit over the COMPRESSED offsets, which depend on the table: they must reproduce them
        if want.uniques:
            ends = [0] + [int(x) for x in np.cumsum([len(table.compress(u)) for u in want.uniques])]
            assert [got.offsets.get_offset(i) for i in range(len(ends))] == ends, f"{typ} case {ci}: offsets"
        assert_arrays_equal(got.to_arrow(), arr, f"{typ} case {ci}: parsed image")
        # and the section sizes in the header add up to the image
        sizes = [int.from_bytes(image[16 + 4 * i:20 + 4 * i], "little") for i in range(5)]
        assert sizes[4] == (4 * len(want.uniques) if fp else 0) and sizes[2] == len(want.shared_prefix)


@pytest.mark.parametrize("typ", TYPES, ids=str)
@pytest.mark.parametrize("fp", [False, True], ids=["plain", "fingerprints"])
def test_restatement_image_becomes_an_entry(cache, typ, fp):
    for ci, vals in enumerate(_cases()):
        arr = _build(vals, typ)
        o = O.OracleByteViewArray.from_arrow(arr, build_fingerprints=fp)
        image = O.byte_view_to_bytes(o)
        scope = 2000 + 100 * TYPES.index(typ) + 20 * int(fp) + ci  # one scope per image: a scope takes one table, once
        cache.load_symbol_table(scope, O.save_symbol_table(o.fsst))
        assert cache.save_symbol_table(scope) == O.save_symbol_table(o.fsst)
        entry = cache.read_from_bytes(image, compressor_scope=scope)
        assert entry.len() == len(arr)
        assert len(arr) == 0 or entry.original_arrow_data_type() == arr.type
        assert_arrays_equal(entry.to_arrow_array(), o.to_arrow(), f"{typ} case {ci}: read_from_bytes")
        assert entry.to_bytes() == image, f"{typ} case {ci}: to_bytes(from_bytes(image)) != image"
        if not len(arr):
            continue
        sel = pa.array(np.random.default_rng(ci).random(len(arr)) < 0.4)
        assert_arrays_equal(entry.filter(sel), o.filter(sel), f"{typ} case {ci}: filter")
        text = pa.types.is_string(arr.type) or pa.types.is_string_view(arr.type) or (
            pa.types.is_dictionary(arr.type) and pa.types.is_string(arr.type.value_type))
        if not text:
            continue
        needle = next((v for v in vals if v is not None), "x")
        for op in ("=", "!=", "<", ">="):
            expr = LiquidExpr.new_unchecked(BinaryExpr(Column("c", 0), op, Literal(needle)))
            assert_masks_equal(entry.try_eval_predicate(expr, sel), o.try_eval_predicate(op, needle, sel), f"{typ} case {ci}: {op} {needle!r}")
        # the image carries fingerprints but not the substring filter this build adds on insert: LIKE must still agree
        like = LiquidExpr.new_unchecked(LikeExpr(False, False, Column("c", 0), Literal("%goog%")))
        assert_masks_equal(entry.try_eval_predicate(like, sel), o.try_eval_predicate("like", "%goog%", sel), f"{typ} case {ci}: LIKE after import")


def test_round_trip_through_a_second_cache(cache):
    from liquid_cache_b200 import LiquidCacheBuilder

    arr = _build(_cases()[5], pa.string())
    liquid = cache.transcode(arr, hint=CacheExpression.SubstringSearch, compressor_scope=42)
    image, table = liquid.to_bytes(), cache.save_symbol_table(42)
    other = LiquidCacheBuilder.new().build()
    other.load_symbol_table(7, table)
    entry = other.read_from_bytes(image, compressor_scope=7)
    assert_arrays_equal(entry.to_arrow_array(), arr, "second cache")
    assert entry.to_bytes() == image
    like = LiquidExpr.new_unchecked(LikeExpr(False, False, Column("c", 0), Literal("%google%")))
    sel = pa.array([True] * len(arr))
    assert_masks_equal(entry.try_eval_predicate(like, sel), liquid.try_eval_predicate(like, sel), "LIKE, imported vs original")
    # later inserts under the loaded scope compress with the loaded table and stay readable
    more = other.transcode(_build(_cases()[6], pa.string()), compressor_scope=7)
    assert_arrays_equal(more.to_arrow_array(), _build(_cases()[6], pa.string()), "insert under a loaded table")


def test_damaged_byte_view_images_are_refused(cache):
    arr = _build(_cases()[5], pa.string())
    o = O.OracleByteViewArray.from_arrow(arr, build_fingerprints=True)
    good = O.byte_view_to_bytes(o)
    cache.load_symbol_table(3000, O.save_symbol_table(o.fsst))
    with pytest.raises(N.NativeError):
        cache.read_from_bytes(good)                                   # no symbol table named
    with pytest.raises(N.NativeError):
        cache.read_from_bytes(good, compressor_scope=3999)            # a scope that has none
    with pytest.raises(N.NativeError):
        cache.load_symbol_table(3000, O.save_symbol_table(o.fsst))    # the scope is taken
    with pytest.raises(N.NativeError):
        cache.load_symbol_table(3001, bytes([3, 1, 9, 1]) + bytes(24))  # symbol length 9
    keys_size, co_size, sp_size, fsst_size, fp_size = (int.from_bytes(good[16 + 4 * i:20 + 4 * i], "little") for i in range(5))
    keys_start = ((40 + fsst_size) + 7) & ~7
    bad_key = bytearray(good)
    vals_off = keys_start + 16  # no nulls: the packed keys follow the 16-byte BitPackedArray header
    bad_key[vals_off:vals_off + 2] = (65535).to_bytes(2, "little")    # a key past the dictionary
    bad_width = bytearray(good)
    bad_width[keys_start + 4] = 9
    bad_sizes = bytearray(good)
    bad_sizes[28:32] = (len(good) * 2).to_bytes(4, "little")          # FSST section longer than the image
    bad_offs = bytearray(good)
    co_start = ((keys_start + keys_size) + 7) & ~7
    bad_offs[co_start + 8] = 3                                        # residual width 3
    bad_resid = bytearray(good)
    ob = good[co_start + 8]
    last = co_start + 9 + (len(o.uniques)) * ob   # the closing offset: pushed past the compressed values
    bad_resid[last:last + ob] = (2 ** (8 * ob - 1) - 1).to_bytes(ob, "little")
    for bad in (good[:30], bytes(bad_key), bytes(bad_width), bytes(bad_sizes), bytes(bad_offs), bytes(bad_resid), good[:-(fp_size // 2)]):
        with pytest.raises(N.NativeError):
            cache.read_from_bytes(bad, compressor_scope=3000)
    # nothing leaked: the good image still loads afterwards
    assert_arrays_equal(cache.read_from_bytes(good, compressor_scope=3000).to_arrow_array(), arr, "after the refusals")
