"""GPU parity of the batched insert (lc_cache_insert_many): every entry it makes must be the entry the one-batch
insert makes — same HBM image byte for byte — and read back as the Arrow batch it was given.

Reference behaviour being batched: LiquidCache::insert per batch (src/core/src/cache/core.rs:122-128) ->
LiquidPrimitiveArray::from_arrow_array (liquid_array/primitive_array.rs:159-206).
"""
import numpy as np
import pyarrow as pa
import pytest

from oracle.liquid_oracle import OracleIntArray
from tests.util import assert_arrays_equal, assert_masks_equal, random_selection

pytestmark = pytest.mark.gpu


def _expr(op, value):
    from liquid_cache_b200 import BinaryExpr, Column, LiquidExpr, Literal

    return LiquidExpr.new_unchecked(BinaryExpr(Column("c", 0), op, Literal(value)))


def _int_batches(rng):
    types = [(pa.int8(), -128, 127), (pa.int16(), -2000, 31000), (pa.int32(), -(2**31), 2**31 - 1), (pa.int64(), -(2**63), 2**63 - 1),
             (pa.uint8(), 0, 255), (pa.uint16(), 100, 160), (pa.uint32(), 0, 2**32 - 1), (pa.uint64(), 0, 2**64 - 1),
             (pa.int64(), 1373832014, 1373832014 + 86400), (pa.int32(), 7, 7), (pa.date32(), 8036, 10556),
             (pa.timestamp("us"), 0, 10**15), (pa.date64(), 0, 86400000 * 20000)]
    out = []
    for ti, (typ, lo, hi) in enumerate(types):
        for n in (1, 500, 8192, 10000):
            if hi < 2**63:
                vals = rng.integers(lo, hi, size=n, endpoint=True, dtype=np.int64)
            else:
                vals = rng.integers(lo, hi, size=n, endpoint=True, dtype=np.uint64)
            mask = (rng.random(n) < 0.2) if (ti + n) % 3 == 0 else None
            store = pa.int32() if pa.types.is_date32(typ) else (pa.int64() if (pa.types.is_timestamp(typ) or pa.types.is_date64(typ)) else typ)
            arr = pa.array(vals, type=store, mask=mask)
            if pa.types.is_date64(typ):
                arr = pa.array((vals // 86400000) * 86400000, type=pa.int64(), mask=mask)
            out.append(arr.cast(typ) if store != typ else arr)
    out.append(pa.array([None] * 300, pa.int32()))   # entirely null
    out.append(pa.array([], pa.int64()))              # empty
    out.append(pa.array([None if i % 3 == 0 else i for i in range(5000)], pa.int32()).slice(13, 3001))  # sliced, bit offset
    return out


def test_insert_many_integers_equal_single_inserts(cache):
    from liquid_cache_b200 import EntryID

    rng = np.random.default_rng(2024)
    arrays = _int_batches(rng)
    ids = [EntryID((9 << 48) | i) for i in range(len(arrays))]
    cache.insert_many(ids, arrays)
    for eid, arr in zip(ids, arrays):
        assert cache.is_cached(eid)
        liquid = cache.try_read_liquid(eid)
        single = cache.transcode(arr)
        assert liquid.entry_image() == single.entry_image(), f"{arr.type} n={len(arr)}: HBM image differs from the one-batch insert"
        assert_arrays_equal(cache.get(eid).read(), arr, f"{arr.type} n={len(arr)}")
        assert liquid.original_arrow_data_type() == arr.type or len(arr) == 0
    # predicates and filters over the batched entries, against the oracle
    for eid, arr in list(zip(ids, arrays))[::5]:
        if len(arr) == 0:
            continue
        oracle = OracleIntArray.from_arrow(arr)
        sel = random_selection(rng, len(arr), 0.5)
        assert_arrays_equal(cache.get(eid).with_selection(sel).read(), oracle.filter(sel), "filter")
        valid = [v for v in arr.to_pylist() if v is not None]
        if not valid or not (pa.types.is_integer(arr.type) or pa.types.is_date32(arr.type)):
            continue  # the mirror lowers int and date literals; timestamps are covered by the round trip above
        lit = valid[0]
        got = cache.eval_predicate(eid, _expr("<=", lit)).with_selection(sel).read()
        assert_masks_equal(got, oracle.try_eval_predicate("<=", lit, sel), f"{arr.type} <= {lit}")


def test_insert_many_overwrites_and_is_all_or_nothing(cache):
    from liquid_cache_b200 import EntryID, _native as N

    a1 = pa.array(np.arange(100), pa.int64())
    a2 = pa.array(np.arange(100, 300), pa.int32())
    ids = [EntryID((10 << 48) | 1), EntryID((10 << 48) | 2)]
    cache.insert_many(ids, [a1, a2])
    cache.insert_many(ids, [a2, a1])  # index insert replaces
    assert_arrays_equal(cache.get(ids[0]).read(), a2, "overwritten 0")
    assert_arrays_equal(cache.get(ids[1]).read(), a1, "overwritten 1")
    fresh = EntryID((10 << 48) | 3)
    with pytest.raises(N.UnsupportedType):
        cache.insert_many([fresh, ids[0]], [a1, pa.array([True, False])])  # Boolean: transcode declines -> nothing inserted
    assert not cache.is_cached(fresh)
    assert_arrays_equal(cache.get(ids[0]).read(), a2, "untouched after the failed call")


def test_insert_many_mixed_kinds_take_the_per_batch_path(cache):
    from liquid_cache_b200 import EntryID

    arrays = [pa.array(["a", "bb", None, "a"]), pa.array([1, 2, 3], pa.int16()), pa.array([1.5, None, 2.25]),
              pa.array(np.arange(9000), pa.uint32())]
    ids = [EntryID((11 << 48) | i) for i in range(len(arrays))]
    cache.insert_many(ids, arrays)
    for eid, arr in zip(ids, arrays):
        assert_arrays_equal(cache.get(eid).read(), arr, str(arr.type))


@pytest.mark.parametrize("typ", [pa.string(), pa.binary(), pa.string_view(), pa.dictionary(pa.uint16(), pa.string())], ids=str)
@pytest.mark.parametrize("hinted", [False, True])
def test_insert_many_byte_views_equal_single_inserts(cache, typ, hinted):
    """The batched byte-view insert (five encode stages over the whole list) must produce the entry the one-batch insert
    produces under the same symbol table: same HBM image, same answers."""
    import zlib

    from liquid_cache_b200 import CacheExpression, EntryID
    from oracle.liquid_oracle import OracleByteViewArray
    from tests.test_gpu_str import build, make_strings

    tag = zlib.crc32(repr((str(typ), hinted)).encode())
    rng = np.random.default_rng(tag)
    hint = CacheExpression.SubstringSearch if hinted else None
    file_id = 12 + tag % 997
    arrays, ids = [], []
    for rg in range(2):           # two column chunks = two symbol tables, each trained on its chunk's first batch
        for b, n in enumerate((8192, 1, 500, 8192, 3000)):
            vals, mask = make_strings(rng, n, max(1, n // 4), null_p=0.1 if b % 2 else 0.0)
            arrays.append(build(vals, mask, typ))
            ids.append(EntryID((file_id << 48) | (rg << 32) | (5 << 16) | b))
    arrays.append(build(["x"] * 40, np.ones(40, dtype=bool), typ))  # entirely null
    ids.append(EntryID((file_id << 48) | (1 << 32) | (5 << 16) | 9))
    cache.insert_many(ids, arrays, hint=hint)
    for eid, arr in zip(ids, arrays):
        liquid = cache.try_read_liquid(eid)
        single = cache.transcode(arr, hint=hint, compressor_scope=int(eid) & ~0xFFFF)  # the chunk's table exists by now
        assert liquid.entry_image() == single.entry_image(), f"{typ} n={len(arr)}: HBM image differs from the one-batch insert"
        assert_arrays_equal(cache.get(eid).read(), arr, f"{typ} n={len(arr)}")
    eid, arr = ids[3], arrays[3]
    oracle = OracleByteViewArray.from_arrow(arr, build_fingerprints=hinted)
    sel = random_selection(rng, len(arr), 0.5)
    plain = arr.cast(arr.type.value_type) if pa.types.is_dictionary(arr.type) else arr
    needle = next(v for v in plain.to_pylist() if v is not None)
    got = cache.eval_predicate(eid, _expr("<=", needle)).with_selection(sel).read()
    assert_masks_equal(got, oracle.try_eval_predicate("<=", needle, sel), "<= after the batched insert")
