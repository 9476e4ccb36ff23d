"""Multi-column OR over cached columns, the way the reference's caller evaluates it (CachedRowGroup::
evaluate_selection_with_predicate, /root/reference/src/datafusion/src/cache/mod.rs:111-150): every column evaluates its
own conjunct on the encoded data under the same selection, the masks are joined with arrow's or_kleene. The three
known-answer tests of the reference (:433-639) through the device entries, plus Kleene logic under nulls."""
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from liquid_cache_b200 import BinaryExpr, Column, EntryID, LiquidExpr, Literal, parquet_array_id
from oracle import liquid_oracle as O
from tests.golden_cases import MULTI_COLUMN_OR_CASES
from tests.util import assert_masks_equal

pytestmark = pytest.mark.gpu

TYPES = {"int32": pa.int32(), "string_view": pa.string_view()}


def _or_over_columns(cache, ids, conjuncts, selection):
    combined = None
    for col, (eid, (op, lit)) in enumerate(zip(ids, conjuncts)):
        expr = LiquidExpr.new_unchecked(BinaryExpr(Column(f"c{col}", col), op, Literal(lit)))
        mask = cache.eval_predicate(eid, expr).with_selection(selection).read()
        combined = mask if combined is None else pc.or_kleene(combined, mask)
    return combined


def test_reference_known_answers(cache):
    for case, (cols, conjuncts, want_rows) in enumerate(MULTI_COLUMN_OR_CASES):
        n = len(cols[0][1])
        ids = [EntryID(parquet_array_id(400 + case, 0, c, 0)) for c in range(len(cols))]
        arrays = [pa.array(vals, TYPES[t]) for t, vals in cols]
        for eid, arr in zip(ids, arrays):
            cache.insert(eid, arr).run()
        sel = pa.array([True] * n)
        got = _or_over_columns(cache, ids, conjuncts, sel)
        assert got.to_pylist() == [i in want_rows for i in range(n)], case
        want = O.evaluate_multi_column_or([(O.transcode(a), op, lit) for a, (op, lit) in zip(arrays, conjuncts)], sel)
        assert_masks_equal(got, want, f"case {case} vs the restatement")
        half = pa.array([i % 2 == 1 for i in range(n)])  # a selection shared by all columns
        got = _or_over_columns(cache, ids, conjuncts, half)
        want = O.evaluate_multi_column_or([(O.transcode(a), op, lit) for a, (op, lit) in zip(arrays, conjuncts)], half)
        assert_masks_equal(got, want, f"case {case} under a selection")


def test_kleene_logic_under_nulls(cache):
    a, b = pa.array([1, None, None, 4], pa.int32()), pa.array([10, 20, 30, None], pa.int32())
    ids = [EntryID(parquet_array_id(410, 0, c, 0)) for c in range(2)]
    for eid, arr in zip(ids, (a, b)):
        cache.insert(eid, arr).run()
    got = _or_over_columns(cache, ids, [("=", 1), ("=", 20)], pa.array([True] * 4))
    assert got.to_pylist() == [True, True, None, None]  # NULL OR TRUE = TRUE, NULL OR FALSE = NULL
