"""GPU parity against the reference's own known-answer tests (the same table that pins the oracle)."""
import pyarrow as pa
import pytest

from tests import golden_cases as G

pytestmark = pytest.mark.gpu


def _expr(op, needle):
    from liquid_cache_b200 import BinaryExpr, Column, LikeExpr, LiquidExpr, Literal

    if op in ("like", "not like"):
        return LiquidExpr.new_unchecked(LikeExpr(op == "not like", False, Column("c", 0), Literal(needle)))
    return LiquidExpr.new_unchecked(BinaryExpr(Column("c", 0), op, Literal(needle)))


@pytest.mark.parametrize("name,values,checks", G.BYTE_VIEW_CASES, ids=[c[0] for c in G.BYTE_VIEW_CASES])
def test_byte_view_known_answers(cache, name, values, checks):
    arr = pa.array(values, pa.string())
    liquid = cache.transcode(arr, compressor_scope=hash(name) & 0xFFFFFFFF)
    assert liquid.to_arrow_array().equals(arr)
    sel = pa.array([True] * len(values))
    for op, needle, expected in checks:
        assert liquid.try_eval_predicate(_expr(op, needle), sel).to_pylist() == expected, f"{name}: {op} {needle!r}"


@pytest.mark.parametrize("name,values,checks", G.FINGERPRINT_CASES)
def test_fingerprint_known_answers(cache, name, values, checks):
    from liquid_cache_b200 import CacheExpression

    liquid = cache.transcode(pa.array(values), hint=CacheExpression.SubstringSearch, compressor_scope=777)
    sel = pa.array([True] * len(values))
    for op, needle, expected in checks:
        assert liquid.try_eval_predicate(_expr(op, needle), sel).to_pylist() == expected


def test_quick_start_through_the_cache_api(cache):
    from liquid_cache_b200 import EntryID

    q = G.QUICK_START
    cache.insert(EntryID(9001), pa.array(q["values"], pa.uint64())).run()
    assert cache.get(EntryID(9001)).with_selection(pa.array(q["selection"])).read().to_pylist() == q["filtered"]
    assert cache.eval_predicate(EntryID(9001), _expr(">", 12)).read().to_pylist() == q["gt12"]
    cache.insert(EntryID(9002), pa.array(q["strings"])).run()
    got = cache.eval_predicate(EntryID(9002), _expr("=", "apple")).with_selection(pa.array(q["string_selection"])).read()
    assert got.to_pylist() == q["eq_apple_selected"]


def test_and_then_known_answer(cache):
    l, r, want = G.AND_THEN_CASE
    got = cache.and_then(pa.array([c == "1" for c in l]), pa.array([c == "1" for c in r]))
    assert "".join("1" if x else "0" for x in got.to_pylist()) == want
