"""Predicate lowering on the host (liquid_cache_b200/expr.py), no device needed: which column-side cast chains keep the
integer value (the reference admits every cast chain in `is_column_like`, cache/liquid_expr.rs:150-164, and then evaluates
the REAL cast with DataFusion on the decoded array, primitive_array.rs:376-379 — so only chains that are the identity on
every value of the source type may be answered in the packed domain) and in which unit a date literal is compared."""
import datetime as dt

import pyarrow as pa
import pytest

from liquid_cache_b200 import BinaryExpr, CastExpr, Column, LiquidExpr, Literal
from liquid_cache_b200 import _native as N
from liquid_cache_b200.expr import _cast_chain_is_integer_identity, _int_literal

COL = Column("c", 0)


def _native(left, op, value, column_type):
    return LiquidExpr.new_unchecked(BinaryExpr(left, op, Literal(value))).to_native(column_type)


def test_widening_chains_are_the_identity():
    # ClickBench: "EventDate"::INT::DATE over a UInt16 column
    chain = CastExpr(CastExpr(COL, pa.int32()), pa.date32())
    assert _cast_chain_is_integer_identity(chain, pa.uint16())
    p = _native(chain, ">=", dt.date(2013, 7, 1), pa.uint16())
    assert p.lit_kind == N.LIT_I64 and p.lit_i64 == (dt.date(2013, 7, 1) - dt.date(1970, 1, 1)).days
    assert _cast_chain_is_integer_identity(CastExpr(COL, pa.int64()), pa.int32())
    assert _cast_chain_is_integer_identity(CastExpr(COL, pa.int64()), pa.uint32())
    assert _cast_chain_is_integer_identity(CastExpr(COL, pa.int32()), pa.date32())


@pytest.mark.parametrize("column_type, target", [
    (pa.int64(), pa.int8()),     # narrowing
    (pa.uint64(), pa.int64()),   # sign change: values above i64::MAX do not fit
    (pa.int32(), pa.uint32()),   # sign change: negatives do not fit
    (pa.int64(), pa.date32()),   # narrowing to a 32-bit day count
    (pa.date64(), pa.date32()),  # rescales (ms -> days)
    (pa.timestamp("us"), pa.int64()),  # a timestamp under a cast: left to DataFusion
])
def test_narrowing_sign_changing_and_rescaling_casts_are_declined(column_type, target):
    chain = CastExpr(COL, target)
    assert not _cast_chain_is_integer_identity(chain, column_type)
    with pytest.raises(N.UnsupportedExpr):
        _native(chain, ">", 3, column_type)


def test_a_chain_that_narrows_in_the_middle_is_declined():
    chain = CastExpr(CastExpr(COL, pa.int8()), pa.int64())
    assert not _cast_chain_is_integer_identity(chain, pa.int32())


def test_date_literal_unit_follows_the_compared_type():
    d = dt.date(2020, 1, 1)
    days = (d - dt.date(1970, 1, 1)).days
    assert _int_literal(Literal(d), pa.date32()) == days
    assert _int_literal(Literal(d), pa.date64()) == days * 86_400_000
    assert _int_literal(Literal(d), pa.timestamp("us")) is None
    assert _native(COL, ">=", d, pa.date64()).lit_i64 == days * 86_400_000
    assert _native(COL, ">=", d, pa.date32()).lit_i64 == days
    with pytest.raises(N.UnsupportedExpr):
        _native(COL, ">=", d, pa.timestamp("us"))
