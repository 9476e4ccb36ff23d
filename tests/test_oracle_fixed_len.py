"""LiquidFixedLenByteArray in the CPU oracle (decimals with a value outside u64): the reference's tests are round trips
and filters over generated decimals (fix_len_byte_array.rs:452-598) plus the transcode dispatch (transcode.rs:118-153)."""
import decimal

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import liquid_oracle as O


def gen_decimals(typ, n, n_distinct, null_p, seed):
    """like liquid_array/utils.rs gen_test_decimal_array: values across the type's range, negatives included"""
    rng = np.random.default_rng(seed)
    digits = min(typ.precision, 60)
    with decimal.localcontext() as cx:
        cx.prec = 100
        pool = [decimal.Decimal(int(rng.integers(-(10**18), 10**18)) * 10 ** int(rng.integers(0, max(1, digits - 18)))).scaleb(-typ.scale) for _ in range(n_distinct)]
        vals = [None if rng.random() < null_p else pool[int(rng.integers(0, n_distinct))] for _ in range(n)]
    return pa.array(vals, typ)


@pytest.mark.parametrize("typ", [pa.decimal128(38, 6), pa.decimal128(20, 0), pa.decimal256(60, 10), pa.decimal256(76, 0)], ids=str)
def test_round_trip_filter_and_dispatch(typ):
    arr = gen_decimals(typ, 3000, 500, 0.1, typ.precision)
    assert not O.OracleDecimalArray.fits_u64(arr)
    o = O.transcode(arr)
    assert isinstance(o, O.OracleFixedLenByteArray) and len(o) == len(arr)
    assert len(o.uniques) <= 500 and all(len(u) == typ.byte_width for u in o.uniques) and o.key_bit_width == O.get_bit_width(len(o.uniques) - 1)
    assert o.to_arrow().equals(arr)
    sel = pa.array(np.random.default_rng(1).random(len(arr)) < 0.3)
    assert o.filter(sel).equals(pc.filter(arr, sel))
    lit = next(v for v in arr.to_pylist() if v is not None)
    got = o.try_eval_predicate(">=", pa.scalar(lit, typ), sel)
    assert got.equals(pc.greater_equal(pc.filter(arr, sel), pa.scalar(lit, typ)))


def test_decimals_that_fit_u64_keep_the_integer_form():
    arr = pa.array([decimal.Decimal("12.50"), None, decimal.Decimal("0.01")], pa.decimal128(15, 2))
    assert isinstance(O.transcode(arr), O.OracleDecimalArray)
    neg = pa.array([decimal.Decimal("-0.01"), decimal.Decimal("1.00")], pa.decimal128(15, 2))
    assert isinstance(O.transcode(neg), O.OracleFixedLenByteArray)  # a negative value is outside u64
    assert O.transcode(neg).to_arrow().equals(neg)
