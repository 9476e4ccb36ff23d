"""GPU parity of LiquidArray::to_bytes / ipc::read_from_bytes (LQDA) for Integer / Float / Decimal entries.

Reference: liquid_array/ipc.rs:158-283 and its tests :308-418; primitive_array.rs:603-679; float_array.rs:393-600;
decimal_array.rs:180-251; raw/bit_pack_array.rs:181-334 (all under /root/reference/src/core/src).
Checked: the image the device entry serializes to is byte for byte the oracle's restatement of the format; an image the
oracle wrote becomes an entry that reads back as the original array; to_bytes -> from_bytes is the identity on the
HBM image; damaged images are refused (the reference panics on them).
"""
import decimal

import numpy as np
import pyarrow as pa
import pytest

from oracle import liquid_oracle as O
from tests.util import assert_arrays_equal, assert_float_bits_equal

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(31)
    out = [pa.array([10, 20, 30, None, 50], pa.int32()),                        # ipc.rs:308-360
           pa.array([None] * 1000, pa.int32()), pa.array(list(range(1000)), pa.int32()), pa.array([42], pa.int32()),
           pa.array([], pa.int32()), pa.array([None if i in (1000, 5000, 9000) else i for i in range(10000)], pa.int32())]
    for typ, lo, hi in [(pa.int8(), -128, 127), (pa.uint16(), 100, 160), (pa.int64(), -(2**63), 2**63 - 1), (pa.uint64(), 0, 2**64 - 1),
                        (pa.int64(), 1373832014, 1373832014 + 86400)]:
        vals = rng.integers(lo, hi, size=8192, endpoint=True, dtype=np.int64 if hi < 2**63 else np.uint64)
        out.append(pa.array(vals, type=typ, mask=rng.random(8192) < 0.1))
    out.append(pa.array(rng.integers(8036, 10556, size=3000), pa.int32()).cast(pa.date32()))
    out.append(pa.array(rng.integers(0, 10**12, size=2000), pa.int64()).cast(pa.timestamp("ms")))
    for typ in (pa.float32(), pa.float64()):
        np_dt = typ.to_pandas_dtype()
        x = np.round(rng.uniform(-100, 100, size=8192), 1).astype(np_dt)
        x[rng.integers(0, 8192, size=60)] = rng.standard_normal(60).astype(np_dt)  # patches
        x[5] = np.nan
        out.append(pa.array(x, typ, mask=rng.random(8192) < 0.05))
        out.append(pa.array(np.arange(2000).astype(np_dt), typ))
        out.append(pa.array([None] * 7, typ))
        out.append(pa.array(rng.standard_normal(600).astype(np_dt), typ))  # every row a patch
    for typ in (pa.decimal128(15, 2), pa.decimal256(50, 4)):
        ints = rng.integers(0, 10**9, size=5000)
        with decimal.localcontext() as cx:
            cx.prec = 100
            vals = [None if i % 11 == 0 else decimal.Decimal(int(v)).scaleb(-typ.scale) for i, v in enumerate(ints)]
        out.append(pa.array(vals, typ))
    return out


def _oracle(arr):
    return O.transcode(arr)


def _same(got, want, what):
    if pa.types.is_floating(want.type):
        assert_float_bits_equal(got, want, what)
    else:
        assert_arrays_equal(got, want, what)


@pytest.mark.parametrize("idx", range(len(_cases())))
def test_to_bytes_and_read_from_bytes(cache, idx):
    arr = _cases()[idx]
    oracle = _oracle(arr)
    liquid = cache.transcode(arr)
    image = liquid.to_bytes()
    want = O.to_bytes(oracle)
    assert len(image) == len(want), f"{arr.type} n={len(arr)}: LQDA image of {len(image)} bytes, the oracle's has {len(want)}"
    if arr.null_count == 0 or pa.types.is_floating(arr.type):
        assert image == want, f"{arr.type} n={len(arr)}: LQDA image differs from the oracle's"
    else:
        # integer / decimal entries keep zeros in the packed slots of null rows where the reference keeps whatever the
        # Arrow buffer held (unspecified payload): everything up to the packed words must agree, and the words must decode alike
        parsed = O.read_from_bytes(image)
        ints_w = oracle.ints if isinstance(oracle, O.OracleDecimalArray) else oracle
        ints_g = parsed.ints if isinstance(parsed, O.OracleDecimalArray) else parsed
        assert (ints_g.n, ints_g.bit_width, ints_g.reference) == (ints_w.n, ints_w.bit_width, ints_w.reference)
        packed_bytes = 0 if ints_w.bit_width is None else ints_w.packed.nbytes
        head = len(want) - packed_bytes
        assert image[:head] == want[:head], "headers / null bitmap differ"
        _same(parsed.to_arrow(), oracle.to_arrow(), "decoded image")
    back = cache.read_from_bytes(want)  # an image written by the CPU restatement
    assert back.len() == len(arr) and back.data_type() == liquid.data_type()
    assert len(arr) == 0 or back.original_arrow_data_type() == arr.type
    _same(back.to_arrow_array(), oracle.to_arrow(), "read_from_bytes(oracle image)")
    again = cache.read_from_bytes(image)  # and our own
    assert again.to_bytes() == image
    if len(arr):
        sel = pa.array(np.random.default_rng(idx).random(len(arr)) < 0.3)
        _same(again.filter(sel), oracle.filter(sel), "filter after the round trip")


def test_damaged_images_are_refused(cache):
    from liquid_cache_b200 import _native as N

    good = O.to_bytes(O.OracleIntArray.from_arrow(pa.array(list(range(5000)), pa.int32())))
    for bad in (good[:10], b"XXXX" + good[4:], good[:4] + (2).to_bytes(2, "little") + good[6:],  # short, magic, version
                good[:28] + bytes([40]) + good[29:],                                            # bit width 40 on Int32
                good[:-100]):                                                                    # values cut short
        with pytest.raises(N.NativeError):
            cache.read_from_bytes(bad)
    f = O.OracleFloatArray.from_arrow(pa.array([0.5, float("nan"), 1.5, 2.5]))
    img = bytearray(O.to_bytes(f))
    assert len(f.patch_indices) >= 1
    img[40:48] = (10**6).to_bytes(8, "little")  # first patch index far past the 4 rows
    with pytest.raises(N.NativeError):
        cache.read_from_bytes(bytes(img))
    s = cache.transcode(pa.array(["a", "b"]))
    with pytest.raises(N.NativeError):  # a byte-view image without its symbol table (tests/test_gpu_zy_ipc_strings.py has the rest)
        cache.read_from_bytes(s.to_bytes())
