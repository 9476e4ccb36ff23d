"""The byte-view half of LQDA in the CPU oracle (byte_view_array/serialization.rs:87-325): round trips of every section
for the four original Arrow types, with and without fingerprints, including empty and entirely null arrays — the cases of
the reference's own serialization tests (byte_view_array/tests.rs round trips through to_bytes/from_bytes). The device
side of the same format is checked against this in tests/test_gpu_zy_ipc_strings.py."""
import pyarrow as pa
import pytest

from oracle import liquid_oracle as O

TYPES = [pa.string(), pa.binary(), pa.string_view(), pa.dictionary(pa.uint16(), pa.string())]
CASES = [["hello_world", "hello_rust", None, "hello_test", "hello_world"], [], [None, None], [f"http://x/{i % 700}" for i in range(5000)]]


def _build(vals, typ):
    if pa.types.is_dictionary(typ):
        return pa.array(vals, pa.string()).dictionary_encode().cast(typ)
    text = pa.types.is_string(typ) or pa.types.is_string_view(typ)
    return pa.array([None if v is None else (v if text else v.encode()) for v in vals], typ)


@pytest.mark.parametrize("typ", TYPES, ids=str)
@pytest.mark.parametrize("fp", [False, True])
def test_byte_view_lqda_round_trip(typ, fp):
    for vals in CASES:
        arr = _build(vals, typ)
        o = O.OracleByteViewArray.from_arrow(arr, build_fingerprints=fp)
        image = O.byte_view_to_bytes(o)
        assert image[0:4] == O.LQDA_MAGIC and int.from_bytes(image[6:8], "little") == 4  # LiquidDataType::ByteViewArray
        assert int.from_bytes(image[8:10], "little") == O._arrow_byte_type_id(arr.type)
        r = O.byte_view_from_bytes(image, o.fsst, arr.type)
        assert r.to_arrow().equals(o.to_arrow()), (typ, len(vals))
        assert r.keys == list(o.keys) and r.prefix_keys == o.prefix_keys and r.shared_prefix == o.shared_prefix
        assert (r.fingerprints or None) == (o.fingerprints or None)  # no fingerprints and zero of them serialize alike
        assert r.compressed == o.compressed and r.offsets.residuals == o.offsets.residuals
        sel = pa.array([i % 3 == 0 for i in range(len(arr))])
        if len(arr):
            needle = next((v for v in vals if v is not None), "x")
            assert r.try_eval_predicate("=", needle, sel).to_pylist() == o.try_eval_predicate("=", needle, sel).to_pylist()


def test_symbol_table_save_format():
    o = O.OracleByteViewArray.from_arrow(pa.array([f"http://host{i % 9}/path" for i in range(300)]))
    blob = O.save_symbol_table(o.fsst)
    n = blob[0]
    assert n == len(o.fsst.symbols) and len(blob) == 1 + n + 8 * n  # count, lengths, u64 symbols (fsst_buffer.rs:854-883)
    assert list(blob[1:1 + n]) == [len(s) for s in o.fsst.symbols]


def test_symbol_table_load_is_the_inverse_of_save():
    o = O.OracleByteViewArray.from_arrow(pa.array([f"https://shop{i % 31}.example/item?id={i}" for i in range(900)]))
    back = O.load_symbol_table(O.save_symbol_table(o.fsst))
    assert back.symbols == o.fsst.symbols
    assert back.decompress(o.fsst.compress(b"https://shop7.example/item?id=12")) == b"https://shop7.example/item?id=12"
