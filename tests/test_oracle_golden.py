"""Pins the CPU oracle against the reference's own known-answer tests and checks the C port
(oracle/c, the timed CPU baseline) against the Python oracle. No GPU needed."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle.liquid_oracle import (CompactOffsets, OracleByteViewArray, OracleIntArray, boolean_buffer_and_then,
                                  fit_line, get_bit_width, string_fingerprint, substring_pattern_bytes, transcode)
from tests import golden_cases as G


def test_quick_start_known_answers():
    q = G.QUICK_START
    arr = pa.array(q["values"], pa.uint64())
    o = OracleIntArray.from_arrow(arr)
    assert o.bit_width == 3 and o.reference == 10  # UInt64 10..15 -> W = 3
    assert o.to_arrow().equals(arr)
    assert o.filter(pa.array(q["selection"])).to_pylist() == q["filtered"]
    assert o.try_eval_predicate(">", 12, pa.array([True] * 6)).to_pylist() == q["gt12"]
    s = OracleByteViewArray.from_arrow(pa.array(q["strings"]))
    assert s.try_eval_predicate("=", "apple", pa.array(q["string_selection"])).to_pylist() == q["eq_apple_selected"]


@pytest.mark.parametrize("name,values,checks", G.BYTE_VIEW_CASES, ids=[c[0] for c in G.BYTE_VIEW_CASES])
def test_byte_view_known_answers(name, values, checks):
    arr = pa.array(values, pa.string())
    o = OracleByteViewArray.from_arrow(arr)
    assert o.to_arrow().equals(arr)
    for op, needle, expected in checks:
        got = o.try_eval_predicate(op, needle, None).to_pylist()
        assert got == expected, f"{name}: {op} {needle!r}"


@pytest.mark.parametrize("name,values,checks", G.FINGERPRINT_CASES + G.LIKE_FALLBACK_CASES)
def test_like_known_answers(name, values, checks):
    arr = pa.array(values, pa.string())
    o = OracleByteViewArray.from_arrow(arr, build_fingerprints=name.startswith("fingerprint"))
    for op, needle, expected in checks:
        assert o.try_eval_predicate(op, needle, None).to_pylist() == expected, f"{name}: {op} {needle!r}"


@pytest.mark.parametrize("values,shared,keys", G.PREFIX_KEY_CASES)
def test_prefix_keys_known_answers(values, shared, keys):
    o = OracleByteViewArray.from_arrow(pa.array(values))
    assert o.shared_prefix == shared
    assert [k[0] for k in o.prefix_keys] == keys


def test_and_then_known_answer():
    l, r, want = G.AND_THEN_CASE
    got = boolean_buffer_and_then(pa.array([c == "1" for c in l]), pa.array([c == "1" for c in r]))
    assert "".join("1" if x else "0" for x in got.to_pylist()) == want


def test_bit_width_and_fingerprint_rules():
    assert [get_bit_width(x) for x in (0, 1, 2, 3, 255, 256, 2**64 - 1)] == [1, 1, 2, 2, 8, 9, 64]
    assert string_fingerprint(b"") == 0
    assert string_fingerprint(b"a") == 1 << (ord("a") & 31)
    assert substring_pattern_bytes(b"%google%") == b"google"
    for bad in (b"%", b"%%", b"abc", b"%a%b%", b"%a_b%", b"a%"):
        assert substring_pattern_bytes(bad) is None


def test_compact_offsets_round_trip():
    """raw/fsst_buffer.rs:931-989: offsets survive slope/intercept + residuals in 1, 2 and 4 bytes."""
    rng = np.random.default_rng(3)
    for offs in ([0], [0, 5], list(range(0, 5000, 10)), [0, 100, 100, 105, 70000, 70001],
                 list(np.cumsum(rng.integers(0, 300, size=2000))), list(np.cumsum(rng.integers(0, 100000, size=300)))):
        offs = [int(x) for x in offs]
        c = CompactOffsets.from_offsets(offs)
        assert [c.get_offset(i) for i in range(len(offs))] == offs
        assert c.offset_bytes in (1, 2, 4)
    assert fit_line([7]) == (0, 7)
    assert fit_line([0, 10, 20, 30]) == (10, 0)


def test_transcode_dtype_matrix():
    """cache/transcode.rs:301-436: 8192-row arrays per type round-trip; Boolean (and floats here) stay Arrow."""
    n = 8192
    cases = [pa.array(np.arange(n, dtype=np.int32)), pa.array(np.arange(n, dtype=np.int64)),
             pa.array(np.arange(n, dtype=np.int64) * 1000, pa.timestamp("us")),
             pa.array([f"test_string_{i}" for i in range(n)]),
             pa.array([f"test_string_{i}".encode() for i in range(n)], pa.binary_view()),
             pa.array([f"value_{i % 100}" for i in range(n)]).dictionary_encode().cast(pa.dictionary(pa.uint16(), pa.string()))]
    for arr in cases:
        o = transcode(arr)
        assert o is not None
        back = o.to_arrow()
        assert back.type == arr.type
        assert back.to_pylist() == arr.to_pylist()
    assert transcode(pa.array([True, False] * 10)) is None


def test_differential_spec_against_arrow():
    """fuzz/fuzz_targets/fsst_view.rs:86-117: compare_with == arrow cmp for arbitrary strings."""
    rng = np.random.default_rng(9)
    alphabet = ["a", "b", "ab", "é", "", "zz", "http://", "\x00", "\xff"]
    for trial in range(40):
        n = int(rng.integers(1, 60))
        vals = [None if rng.random() < 0.15 else "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), size=int(rng.integers(0, 6))))
                for _ in range(n)]
        arr = pa.array(vals, pa.string())
        o = OracleByteViewArray.from_arrow(arr)
        assert o.to_arrow().equals(arr)
        needle = vals[int(rng.integers(0, n))] or "a"
        for op, fn in (("=", pc.equal), ("!=", pc.not_equal), ("<", pc.less), ("<=", pc.less_equal), (">", pc.greater), (">=", pc.greater_equal)):
            assert o.try_eval_predicate(op, needle, None).to_pylist() == fn(arr, pa.scalar(needle)).to_pylist()


def test_differential_spec_hypothesis():
    """Same spec, searched by hypothesis (the reference drives it with libfuzzer): arbitrary unicode / control bytes,
    nulls, a selection, every comparison operator and LIKE '%x%' with and without fingerprints."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    text = st.text(alphabet=st.sampled_from(list("abgoleGOLE/%._-:?=&0189 \x00\x7f") + ["é", "к", "\U0001F600"]), max_size=14)

    @settings(max_examples=150, deadline=None)
    @given(vals=st.lists(st.one_of(st.none(), text), min_size=1, max_size=40), needle=text, data=st.data())
    def run(vals, needle, data):
        arr = pa.array(vals, pa.string())
        sel_bits = data.draw(st.lists(st.booleans(), min_size=len(vals), max_size=len(vals)))
        sel = pa.array(sel_bits)
        filt = arr.filter(sel)
        for fp in (False, True):
            o = OracleByteViewArray.from_arrow(arr, build_fingerprints=fp)
            assert o.to_arrow().equals(arr)
            assert o.filter(sel).equals(filt)
            for op, fn in (("=", pc.equal), ("!=", pc.not_equal), ("<", pc.less), ("<=", pc.less_equal), (">", pc.greater),
                           (">=", pc.greater_equal)):
                assert o.try_eval_predicate(op, needle, sel).to_pylist() == fn(filt, pa.scalar(needle)).to_pylist()
            inner = needle.replace("%", "").replace("_", "").replace("\\", "")
            if inner:
                got = o.try_eval_predicate("like", f"%{inner}%", sel).to_pylist()
                assert got == pc.match_substring(filt, inner).to_pylist()

    run()


def test_c_port_matches_python_oracle():
    """The timed CPU baseline (oracle/c) must agree with the oracle it is a port of."""
    from oracle import c_oracle as CO

    CO.lib(rebuild=True)
    rng = np.random.default_rng(21)
    for t, lo, hi in ((pa.int64(), -5000, 5000), (pa.int16(), -3, 12), (pa.uint32(), 0, 2**32 - 1), (pa.int8(), -128, 127)):
        vals = rng.integers(lo, hi, size=3000, endpoint=True)
        arr = pa.array(vals, t, mask=rng.random(3000) < 0.1)
        c, o = CO.CIntArray(arr), OracleIntArray.from_arrow(arr)
        sel = pa.array(rng.random(3000) < 0.4)
        assert c.filter(None).equals(arr) and c.filter(sel).equals(o.filter(sel))
        lit = int(vals[5])
        for op in ("=", "!=", "<", "<=", ">", ">="):
            assert c.eval(op, lit, sel).to_pylist() == o.try_eval_predicate(op, lit, sel).to_pylist()
    words = ["http://", "google", ".ru", "/", "maps", "yandex", "?q=", "%D0%BA", "x" * 300]
    vals = ["".join(words[int(i)] for i in rng.integers(0, len(words), size=int(rng.integers(0, 9)))) for _ in range(2000)]
    arr = pa.array(vals, mask=rng.random(2000) < 0.1)
    fsst = CO.CFsst(arr)
    c = CO.CStrArray(arr, fsst, build_fingerprints=True)
    o = OracleByteViewArray.from_arrow(arr, build_fingerprints=True)
    sel = pa.array(rng.random(2000) < 0.5)
    assert c.filter(None).equals(arr) and c.filter(sel).equals(o.filter(sel))
    for inner in (b"google", b"maps", b"zzz", b"D0"):
        for neg in (False, True):
            want = o.try_eval_predicate("not like" if neg else "like", b"%" + inner + b"%", sel)
            assert c.like(inner, sel, neg).to_pylist() == want.to_pylist()
    for needle in (vals[0].encode(), b"", b"http://google", vals[3].encode()):
        for neg in (False, True):
            assert c.eq(needle, sel, neg).to_pylist() == o.try_eval_predicate("!=" if neg else "=", needle, sel).to_pylist()
    l = rng.random(5000) < 0.3
    r = rng.random(int(l.sum())) < 0.5
    assert CO.and_then(l, r).tolist() == boolean_buffer_and_then(pa.array(l), pa.array(r)).to_pylist()


def test_multi_column_or_known_answers():
    """src/datafusion/src/cache/mod.rs:433-639: per-column masks under one selection, joined with or_kleene"""
    from oracle.liquid_oracle import evaluate_multi_column_or, transcode
    from tests.golden_cases import MULTI_COLUMN_OR_CASES

    types = {"int32": pa.int32(), "string_view": pa.string_view()}
    for cols, conjuncts, want_rows in MULTI_COLUMN_OR_CASES:
        n = len(cols[0][1])
        liquid = [transcode(pa.array(vals, types[t])) for t, vals in cols]
        got = evaluate_multi_column_or([(la, op, lit) for la, (op, lit) in zip(liquid, conjuncts)], pa.array([True] * n))
        assert got.to_pylist() == [i in want_rows for i in range(n)]
    # nulls follow Kleene logic: NULL OR TRUE = TRUE, NULL OR FALSE = NULL
    a = transcode(pa.array([1, None, None, 4], pa.int32()))
    b = transcode(pa.array([10, 20, 30, None], pa.int32()))
    got = evaluate_multi_column_or([(a, "=", 1), (b, "=", 20)], pa.array([True] * 4))
    assert got.to_pylist() == [True, True, None, None]
