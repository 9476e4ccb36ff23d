"""GPU parity: byte-view path (u16 dictionary + FSST + prefix keys + fingerprints) vs. the CPU oracle.

Known answers transcribed from the reference:
  src/core/README.md:60-104                                  5-row string example (= "apple" under a selection)
  src/core/src/liquid_array/byte_view_array/tests.rs:478-680  ordering compares incl. UTF-8 and nulls
  .../byte_view_array/tests.rs:689-806                        equality incl. len >= 255, NotEq null preservation
  .../byte_view_array/tests.rs:808-851                        needles shorter than the shared prefix, LIKE semantics
  fuzz/fuzz_targets/fsst_view.rs:86-117                       compare_with == arrow cmp on arbitrary strings
"""
import zlib

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle.liquid_oracle import OracleByteViewArray
from tests.util import assert_arrays_equal, assert_masks_equal, random_selection

pytestmark = pytest.mark.gpu


def _bin(op, value):
    from liquid_cache_b200 import BinaryExpr, Column, LiquidExpr, Literal

    return LiquidExpr.new_unchecked(BinaryExpr(Column("liquid_predicate_col", 0), op, Literal(value)))


def _like(pattern, negated=False):
    from liquid_cache_b200 import Column, LikeExpr, LiquidExpr, Literal

    return LiquidExpr.new_unchecked(LikeExpr(negated, False, Column("liquid_predicate_col", 0), Literal(pattern)))


def _hint():
    from liquid_cache_b200 import CacheExpression

    return CacheExpression.SubstringSearch


def test_core_readme_string_example(cache):
    """src/core/README.md: ["apple","banana",NULL,"apple","cherry"], = "apple" with selection [T,T,F,T,T] -> [T,F,T,F]."""
    from liquid_cache_b200 import EntryID

    arr = pa.array(["apple", "banana", None, "apple", "cherry"])
    eid = EntryID(77)
    cache.insert(eid, arr).run()
    assert cache.get(eid).read().equals(arr)
    sel = pa.array([True, True, False, True, True])
    assert cache.get(eid).with_selection(sel).read().to_pylist() == ["apple", "banana", "apple", "cherry"]
    got = cache.eval_predicate(eid, _bin("=", "apple")).with_selection(sel).read()
    assert got.to_pylist() == [True, False, True, False]
    got = cache.eval_predicate(eid, _bin("=", "apple")).read()
    assert got.to_pylist() == [True, False, None, True, False]


WORDS = ["http://", "https://", "www.", "google", "yandex", ".ru", ".com", "/search?q=", "%D0%BA", "%D1%83", "auto",
         "maps", "tours", "/", "-", "_", "&", "=", "id", "page", "1", "2", "3", "77", "2013", ""]


def make_strings(rng, n, n_unique, null_p=0.1, min_tokens=0, max_tokens=12, prefix=""):
    uniq = []
    for _ in range(n_unique):
        k = int(rng.integers(min_tokens, max_tokens + 1))
        uniq.append(prefix + "".join(WORDS[int(i)] for i in rng.integers(0, len(WORDS), size=k)))
    idx = rng.integers(0, n_unique, size=n)
    vals = [uniq[int(i)] for i in idx]
    mask = rng.random(n) < null_p if null_p else None
    return vals, mask


STRING_TYPES = [pa.string(), pa.binary(), pa.string_view(), pa.binary_view(), pa.dictionary(pa.uint16(), pa.string()),
                pa.dictionary(pa.uint16(), pa.binary())]


def build(vals, mask, typ):
    base = typ.value_type if pa.types.is_dictionary(typ) else typ
    is_text = pa.types.is_string(base) or pa.types.is_string_view(base)
    data = [None if (mask is not None and mask[i]) else (v if is_text else v.encode()) for i, v in enumerate(vals)]
    if pa.types.is_dictionary(typ):
        return pa.array(data, type=pa.string() if is_text else pa.binary()).dictionary_encode().cast(typ)
    # built directly in the target type: pyarrow 24's cast() to a view type can leave a null variadic buffer
    # behind, which its own C-data exporter then dereferences (segfault in arrow::ExportArray)
    return pa.array(data, type=typ)


@pytest.mark.parametrize("typ", STRING_TYPES, ids=str)
@pytest.mark.parametrize("n,n_unique", [(1, 1), (100, 7), (3000, 400), (8192, 1900), (10000, 9000)])
def test_round_trip_and_filter(cache, typ, n, n_unique):
    rng = np.random.default_rng(zlib.crc32(repr((str(typ), n, n_unique)).encode()))
    vals, mask = make_strings(rng, n, n_unique, prefix="http://" if n % 2 == 0 else "")
    arr = build(vals, mask, typ)
    liquid = cache.transcode(arr, compressor_scope=zlib.crc32(repr((str(typ), n)).encode()))
    assert liquid.len() == n
    got = liquid.to_arrow_array()
    assert_arrays_equal(got, arr, f"to_arrow {typ}")
    oracle = OracleByteViewArray.from_arrow(arr)
    for p in (0.5, 0.02, 0.0):
        sel = random_selection(rng, n, p)
        assert_arrays_equal(liquid.filter(sel), oracle.filter(sel), f"filter p={p} {typ}")


@pytest.mark.parametrize("with_fp", [False, True])
@pytest.mark.parametrize("n,n_unique,prefix", [(64, 5, ""), (5000, 700, "https://"), (8192, 1900, ""), (9000, 8000, "ab")])
def test_predicates_match_oracle_and_arrow(cache, with_fp, n, n_unique, prefix):
    """compare_with == arrow cmp (the differential spec of fuzz/fuzz_targets/fsst_view.rs)."""
    rng = np.random.default_rng(zlib.crc32(repr((with_fp, n, n_unique, prefix)).encode()))
    vals, mask = make_strings(rng, n, n_unique, prefix=prefix)
    arr = build(vals, mask, pa.string())
    hint = _hint() if with_fp else None
    liquid = cache.transcode(arr, hint=hint, compressor_scope=1000 + n + (1 if with_fp else 0))
    oracle = OracleByteViewArray.from_arrow(arr, build_fingerprints=with_fp)
    present = [v for v in vals[:50]]
    needles = present[:6] + ["", prefix, prefix[:1], prefix + "zzzzzzzzzzzzzzzz", "a", "http", "https://www.google.com/search?q=",
                             vals[0][: max(1, len(vals[0]) // 2)], vals[1] + "x", "éè", "zzzz"]
    for p in (1.0, 0.3):
        sel = random_selection(rng, n, p)
        filt = arr.filter(sel)
        for needle in needles:
            for op, fn in (("=", pc.equal), ("!=", pc.not_equal), ("<", pc.less), ("<=", pc.less_equal),
                           (">", pc.greater), (">=", pc.greater_equal)):
                got = liquid.try_eval_predicate(_bin(op, needle), sel)
                assert_masks_equal(got, oracle.try_eval_predicate(op, needle, sel), f"{op} {needle!r} vs oracle")
                assert_masks_equal(got, fn(filt, pa.scalar(needle)), f"{op} {needle!r} vs arrow")
        for inner in ("google", "tours", "D0", "q", "zzzz", "://", (vals[2][1:6] or "x").replace("%", "5").replace("_", "-"), "2013"):
            pat = f"%{inner}%"
            for negated in (False, True):
                got = liquid.try_eval_predicate(_like(pat, negated), sel)
                want = oracle.try_eval_predicate("not like" if negated else "like", pat, sel)
                assert_masks_equal(got, want, f"like {pat} neg={negated} fp={with_fp} vs oracle")
                if not (negated and with_fp):
                    # with fingerprints NOT LIKE carries the reference's "no candidate -> all false" behaviour
                    arrow = pc.match_substring(filt, inner)
                    assert_masks_equal(got, pc.invert(arrow) if negated else arrow, f"like {pat} vs arrow")


@pytest.mark.parametrize("with_fp", [False, True])
def test_like_needle_length_boundaries(cache, with_fp):
    """The code-domain matcher handles needles up to 31 bytes (state bit 31 is reserved); 32 and beyond take the
    decode + KMP path. Both must agree with arrow's match_substring around the switch."""
    rng = np.random.default_rng(31 + with_fp)
    vals, mask = make_strings(rng, 6000, 1500, min_tokens=6, max_tokens=30)
    arr = build(vals, mask, pa.string())
    liquid = cache.transcode(arr, hint=_hint() if with_fp else None, compressor_scope=4400 + with_fp)
    oracle = OracleByteViewArray.from_arrow(arr, build_fingerprints=with_fp)
    sel = random_selection(rng, len(vals), 0.8)
    filt = arr.filter(sel)
    longest = sorted(set(vals), key=len, reverse=True)[:20]
    for m in (1, 2, 7, 8, 9, 29, 30, 31, 32, 33, 64, 70):
        for src in longest[:4]:
            if len(src) < m + 2:
                continue
            inner = src[1:1 + m].replace("%", "5").replace("_", "-")
            for needle in (inner, inner[:-1] + "~"):
                pat = f"%{needle}%"
                got = liquid.try_eval_predicate(_like(pat), sel)
                assert_masks_equal(got, oracle.try_eval_predicate("like", pat, sel), f"like m={m} vs oracle")
                assert_masks_equal(got, pc.match_substring(filt, needle), f"like m={m} vs arrow")


def test_long_values_and_len255_path(cache):
    """byte_view_array/tests.rs:689-774: values with >= 255 byte suffixes take the len==255 gates."""
    base = "x" * 300
    vals = [base, base + "a", base + "b", "x" * 254, "x" * 255, "x" * 256, None, "short", ""]
    arr = pa.array(vals)
    liquid = cache.transcode(arr, compressor_scope=31337)
    oracle = OracleByteViewArray.from_arrow(arr)
    assert_arrays_equal(liquid.to_arrow_array(), arr, "long round trip")
    sel = pa.array([True] * len(vals))
    for needle in (base, base + "a", "x" * 255, "x" * 254, "x" * 299, "", "short", base + "c"):
        for op, fn in (("=", pc.equal), ("!=", pc.not_equal), ("<", pc.less), (">=", pc.greater_equal), ("<=", pc.less_equal), (">", pc.greater)):
            got = liquid.try_eval_predicate(_bin(op, needle), sel)
            assert_masks_equal(got, fn(arr, pa.scalar(needle)), f"{op} len={len(needle)}")
            assert_masks_equal(got, oracle.try_eval_predicate(op, needle, sel), f"{op} len={len(needle)} oracle")


def test_constant_and_unsupported_predicates(cache):
    from liquid_cache_b200 import LiquidExpr, Literal
    from liquid_cache_b200 import _native as N

    arr = pa.array(["a", None, "b"])
    liquid = cache.transcode(arr, hint=_hint(), compressor_scope=5)
    sel = pa.array([True, True, True])
    assert liquid.try_eval_predicate(LiquidExpr.new_unchecked(Literal(True)), sel).to_pylist() == [True, None, True]
    assert liquid.try_eval_predicate(LiquidExpr.new_unchecked(Literal(False)), sel).to_pylist() == [False, None, False]
    for pat in ("https://%", "%a_b%", "%", "%%", "abc"):
        with pytest.raises(N.UnsupportedExpr):
            liquid.try_eval_predicate(_like(pat), sel)


def test_all_null_and_empty(cache):
    for arr in (pa.array([None, None], pa.string()), pa.array([], pa.string()), pa.array(["", "", None]), pa.array(["same"] * 100)):
        liquid = cache.transcode(arr, compressor_scope=99)
        assert_arrays_equal(liquid.to_arrow_array(), arr, "edge")
        if len(arr):
            sel = pa.array([True] * len(arr))
            assert_masks_equal(liquid.try_eval_predicate(_bin("=", ""), sel), pc.equal(arr, pa.scalar("")), "eq empty")
            assert_masks_equal(liquid.try_eval_predicate(_bin("!=", ""), sel), pc.not_equal(arr, pa.scalar("")), "ne empty")
            assert_masks_equal(liquid.try_eval_predicate(_bin(">", "s"), sel), pc.greater(arr, pa.scalar("s")), "gt")


def test_scan_pipeline_matches_per_call_path(cache):
    """liquid_cache_reader.rs:297-391: conjuncts evaluated in order, nulls -> false, and_then, then projected reads."""
    from liquid_cache_b200 import CacheExpression
    from oracle.liquid_oracle import OracleIntArray, boolean_buffer_and_then, prep_null_mask_filter

    rng = np.random.default_rng(5)
    n_batches, rows = 6, 8192
    ints, strs, li, ls = [], [], [], []
    for b in range(n_batches):
        iv = pa.array(rng.integers(0, 1000, size=rows), pa.int64(), mask=rng.random(rows) < 0.05)
        sv, sm = make_strings(rng, rows, 500, prefix="http://")
        sa = build(sv, sm, pa.string())
        ints.append(iv)
        strs.append(sa)
        li.append(cache.transcode(iv))
        ls.append(cache.transcode(sa, hint=CacheExpression.SubstringSearch, compressor_scope=4242))
    hi = np.array([l.handle for l in li], dtype=np.uint64)
    hs = np.array([l.handle for l in ls], dtype=np.uint64)
    with cache.scan([rows] * n_batches) as scan:
        scan.filter(hi, _bin(">=", 100), pa.int64())
        scan.filter(hi, _bin("<", 600), pa.int64())
        scan.filter(hs, _like("%google%"), pa.string())
        counts, total = scan.counts()
        got_i = scan.read(hi)
        got_s = scan.read(hs)
        want_i, want_s = [], []
        for b in range(n_batches):
            sel = pa.array([True] * rows)
            for m in (pc.greater_equal(ints[b], 100), pc.less(ints[b], 600)):
                sel = boolean_buffer_and_then(sel, prep_null_mask_filter(m.filter(sel)))
            m = pc.match_substring(strs[b], "google")
            sel = boolean_buffer_and_then(sel, prep_null_mask_filter(m.filter(sel)))
            assert int(counts[b]) == sum(sel.to_pylist())
            assert scan.selection(b).to_pylist() == sel.to_pylist()
            want_i.append(ints[b].filter(sel))
            want_s.append(strs[b].filter(sel))
        assert total == sum(len(x) for x in want_i)
        assert_arrays_equal(got_i, pa.concat_arrays(want_i), "scan ints")
        assert_arrays_equal(got_s, pa.concat_arrays(want_s), "scan strings")


@pytest.mark.parametrize("with_nulls", [False, True])
def test_scan_read_device_matches_host_read(cache, with_nulls):
    """lc_scan_read_device: the same concatenated result as lc_scan_read, left in caller-owned device memory
    (values / int32 offsets / validity words concatenated at bit granularity on the device)."""
    import torch

    from liquid_cache_b200.dist import gather_device_result_to_rank0

    rng = np.random.default_rng(77 + with_nulls)
    sizes = [8192, 1000, 8192, 37, 4096, 1]  # ragged batches: row bases that are not multiples of 32
    ints, strs, li, ls = [], [], [], []
    for rows in sizes:
        mask = (rng.random(rows) < 0.2) if with_nulls else None
        iv = pa.array(rng.integers(-1000, 1000, size=rows), pa.int32(), mask=mask)
        sv, sm = make_strings(rng, rows, max(1, rows // 4), null_p=0.2 if with_nulls else 0.0)
        ints.append(iv)
        strs.append(build(sv, sm, pa.string()))
        li.append(cache.transcode(iv))
        ls.append(cache.transcode(strs[-1], compressor_scope=5151 + with_nulls))
    hi = np.array([l.handle for l in li], dtype=np.uint64)
    hs = np.array([l.handle for l in ls], dtype=np.uint64)
    dev = torch.device("cuda", 0)
    with cache.scan(sizes) as scan:
        for pred in (None, _bin(">=", -300)):
            if pred is not None:
                scan.filter(hi, pred, pa.int32())
            want_i, want_s = scan.read(hi), scan.read(hs)
            v, o, b, rows, nulls = scan.read_torch(hi, dev)
            assert o is None and rows == len(want_i) and nulls == want_i.null_count
            got_i = gather_device_result_to_rank0(v, o, b, rows, nulls, pa.int32(), 0, 1)
            assert_arrays_equal(got_i, want_i, "device ints")
            v, o, b, rows, nulls = scan.read_torch(hs, dev)
            assert rows == len(want_s) and nulls == want_s.null_count and int(o[-1].item()) == v.numel()
            got_s = gather_device_result_to_rank0(v, o, b, rows, nulls, pa.string(), 0, 1)
            assert_arrays_equal(got_s, want_s, "device strings")


def test_differential_spec_hypothesis_gpu(cache):
    """fuzz/fuzz_targets/fsst_view.rs:86-117 through the C ABI: arbitrary unicode / control bytes, nulls, selections,
    every comparison operator and LIKE, CUDA path == arrow (and == the oracle, which passes the same search on CPU)."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    text = st.text(alphabet=st.sampled_from(list("abgoleGOLE/%._-:?=&0189 \x00\x7f") + ["é", "к", "\U0001F600"]), max_size=14)
    scope = [7000]

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(vals=st.lists(st.one_of(st.none(), text), min_size=1, max_size=40), needle=text, data=st.data())
    def run(vals, needle, data):
        arr = pa.array(vals, pa.string())
        sel = pa.array(data.draw(st.lists(st.booleans(), min_size=len(vals), max_size=len(vals))))
        filt = arr.filter(sel)
        for fp in (False, True):
            scope[0] += 1
            liquid = cache.transcode(arr, hint=_hint() if fp else None, compressor_scope=scope[0])
            assert_arrays_equal(liquid.to_arrow_array(), arr, "round trip")
            assert_arrays_equal(liquid.filter(sel), filt, "filter")
            for op, fn in (("=", pc.equal), ("!=", pc.not_equal), ("<", pc.less), ("<=", pc.less_equal), (">", pc.greater),
                           (">=", pc.greater_equal)):
                assert_masks_equal(liquid.try_eval_predicate(_bin(op, needle), sel), fn(filt, pa.scalar(needle)), f"{op} {needle!r}")
            inner = needle.replace("%", "").replace("_", "").replace("\\", "")
            if inner:
                got = liquid.try_eval_predicate(_like(f"%{inner}%"), sel)
                assert_masks_equal(got, pc.match_substring(filt, inner), f"like {inner!r} fp={fp}")

    run()


@pytest.mark.parametrize("with_fp", [False, True])
def test_streaming_like_walks_agree_with_arrow(cache, with_fp):
    """k_str_like (full-length outputs: no selection, or a scan's refine) walks the candidates the gate leaves on their FSST
    codes in two ways — a handful of candidates: the warp takes 32 codes of one value at once and composes their Shift-And
    steps (k_str.cu like_candidates_warp); many: 32 values at once, a lane each. Both must give Arrow's match_substring for
    needles planted at every offset around the 32-code block boundary, in values full of bytes the symbol table has to
    escape, for LIKE and NOT LIKE, with the private filter (planes) and without it (no fingerprints: every value walks)."""
    rng = np.random.default_rng(77 + with_fp)
    alphabet = [chr(c) for c in range(0x21, 0x7f)] + ["é", "ß", "й", "中", " "]  # rare code points end up escaped
    needle = "Zq~mark§x"
    vals = []
    for i in range(3000):
        body = "".join(alphabet[int(x)] for x in rng.integers(0, len(alphabet), size=int(rng.integers(5, 120))))
        if i % 17 == 0:
            at = int(rng.integers(0, len(body) + 1)) if i % 34 else (i // 34) % 70  # every offset 0..69 once
            body = body[:at] + needle + body[at:]
        if i % 97 == 0:
            body = body + needle[:-1]  # a near miss at the very end
        vals.append(body)
    vals += [needle, "", needle * 3, "x" * 400 + needle, needle + "y" * 400]
    vals = vals * 2  # dictionary of ~3000 values, each key twice
    arr = pa.array(vals)
    liquid = cache.transcode(arr, hint=_hint() if with_fp else None, compressor_scope=4600 + with_fp)
    for nd in (needle, needle[:3], needle[2:], "q~", "§", "中", "no such thing in here", needle + "y"):
        want = pc.match_substring(arr, nd)
        assert_masks_equal(liquid.try_eval_predicate(_like(f"%{nd}%"), None), want, f"like {nd!r} fp={with_fp}")
        assert_masks_equal(liquid.try_eval_predicate(_like(f"%{nd}%", True), None), pc.invert(want), f"not like {nd!r} fp={with_fp}")
    # the same through a scan (MODE_REFINE), after another conjunct has thinned the selection
    h = np.array([liquid.handle], dtype=np.uint64)
    with cache.scan([len(vals)]) as scan:
        scan.filter(h, _bin(">=", "P"), pa.string())
        scan.filter(h, _like(f"%{needle}%"), pa.string())
        keep = pc.and_(pc.greater_equal(arr, pa.scalar("P")), pc.match_substring(arr, needle))
        assert_arrays_equal(scan.read(h), arr.filter(keep), "refine + read")
