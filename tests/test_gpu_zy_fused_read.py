"""lc_scan_read with device-side bookkeeping (scan_host.cc scan_read_fused, k_scan_plan.cu): after the first read of a scan
has taught it its sizes, a get over the device-resident selection costs one synchronisation — row / byte prefix sums run on
the device against capacities taken from the previous read, the result comes down with a speculative prefix. Whatever the
path (first read planned on the host, device-planned, a read that outgrows the speculative download, a read that outgrows the
capacities and falls back), the array must be Arrow's filter of the column under the reader's selection
(liquid_cache_reader.rs:342-391)."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from liquid_cache_b200 import BinaryExpr, CacheExpression, Column, LikeExpr, LiquidExpr, Literal
from tests.util import assert_arrays_equal

pytestmark = pytest.mark.gpu


def _bin(op, v):
    return LiquidExpr.new_unchecked(BinaryExpr(Column("c", 0), op, Literal(v)))


def _like(p):
    return LiquidExpr.new_unchecked(LikeExpr(False, False, Column("c", 0), Literal(p)))


def _columns(cache, rng, n_batches, rows, scope):
    ints, strs, hi, hs = [], [], [], []
    hosts = ["google.com", "yandex.ru", "example.org", "mail.ru", "bing.com"]
    for b in range(n_batches):
        n = rows if b % 3 else rows - 37  # ragged batches
        iv = pa.array(rng.integers(0, 1000, size=n), pa.int64())
        sv = pa.array([f"http://{hosts[int(rng.integers(0, 5))]}/p/{int(rng.integers(0, 400))}?q={int(rng.integers(0, 50))}" for _ in range(n)])
        ints.append(iv)
        strs.append(sv)
        hi.append(cache.transcode(iv))
        hs.append(cache.transcode(sv, hint=CacheExpression.SubstringSearch, compressor_scope=scope))
    return ints, strs, hi, hs


def test_repeated_reads_take_the_device_planned_path_and_agree_with_arrow(cache):
    rng = np.random.default_rng(11)
    n_batches, rows = 24, 4096
    ints, strs, li, ls = _columns(cache, rng, n_batches, rows, 8801)
    hi = np.array([l.handle for l in li], dtype=np.uint64)
    hs = np.array([l.handle for l in ls], dtype=np.uint64)
    sizes = [len(a) for a in ints]
    # thresholds chosen so that consecutive reads shrink, grow a little (speculative download too small) and grow a lot
    # (capacities too small: the call falls back to the host-planned path and re-learns)
    plan = [(">=", 500), (">=", 520), (">=", 480), (">=", 10), (">=", 900), (">=", 880)]
    with cache.scan(sizes) as scan:
        for op, thr in plan:
            scan.reset()
            scan.filter(hi, _bin(op, thr), pa.int64())
            got_i = scan.read(hi)
            got_s = scan.read(hs)
            want_i = pa.concat_arrays([a.filter(pc.greater_equal(a, thr)) for a in ints])
            want_s = pa.concat_arrays([s.filter(pc.greater_equal(a, thr)) for a, s in zip(ints, strs)])
            assert_arrays_equal(got_i, want_i, f"ints {op} {thr}")
            assert_arrays_equal(got_s, want_s, f"strings {op} {thr}")


def test_selective_like_then_get_of_the_hits(cache):
    rng = np.random.default_rng(12)
    n_batches, rows = 40, 2048
    ints, strs, li, ls = _columns(cache, rng, n_batches, rows, 8802)
    hs = np.array([l.handle for l in ls], dtype=np.uint64)
    hi = np.array([l.handle for l in li], dtype=np.uint64)
    sizes = [len(a) for a in ints]
    with cache.scan(sizes) as scan:
        for pattern in ("%bing%", "%bing.com/p/7%", "%yandex%", "%nothing-matches-this%", "%bing%"):
            inner = pattern.strip("%")
            scan.reset()
            scan.filter(hs, _like(pattern), pa.string())
            got_s = scan.read(hs)
            got_i = scan.read(hi)
            masks = [pc.match_substring(s, inner) for s in strs]
            assert_arrays_equal(got_s, pa.concat_arrays([s.filter(m) for s, m in zip(strs, masks)]), pattern)
            assert_arrays_equal(got_i, pa.concat_arrays([a.filter(m) for a, m in zip(ints, masks)]), pattern + " ints")
            counts, total = scan.counts()
            assert total == len(got_s) and [int(c) for c in counts] == [int(pc.sum(m).as_py() or 0) for m in masks]
