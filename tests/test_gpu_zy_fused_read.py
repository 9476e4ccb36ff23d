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


def test_async_read_into_caller_buffers_matches_arrow_and_reports_short_capacities(cache):
    """lc_scan_read_async (scan_host.cc scan_read_async): the same get, with no host synchronisation — result and the
    64-byte header land in caller-owned device memory. DeviceGather is the caller bench.py uses; with world = 1 its
    exchange() is only the header download. Short capacities are reported in the header (rows stay valid), nothing else is
    written, and growing from the headers converges."""
    import torch

    from liquid_cache_b200.dist import DeviceGather

    rng = np.random.default_rng(13)
    n_batches, rows = 30, 4096
    ints, strs, li, ls = _columns(cache, rng, n_batches, rows, 8803)
    hi = np.array([l.handle for l in li], dtype=np.uint64)
    hs = np.array([l.handle for l in ls], dtype=np.uint64)
    sizes = [len(a) for a in ints]
    dev = torch.device("cuda", 0)
    gi = DeviceGather(pa.int64(), 0, 1, dev, rows_cap=256)
    gs = DeviceGather(pa.string(), 0, 1, dev, rows_cap=256, values_cap=1024)
    with cache.scan(sizes) as scan:
        for thr in (990, 500, 0, 2000, 700):
            scan.reset()
            scan.filter(hi, _bin(">=", thr), pa.int64())
            want_i = pa.concat_arrays([a.filter(pc.greater_equal(a, thr)) for a in ints])
            want_s = pa.concat_arrays([s.filter(pc.greater_equal(a, thr)) for a, s in zip(ints, strs)])
            for g, h, want in ((gi, hi, want_i), (gs, hs, want_s)):
                for _ in range(8):
                    assert scan.read_async(h, *g.addresses())
                    cache.synchronize()  # the session cache runs on its own stream, not torch's (bench.py shares one stream)
                    (n_rows, _nbytes, overflow), = g.exchange()
                    assert n_rows == len(want)  # rows are reported even when nothing could be written
                    if not overflow:
                        break
                    g.grow()
                assert not g.overflowed()
                assert_arrays_equal(g.to_arrow(), want, f"async >= {thr}")
    assert gi.grows > 0 and gs.grows > 0


def test_async_read_refuses_what_the_device_plan_does_not_cover(cache):
    import torch

    from liquid_cache_b200.dist import DeviceGather

    vals = pa.array([1, None, 3, 4] * 512, pa.int64())
    l = cache.transcode(vals)
    h = np.array([l.handle], dtype=np.uint64)
    g = DeviceGather(pa.int64(), 0, 1, torch.device("cuda", 0), rows_cap=4096)
    with cache.scan([len(vals)]) as scan:
        scan.filter(h, _bin(">=", 2), pa.int64())
        assert scan.read_async(h, *g.addresses()) is False  # nulls: the caller takes lc_scan_read / lc_scan_read_device
        assert_arrays_equal(scan.read(h), vals.filter(pc.fill_null(pc.greater_equal(vals, 2), False)), "fallback")
