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


def test_one_pass_read_of_a_selective_scan(cache):
    """k_str_read_onepass (k_str.cu): when few rows survive per batch the whole read — sizes, the prefix sums across entries
    (a chained scan over the kernel's own CTAs), decode — is one kernel. Survivors are skewed on purpose: most batches keep
    none or a couple, one keeps hundreds (more than the 64 a warp holds in shared memory, so the lane-owned second pass
    runs), some surviving values are empty and some are longer than 254 bytes (their length is not in the PrefixKey)."""
    import torch

    from liquid_cache_b200.dist import DeviceGather

    rng = np.random.default_rng(14)
    n_batches, rows = 128, 1024
    ints, strs, li, ls = [], [], [], []  # li / ls keep the transcoded arrays (and with them their handles) alive
    for b in range(n_batches):
        n = rows if b % 5 else rows - 13
        iv = rng.integers(0, 1000, size=n)
        if b == 77:
            iv[rng.random(n) < 0.6] = 999  # one dense batch
        if b in (3, 4, 5, 90):
            iv[:] = 0                      # batches without survivors, some of them neighbours (a CTA with nothing to do)
        vals = []
        for i in range(n):
            r = int(rng.integers(0, 40))
            vals.append("" if r == 0 else ("x" * 300 + str(i % 7)) if r == 1 else f"http://h{int(rng.integers(0, 9))}.example/p/{int(rng.integers(0, 300))}")
        ints.append(pa.array(iv, pa.int64()))
        strs.append(pa.array(vals))
        li.append(cache.transcode(ints[-1]))
        ls.append(cache.transcode(strs[-1], compressor_scope=8804))
    hi, hs = np.array([l.handle for l in li], dtype=np.uint64), np.array([l.handle for l in ls], dtype=np.uint64)
    sizes = [len(a) for a in ints]
    thr = 998
    want_s = pa.concat_arrays([s.filter(pc.greater_equal(a, thr)) for a, s in zip(ints, strs)])
    assert 64 < int(pc.sum(pc.greater_equal(ints[77], thr)).as_py()) and len(want_s) <= 8 * n_batches
    dev = torch.device("cuda", 0)
    with cache.scan(sizes) as scan:
        for _ in range(3):  # the first read is planned on the host and teaches the sizes; the next ones take the one-pass kernel
            scan.reset()
            scan.filter(hi, _bin(">=", thr), pa.int64())
            assert_arrays_equal(scan.read(hs), want_s, "one-pass read, host result")
        # the asynchronous form: row capacity within 16 per batch selects the one-pass kernel; a short byte capacity is reported
        g = DeviceGather(pa.string(), 0, 1, dev, rows_cap=16 * n_batches, values_cap=2048)
        for attempt in range(6):
            assert scan.read_async(hs, *g.addresses())
            cache.synchronize()
            (n_rows, n_bytes, overflow), = g.exchange()
            assert n_rows == len(want_s)
            if not overflow:
                break
            assert overflow == 2 and n_bytes == sum(len(v) for v in want_s.to_pylist())  # both totals are reported
            g.grow()
        assert attempt > 0 and not g.overflowed()
        assert_arrays_equal(g.to_arrow(), want_s, "one-pass read, device result")
        # nothing survives: an empty array, closing offset 0
        scan.reset()
        scan.filter(hi, _bin(">=", 5000), pa.int64())
        assert scan.read_async(hs, *g.addresses())
        cache.synchronize()
        assert g.exchange() == [(0, 0, 0)]
        assert len(g.to_arrow()) == 0
