"""The reference's `Arc<LiquidCache>` is hit by every DataFusion partition task at once — every method takes `&self`, the
index is a lock-free ART, the budget a CAS loop (cache/core.rs:52-63, index.rs:12-60, budget.rs:37-53). Here every calling
thread gets its own lane (CUDA stream, scratch, staging buffers) and only the short operations on shared state take the
context's lock: eight host threads issue mixed insert / eval_predicate / get / scan calls against one cache and every answer
must equal the serial one (Arrow on the arrays that were inserted)."""
import threading

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from liquid_cache_b200 import (BinaryExpr, CacheExpression, Column, EntryID, LikeExpr, LiquidCacheBuilder, LiquidExpr, Literal,
                               parquet_array_id)
from tests.util import assert_arrays_equal, assert_masks_equal

pytestmark = pytest.mark.gpu

N_THREADS = 8
ROUNDS = 12


def _bin(op, v):
    return LiquidExpr.new_unchecked(BinaryExpr(Column("c", 0), op, Literal(v)))


def _like(p):
    return LiquidExpr.new_unchecked(LikeExpr(False, False, Column("c", 0), Literal(p)))


def _strings(rng, n):
    hosts = ["google.com", "yandex.ru", "example.org", "mail.ru", "bing.com"]
    return pa.array([None if rng.random() < 0.03 else f"http://{hosts[int(rng.integers(0, 5))]}/p/{int(rng.integers(0, 300))}"
                     for _ in range(n)])


def test_eight_threads_mixed_calls_match_the_serial_answers():
    cache = LiquidCacheBuilder.new().build()
    errors = []
    try:
        # shared, read-only entries every thread queries: one string column chunk (32 batches, ONE symbol table) + ints
        rng = np.random.default_rng(0)
        shared_s = [_strings(rng, 4096) for _ in range(16)]
        shared_i = [pa.array(rng.integers(-500, 500, size=4096), pa.int32(), mask=rng.random(4096) < 0.02) for _ in range(16)]
        ids_s = [parquet_array_id(7, 0, 3, b) for b in range(16)]
        ids_i = [parquet_array_id(7, 0, 4, b) for b in range(16)]
        cache.insert_many(ids_s, shared_s, hint=CacheExpression.SubstringSearch)
        cache.insert_many(ids_i, shared_i)
        hs, hi = cache.handles([int(x) for x in ids_s]), cache.handles([int(x) for x in ids_i])
        start = threading.Barrier(N_THREADS)

        def worker(t):
            try:
                rng = np.random.default_rng(100 + t)
                start.wait()
                for r in range(ROUNDS):
                    # 1. insert this thread's own batches: strings of ONE column chunk shared by all threads (row group 1,
                    #    column 5: whoever comes first trains the table, the others wait for it), then integers
                    mine_s = _strings(rng, 2048)
                    mine_i = pa.array(rng.integers(0, 1 << 40, size=3000, dtype=np.int64), pa.int64())
                    sid = parquet_array_id(7, 1, 5, t * ROUNDS + r)
                    iid = parquet_array_id(7, 1 + t, 6, r)
                    cache.insert(sid, mine_s).with_squeeze_hint(CacheExpression.SubstringSearch).run()
                    cache.insert(iid, mine_i).run()
                    # 2. read them back, filtered
                    sel = pa.array(rng.random(2048) < 0.5)
                    assert_arrays_equal(cache.get(sid).with_selection(sel).read(), mine_s.filter(sel), f"t{t} get strings")
                    thr = int(rng.integers(0, 1 << 40))
                    assert_masks_equal(cache.eval_predicate(iid, _bin(">=", thr)).read(), pc.greater_equal(mine_i, pa.scalar(thr, pa.int64())),
                                       f"t{t} int predicate")
                    # 3. predicates on the shared entries
                    b = int(rng.integers(0, 16))
                    got = cache.eval_predicate(ids_s[b], _like("%google%")).read()
                    assert_masks_equal(got, pc.match_substring(shared_s[b], "google"), f"t{t} shared like")
                    k = int(rng.integers(-500, 500))
                    got = cache.eval_predicate(ids_i[b], _bin("<", k)).read()
                    assert_masks_equal(got, pc.less(shared_i[b], pa.scalar(k, pa.int32())), f"t{t} shared int")
                    # 4. a scan over all shared batches on this thread's own lane: two conjuncts, then both columns read
                    with cache.scan([4096] * 16) as scan:
                        scan.filter(hi, _bin(">=", k), pa.int32())
                        scan.filter(hs, _like("%yandex%"), pa.string())
                        got_s = scan.read(hs)
                        got_i = scan.read(hi)
                    masks = [pc.and_(pc.fill_null(pc.greater_equal(a, pa.scalar(k, pa.int32())), False),
                                     pc.fill_null(pc.match_substring(s, "yandex"), False)) for a, s in zip(shared_i, shared_s)]
                    assert_arrays_equal(got_s, pa.concat_arrays([s.filter(m) for s, m in zip(shared_s, masks)]), f"t{t} scan strings")
                    assert_arrays_equal(got_i, pa.concat_arrays([a.filter(m) for a, m in zip(shared_i, masks)]), f"t{t} scan ints")
                    # 5. replace and remove this thread's older entries while the others keep going
                    if r >= 2:
                        cache.remove(parquet_array_id(7, 1 + t, 6, r - 2))
            except Exception as e:  # noqa: BLE001 - reported by the main thread
                errors.append((t, repr(e)))

        threads = [threading.Thread(target=worker, args=(t,)) for t in range(N_THREADS)]
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=600)
        assert not errors, errors[:3]
        assert all(not th.is_alive() for th in threads)
        # the cache is intact afterwards: shared entries still answer, every thread's last entries are there
        assert_arrays_equal(cache.get(ids_s[3]).read(), shared_s[3], "shared after the storm")
        for t in range(N_THREADS):
            assert cache.is_cached(parquet_array_id(7, 1 + t, 6, ROUNDS - 1))
            assert not cache.is_cached(parquet_array_id(7, 1 + t, 6, 0))
    finally:
        cache.close()
