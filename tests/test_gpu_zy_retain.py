"""`LiquidCache::try_read_liquid` hands out an `Arc<dyn LiquidArray>` (cache/core.rs:243-252): the array a caller holds
keeps answering for the entry AS IT WAS, whatever happens to the id in the cache meanwhile (lc_cache_retain)."""
import pyarrow as pa
import pytest

from liquid_cache_b200 import BinaryExpr, Column, EntryID, LiquidCacheBuilder, LiquidExpr, Literal

pytestmark = pytest.mark.gpu


def _gt(v):
    return LiquidExpr.new_unchecked(BinaryExpr(Column("c", 0), ">", Literal(v)))


def test_a_held_array_survives_reinsert_remove_and_reset():
    cache = LiquidCacheBuilder.new().build()
    try:
        first = pa.array(list(range(100, 1100)), pa.int64())
        cache.insert(EntryID(7), first).run()
        held = cache.try_read_liquid(EntryID(7))
        assert held is not None and held.len() == 1000
        # the id is replaced by a different array: the held one still reads as the first
        second = pa.array(list(range(5)), pa.int64())
        cache.insert(EntryID(7), second).run()
        assert cache.get(EntryID(7)).read().to_pylist() == [0, 1, 2, 3, 4]
        assert held.to_arrow_array().equals(first)
        assert held.try_eval_predicate(_gt(1000), None).to_pylist() == [v > 1000 for v in first.to_pylist()]
        # churn the arena so a freed blob would be overwritten
        for i in range(64):
            cache.insert(EntryID(100 + i), pa.array([i] * 4096, pa.int64())).run()
        assert held.to_arrow_array().equals(first)
        cache.reset()
        assert cache.try_read_liquid(EntryID(7)) is None
        assert held.len() == 1000 and held.to_arrow_array().equals(first)
        del held
    finally:
        cache.close()


def test_absent_entries_read_as_none():
    cache = LiquidCacheBuilder.new().build()
    try:
        assert cache.try_read_liquid(EntryID(1)) is None
        assert cache.eval_predicate(EntryID(1), _gt(0)).read() is None
    finally:
        cache.close()


def test_churn_under_a_budget_recycles_hbm():
    """with_max_memory_bytes bounds what the arena RESERVES: replacing and removing entries over and over reuses the freed
    ranges instead of growing the reservation, and an insert that cannot fit reports CacheFull (cache/budget.rs:37-53)."""
    import numpy as np

    from liquid_cache_b200 import _native as N

    budget = 8 << 20
    cache = LiquidCacheBuilder.new().with_max_memory_bytes(budget).build()
    try:
        rng = np.random.default_rng(3)
        for round_ in range(40):
            for i in range(24):
                # 8192 x u64 at W = 64: ~64 KB per entry; 24 live entries ~1.6 MB, far below the budget
                arr = pa.array(rng.integers(0, 1 << 63, size=8192, dtype=np.int64), pa.int64())
                cache.insert(EntryID(i), arr).run()
            st = cache.stats()
            assert st.hbm_bytes_used <= budget
            assert st.hbm_bytes_reserved <= budget, (round_, st.hbm_bytes_reserved)
        last = pa.array(rng.integers(0, 1 << 63, size=8192, dtype=np.int64), pa.int64())
        cache.insert(EntryID(3), last).run()
        assert cache.get(EntryID(3)).read().equals(last)
        # fill up until the cache says it is full; the reservation never passes the budget
        with pytest.raises(N.CacheFull):
            for i in range(100, 400):
                cache.insert(EntryID(i), pa.array(rng.integers(0, 1 << 63, size=8192, dtype=np.int64), pa.int64())).run()
        assert cache.stats().hbm_bytes_reserved <= budget
        assert cache.get(EntryID(3)).read().equals(last)
    finally:
        cache.close()
