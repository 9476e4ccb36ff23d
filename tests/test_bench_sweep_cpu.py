"""The ClickBench sweep harness (bench_sweep.py) over a pyarrow test double: every query of the table lowers, runs, and its
survivor counts equal Arrow's; the literals taken from the sample select something; the IN-list conjunct takes the host
fallback. No GPU needed (the CUDA path runs the same harness through bench.py --workload clickbench_sweep)."""
import time

import pyarrow as pa
import pytest

import bench_sweep as S
from tests.fake_cache import FakeCache


def _timer():
    t = [0.0]

    def start():
        t[0] = time.perf_counter()

    return start, (lambda: (time.perf_counter() - t[0]) * 1e3)


def test_query_table_is_complete_and_well_formed():
    assert [q for q, _c, _p in S.QUERIES] == list(range(43))
    from synth.hits import HitsSample

    sample = HitsSample()
    for c in S.columns_used():
        assert c in sample.cols, c
    for _q, conj, proj in S.QUERIES:
        for column, op, _lit in conj:
            assert op in ("=", "!=", "like", "not like", ">=", "<=", "in"), op
            if op in ("like", "not like"):
                assert pa.types.is_string(sample.cols[column].type)


@pytest.mark.parametrize("bulk", [False, True])  # True: the IN conjunct edits the selection words in bulk, as on a GPU
def test_sweep_runs_and_matches_arrow_on_the_test_double(bulk):
    res = S.run_sweep(FakeCache(bulk_selections=bulk), rows=8192 * 24, steps=1, warmup=1, timer=_timer, check_batches=24)
    assert res["all_counts_match_arrow"] and len(res["queries"]) == 43
    by_q = {r["q"]: r for r in res["queries"]}
    assert by_q[0]["note"] == "no column touched"
    assert by_q[19]["rows_out"] > 0, "the UserID literal comes from the sample and must select rows"
    assert by_q[36]["rows_out"] > 0 and by_q[42]["rows_out"] > 0, "CounterID from the sample + July 2013 dates select rows"
    assert by_q[38]["rows_out"] == 0  # IsLink is constantly 0 in the sample
    assert by_q[40]["rows_out"] <= by_q[39]["rows_out"] and "IN" in by_q[40]["note"]
    assert by_q[12]["selectivity"] < 0.5  # most SearchPhrase values are empty
    assert not any("not run" in r.get("note", "") for r in res["queries"])
