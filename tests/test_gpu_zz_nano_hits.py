"""GPU parity on the reference's SQL-level snapshots: the six datafusion-local queries over examples/nano_hits.parquet
(src/datafusion-local/src/tests/mod.rs:187-441), answered through the CUDA path — insert / eval_predicate /
get().with_selection() of the C ABI under the reader loop of tests/nano_hits.py — and compared with the tables the
REFERENCE produced (tests/golden/nano_hits_answers.json). The CPU oracle passes the same driver in
tests/test_oracle_nano_hits.py. (Named to sort after the other GPU test files.)"""
import pyarrow as pa
import pytest

from tests import nano_hits as NH

pytestmark = pytest.mark.gpu

COLUMN_IDS = {"WatchID": 0, "OS": 42, "EventTime": 4, "URL": 13, "Referer": 14}  # positions in `hits`


@pytest.mark.parametrize("qi,name", list(enumerate(NH.QUERIES)))
def test_reference_snapshot_answers_on_the_gpu(cache, qi, name):
    from liquid_cache_b200 import (BinaryExpr, CacheExpression, Column, LikeExpr, LiquidExpr, Literal, _native as N,
                                   parquet_array_id)

    batches, answers = NH.load()
    conjuncts, projection, _finish, hinted = NH.QUERIES[name]
    used = {c for c, _op, _lit in conjuncts} | set(projection)
    types = {}
    for rg, bi, cols in batches:  # what LiquidCacheReader inserts while the query first runs
        for c in used:
            eid = parquet_array_id(20 + qi, rg, COLUMN_IDS[c], bi)
            ins = cache.insert(eid, cols[c])
            if c in hinted:
                ins = ins.with_squeeze_hint(CacheExpression.SubstringSearch)
            ins.run()
            types[c] = cols[c].type

    def eid_of(key, column):
        return parquet_array_id(20 + qi, key[0], COLUMN_IDS[column], key[1])

    def eval_predicate(key, column, op, lit, sel):
        col = Column(column, 0)
        expr = LikeExpr(False, False, col, Literal(lit)) if op == "like" else BinaryExpr(col, op, Literal(lit))
        hint = CacheExpression.SubstringSearch if column in hinted else None
        lexpr = LiquidExpr.try_new(expr, types[column], hint)
        if lexpr is None:
            return None  # LiquidExpr::try_new -> None: the reader decodes and evaluates with Arrow (column.rs:143-151)
        try:
            return cache.eval_predicate(eid_of(key, column), lexpr).with_selection(pa.array(sel)).read()
        except N.UnsupportedExpr:
            return None

    def get(key, column, sel):
        return cache.get(eid_of(key, column)).with_selection(pa.array(sel)).read()

    got = NH.run_query(name, batches, eval_predicate, get)
    assert NH.rows_match(name, got, answers[name]["rows"]), f"{answers[name]['sql']}\n got {got[:3]}\nwant {answers[name]['rows'][:3]}"
