"""The reference's own SQL-level tests on examples/nano_hits.parquet (src/datafusion-local/src/tests/mod.rs:187-441),
restated at the level of THIS path: what LiquidCacheReader does per 8192-row batch
(src/datafusion/src/reader/runtime/liquid_cache_reader.rs:297-391) — evaluate every conjunct on the cached column under
the running selection (CachedColumn::eval_predicate_with_filter, src/datafusion/src/cache/column.rs:114-152: pushdown when
LiquidExpr::try_new admits the predicate and the array can evaluate it, else get-with-selection + Arrow on the CPU), fold
it in with boolean_buffer_and_then, then get-with-selection of the projected columns — followed by the query's ORDER BY /
LIMIT / COUNT done with pyarrow (those operators are DataFusion's, above the cache).

The expected tables are the reference's insta snapshots (tests/golden/nano_hits_answers.json, transcribed by
tests/golden/make_fixtures.py); the data is tests/golden/nano_hits_subset.parquet (the five columns the queries touch).
A backend supplies eval_predicate / get for a (column, batch); the CPU oracle and the CUDA path both run under this driver.
"""
import json
import os

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq

HERE = os.path.dirname(os.path.abspath(__file__))
BATCH = 8192  # LiquidCacheBuilder default batch size (builders.rs:54, datafusion-local/src/lib.rs:77)

# name -> (conjuncts [(column, op, literal)] in the order the reader evaluates them (`=`/`<>` first, then LIKE:
#          row_filter.rs:501-515), projection, finishing step, columns cached under the SubstringSearch hint).
# The hint — and with it the fingerprints — is given to a string column whose only uses in the query are
# `LIKE '%x%'` patterns or plain projection (lineage_opt.rs:655-686, 939-950); 'https://%' is not such a pattern, so that
# query meets a URL column without fingerprints and LIKE takes the Arrow fallback (helpers.rs:83-88).
QUERIES = {
    "url_prefix_filtering": ([("URL", "like", "https://%")], [], "count", set()),
    "url_selection_and_ordering": ([("URL", "like", "%tours%")], ["URL"], ("order_desc", "URL", None), {"URL"}),
    "os_selection": ([("URL", "like", "%tours%")], ["OS"], ("order_desc", "OS", None), {"URL"}),
    "referer_filtering": ([("Referer", "!=", ""), ("URL", "like", "%tours%")], ["Referer"], ("order_desc", "Referer", None), {"URL"}),
    "single_column_filter_projection": ([("WatchID", "=", 6978470580070504163)], ["WatchID"], None, set()),
    "provide_schema_with_filter": ([("OS", "!=", 2)], ["WatchID", "OS", "EventTime"], ("order_desc", "WatchID", 10), set()),
}


def load():
    f = pq.ParquetFile(os.path.join(HERE, "golden", "nano_hits_subset.parquet"))
    answers = json.load(open(os.path.join(HERE, "golden", "nano_hits_answers.json"), encoding="utf-8"))
    batches = []  # [(row group, batch index, {column: Array})], batches never span row groups
    for rg in range(f.metadata.num_row_groups):
        t = f.read_row_group(rg)
        for b0 in range(0, t.num_rows, BATCH):
            sl = t.slice(b0, min(BATCH, t.num_rows - b0))
            batches.append((rg, b0 // BATCH, {c: sl[c].combine_chunks() for c in t.column_names}))
    return batches, answers


def and_then(left: np.ndarray, right: np.ndarray) -> np.ndarray:
    """boolean_buffer_and_then (src/datafusion/src/utils.rs:62-83) on bool arrays."""
    out = np.zeros(len(left), dtype=bool)
    out[np.flatnonzero(left)] = right
    return out


def run_query(name, batches, eval_predicate, get):
    """eval_predicate(key, column, op, literal, selection bool array) -> BooleanArray over the selected rows, or None when
    the backend cannot push the predicate down; get(key, column, selection) -> Arrow array of the selected rows."""
    conjuncts, projection, finish, _hinted = QUERIES[name]
    out = {c: [] for c in projection}
    total = 0
    for rg, bi, cols in batches:
        key = (rg, bi)
        n = len(next(iter(cols.values())))
        sel = np.ones(n, dtype=bool)
        for column, op, lit in conjuncts:
            if not sel.any():
                break
            mask = eval_predicate(key, column, op, lit, sel)
            if mask is None:  # fallback of column.rs:143-151: decode the selected rows, evaluate with Arrow
                vals = get(key, column, sel)
                if op == "like":
                    mask = pc.match_like(vals, lit)
                else:
                    fn = {"=": pc.equal, "!=": pc.not_equal}[op]
                    mask = fn(vals, pa.scalar(lit, vals.type))
            m = np.asarray(mask.fill_null(False).to_numpy(zero_copy_only=False), dtype=bool)  # prep_null_mask_filter
            assert len(m) == int(sel.sum()), "a predicate mask covers exactly the selected rows"
            sel = and_then(sel, m)
        total += int(sel.sum())
        if sel.any():  # read_from_cache returns early on an empty selection (liquid_cache_reader.rs:346-349)
            for c in projection:
                out[c].append(get(key, c, sel))
    if finish == "count":
        return [[str(total)]]
    table = pa.table({c: pa.concat_arrays(out[c]) if out[c] else pa.array([], batches[0][2][c].type) for c in projection})
    if finish is not None:
        _, col, limit = finish
        table = table.sort_by([(col, "descending")])
        if limit:
            table = table.slice(0, limit)
    return [[str(v) for v in row.values()] for row in table.to_pylist()]


def rows_match(name, got, want) -> bool:
    """ORDER BY leaves ties in any order and the snapshot pads cells: compare trimmed rows; ties compare as multisets
    through the sort key, which all these queries project."""
    want = [[c.strip() for c in r] for r in want]
    return got == want or sorted(map(tuple, got)) == sorted(map(tuple, want)) and [r[0] for r in got] == [r[0] for r in want]
