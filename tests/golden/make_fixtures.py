#!/usr/bin/env python
"""Regenerates the golden fixtures of this directory FROM THE REFERENCE CHECKOUT (/root/reference, read-only). It only
runs where that checkout exists (the build container); the GPU box gets the committed outputs.

Inputs (nothing is copied verbatim but the data the reference's own tests run on):
  examples/nano_hits.parquet                                   the 24 586-row ClickBench sample the reference's
                                                               datafusion-local tests query (row groups 24 576 + 10)
  src/datafusion-local/src/tests/mod.rs:187-441                the SQL of those tests
  src/datafusion-local/src/tests/snapshots/*.snap              the answers the REFERENCE ITSELF produced for them (insta)

Outputs:
  nano_hits_subset.parquet   the five columns those queries touch (WatchID, OS, EventTime, URL, Referer), same row groups
  nano_hits_answers.json     per test: the SQL and the result table transcribed from the snapshot
"""
import json
import os
import re

import pyarrow.parquet as pq

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
SNAP = os.path.join(REF, "src/datafusion-local/src/tests/snapshots")
TESTS = {  # snapshot name -> SQL (src/datafusion-local/src/tests/mod.rs)
    "url_prefix_filtering": 'select COUNT(*) from hits where "URL" like \'https://%\'',
    "url_selection_and_ordering": 'select "URL" from hits where "URL" like \'%tours%\' order by "URL" desc',
    "os_selection": 'select "OS" from hits where "URL" like \'%tours%\' order by "OS" desc',
    "referer_filtering": 'select "Referer" from hits where "Referer" <> \'\' AND "URL" like \'%tours%\' order by "Referer" desc',
    "single_column_filter_projection": 'select "WatchID" from hits where "WatchID" = 6978470580070504163',
    "provide_schema_with_filter": 'select "WatchID", "OS", "EventTime" from hits where "OS" <> 2 order by "WatchID" desc limit 10',
}


def parse_values_table(text: str):
    """The `values:` block of an insta snapshot is an Arrow pretty-printed table."""
    block = text.split("\nvalues:", 1)[1].split("\nstats:", 1)[0]  # the header line quotes the format string: anchor on line starts
    lines = [l for l in block.splitlines() if l.startswith("|")]
    cells = [[c.strip() for c in l.strip().strip("|").split("|")] for l in lines]
    return cells[0], cells[1:]


def main():
    src = pq.ParquetFile(os.path.join(REF, "examples/nano_hits.parquet"))
    table = src.read(columns=["WatchID", "OS", "EventTime", "URL", "Referer"])
    pq.write_table(table, os.path.join(HERE, "nano_hits_subset.parquet"), compression="zstd", compression_level=19,
                   row_group_size=src.metadata.row_group(0).num_rows)
    answers = {}
    for name, sql in TESTS.items():
        text = open(os.path.join(SNAP, f"liquid_cache_datafusion_local__tests__{name}.snap"), encoding="utf-8").read()
        cols, rows = parse_values_table(text)
        # confirm the SQL against the test source so a drifting reference is noticed
        mod = open(os.path.join(REF, "src/datafusion-local/src/tests/mod.rs"), encoding="utf-8").read()
        assert sql in mod, f"{name}: SQL not found in tests/mod.rs"
        answers[name] = {"sql": sql, "columns": cols, "rows": rows}
    with open(os.path.join(HERE, "nano_hits_answers.json"), "w", encoding="utf-8") as f:
        json.dump(answers, f, ensure_ascii=False, indent=1)
    print({k: len(v["rows"]) for k, v in answers.items()})


if __name__ == "__main__":
    main()
