"""Pins the CPU oracle on the reference's SQL-level snapshots: the answers the REFERENCE produced on
examples/nano_hits.parquet, reproduced by running the oracle's insert / eval_predicate / get under the reader loop of
tests/nano_hits.py. Fixtures: tests/golden/ (generator: make_fixtures.py). No GPU needed."""
import numpy as np
import pyarrow as pa
import pytest

from oracle.liquid_oracle import OracleFsst, transcode
from tests import nano_hits as NH


@pytest.fixture(scope="module")
def world():
    batches, answers = NH.load()
    store = {}
    fsst = {}
    for rg, bi, cols in batches:
        for c, arr in cols.items():
            if pa.types.is_string(arr.type):
                # one FSST table per (row group, column), trained on the chunk's first batch (transcode.rs:16-33)
                if (rg, c) not in fsst:
                    fsst[(rg, c)] = OracleFsst.train([v.encode() for v in arr.to_pylist() if v is not None])
                for hinted in (False, True):  # with and without the SubstringSearch hint (fingerprints)
                    store[(rg, bi, c, hinted)] = transcode(arr, fsst[(rg, c)], build_fingerprints=hinted)
            else:
                store[(rg, bi, c, False)] = transcode(arr)
    return batches, answers, store


def test_fixture_shape():
    batches, answers = NH.load()
    assert sum(len(c["URL"]) for _, _, c in batches) == 24586  # 24 576 + 10 rows
    assert [(rg, bi) for rg, bi, _ in batches] == [(0, 0), (0, 1), (0, 2), (1, 0)]
    assert set(answers) == set(NH.QUERIES)


@pytest.mark.parametrize("name", list(NH.QUERIES))
def test_reference_snapshot_answers(world, name):
    batches, answers, store = world

    hinted = NH.QUERIES[name][3]

    def eval_predicate(key, column, op, lit, sel):
        o = store[(key[0], key[1], column, column in hinted)]
        return o.try_eval_predicate(op, lit, pa.array(sel))

    def get(key, column, sel):
        return store[(key[0], key[1], column, column in hinted)].filter(pa.array(sel))

    got = NH.run_query(name, batches, eval_predicate, get)
    assert NH.rows_match(name, got, answers[name]["rows"]), f"{answers[name]['sql']}\n got {got[:3]}\nwant {answers[name]['rows'][:3]}"
