#!/usr/bin/env python
"""Builds synth/hits_sample.parquet FROM THE REFERENCE CHECKOUT: the 25 columns of examples/nano_hits.parquet (24 586 real
ClickBench rows shipped with the reference) that the 43 ClickBench queries touch (benchmark/clickbench/queries/q0-q42.sql).
bench.py's clickbench_sweep workload draws its rows from this sample (synth/hits.py). Runs only where /root/reference
exists; the output is committed so the GPU box has it."""
import os

import pyarrow.parquet as pq

COLUMNS = ["AdvEngineID", "ResolutionWidth", "UserID", "SearchPhrase", "EventDate", "RegionID", "MobilePhoneModel", "MobilePhone",
           "SearchEngineID", "EventTime", "URL", "Title", "CounterID", "Referer", "IsRefresh", "ClientIP", "WatchID",
           "DontCountHits", "IsLink", "IsDownload", "TraficSourceID", "URLHash", "RefererHash", "WindowClientWidth",
           "WindowClientHeight"]

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    t = pq.read_table("/root/reference/examples/nano_hits.parquet", columns=COLUMNS)
    pq.write_table(t, os.path.join(here, "hits_sample.parquet"), compression="zstd", compression_level=19)
    print(t.num_rows, "rows,", os.path.getsize(os.path.join(here, "hits_sample.parquet")), "bytes")
