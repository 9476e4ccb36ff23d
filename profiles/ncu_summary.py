#!/usr/bin/env python
"""Summarise an .ncu-rep (details page) into a text file: python profiles/ncu_summary.py in.ncu-rep out.txt"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
ni, vi, ui, si, ki, ii = (hdr.index(x) for x in ("Metric Name", "Metric Value", "Metric Unit", "Section Name", "Kernel Name", "ID"))
with open(out, "w") as f:
    last = None
    for r in rows[1:]:
        if r[ii] != last:
            f.write(f"\n==== launch {r[ii]}: {r[ki]}\n")
            last = r[ii]
        f.write(f"{r[si]:38s} {r[ni]:52s} {r[vi]} {r[ui]}\n")
