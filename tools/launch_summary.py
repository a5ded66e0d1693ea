#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = []
for ln in csv.reader(open(sys.argv[1], errors="ignore")):
    if len(ln) >= 15 and ln[0].isdigit():
        rows.append((ln[4].split("(")[0].replace("void ", ""), ln[8], float(ln[14])))
d = collections.defaultdict(list)
for k, g, t in rows:
    d[k].append(t)
tot = sum(t for _, _, t in rows)
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:34s} n={len(v):5d} mean={sum(v)/len(v)/1e3:9.1f} us  min={min(v)/1e3:8.1f} max={max(v)/1e3:8.1f}  share={100*sum(v)/tot:5.1f}%")
