#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total and share.
    python tools/launch_list.py gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = collections.OrderedDict()
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    key = (name[:90], r["Grid Size"], r["Block Size"])
    c = tot.setdefault(key, [0, 0.0])
    c[0] += 1
    c[1] += float(r["Metric Value"].replace(",", "")) / 1e3
total = sum(v[1] for v in tot.values())
print(f"{'kernel':92s} {'grid':>16s} {'block':>14s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
for (name, grid, block), (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:92s} {grid:>16s} {block:>14s} {n:4d} {us:10.1f} {us / n:9.1f} {100 * us / total:6.1f}%")
print(f"total {total:.1f} us over {sum(v[0] for v in tot.values())} launches (cold-cache, serialised: compare shares)")
