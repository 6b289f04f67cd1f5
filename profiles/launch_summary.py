#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python profiles/launch_summary.py launches.csv"""
import csv, collections, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0][:58]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':60s} {'n':>4s} {'avg us':>9s} {'min':>8s} {'max':>9s} {'share':>6s}")
for k, v in agg.items():
    print(f"{k:60s} {len(v):4d} {sum(v)/len(v):9.2f} {min(v):8.2f} {max(v):9.2f} {100*sum(v)/tot:5.1f}%")
