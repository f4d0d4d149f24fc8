#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: launch_summary.py launches.csv [top]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for x in csv.DictReader(lines):
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", x["Kernel Name"]))[:64]
    agg[name][0] += 1; agg[name][1] += ms; tot += ms
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{ms:10.3f} ms {100 * ms / tot:5.1f} % {n:5d} x  {k}")
print(f"{tot:10.3f} ms total (serialised, cold-cache launch times)")
