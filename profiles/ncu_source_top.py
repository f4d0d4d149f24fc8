#!/usr/bin/env python
"""Top stall sites from `ncu --page source --csv` (SASS view).  usage: ncu_source_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] in ("Kernel Name", "Address"):
        if r[0] == "Kernel Name": break      # next view / kernel
        continue
    if len(r) == len(hdr): data.append(r)
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, " instructions executed", sum(int(r[ix["Instructions Executed"]] or 0) for r in data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    top = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*s/tot:5.1f}%  exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:70]:70s} {top}")
