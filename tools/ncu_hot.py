#!/usr/bin/env python
"""Development aid: from `ncu -i rep --page source --csv`, print the most-executed straight run of SASS (the hot loop) with
per-instruction executed counts, stall samples and the dominant stall reason.

    ncu -i k2.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [min_exec_fraction]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
data = rows[2:]
ex = [int(r[ix["Instructions Executed"]] or 0) for r in data]
mx = max(ex)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_samples = sum(int(r[ix["# Samples"]] or 0) for r in data)
hot_samples = 0
print(f"max executed {mx}, total samples {tot_samples}")
for r, e in zip(data, ex):
    if e < frac * mx:
        continue
    s = int(r[ix["# Samples"]] or 0)
    hot_samples += s
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{e/mx:5.2f} {s:6d}  {r[ix['Source']].strip():70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
print(f"samples in listed instructions: {hot_samples} of {tot_samples} ({100*hot_samples/tot_samples:.1f}%)")
