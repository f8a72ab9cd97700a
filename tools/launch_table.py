#!/usr/bin/env python
"""Per-kernel averages from an `ncu --metrics a,b,c --csv` launch list."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: defaultdict(list))
for r in rows[hi + 1:]:
    if len(r) > vi:
        try:
            agg[r[ki][:60]][r[mi] + " [" + r[ui] + "]"].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(d.get("gpu__time_duration.sum [ns]", [0])) for d in agg.values())
for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum [ns]", [0]))):
    t = d.get("gpu__time_duration.sum [ns]", [0])
    line = f"{k:60s} n={len(t):3d} avg_us={sum(t)/max(len(t),1)/1e3:9.1f} share={sum(t)/max(tot,1):.3f}"
    for m, v in d.items():
        if "dram" in m:
            line += f" {m.split('.')[0].replace('dram__bytes_','dram_')}={sum(v)/len(v)/1e6:9.1f}MB"
    print(line)
