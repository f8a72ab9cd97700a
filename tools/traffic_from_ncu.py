#!/usr/bin/env python
"""profiles/secondary_traffic.json from per-op ncu launch lists (one CSV per op of tools/profile_ops.py, captured with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum): DRAM bytes (read + write) of ALL kernels of one
invocation of the op, which bench.py attaches to the matching secondary record as `traffic`.
    python tools/traffic_from_ncu.py gpurun_out/r2_ops_ profiles/secondary_traffic.json
"""
import csv
import json
import os
import sys

prefix, out = sys.argv[1], sys.argv[2]
# op of tools/profile_ops.py -> (bench.py secondary record, invocations per profile_ops run, kernels to leave out)
OPS = {
    "lookup_conv": ("lookup_convc1_fused", 2, ("pack_weights", "absmax", "scales", "convert_kmajor", "pool2x2", "pack_record", "pool_pack", "corr_pyramid_gemm")),
    "cfg2": ("cfg2_flownetc_corr_fwd", 2, ()),
    "cfg3": ("cfg3_pwc_corr_fwd", 2, ()),
    "cfg4": ("cfg4_dcvnet_matryoshka_fwd", 2, ()),
}
res = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, tools/profile_ops.py <op>, sum over the "
                 "kernels of one invocation (cold caches: small ops whose output stays in L2 under-report their writes)"}
for op, (rec, ninv, skip) in OPS.items():
    path = f"{prefix}{op}.csv"
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    tot = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[mi].startswith("dram__bytes") or "at::" in r[ki] or any(s in r[ki] for s in skip):
            continue
        tot += float(r[vi].replace(",", ""))
    res[rec] = tot / ninv
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
