#!/usr/bin/env python
"""Summarise ncu outputs into text for profiles/: launch list shares and key raw metrics per kernel.

  python tools/ncu_summary.py launches <launches.csv>
  python tools/ncu_summary.py raw <report.ncu-rep> [regex]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEYS = re.compile(
    r"^(gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|lts__throughput.avg.pct_of_peak_sustained_elapsed|"
    r"l1tex__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|"
    r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.pct_of_peak_sustained_active|"
    r"launch__registers_per_thread|launch__grid_size|launch__block_size|launch__occupancy_limit_\w+|"
    r"sm__cycles_elapsed.avg|smsp__inst_executed.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|"
    r"l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum|l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active.ratio)$"
)


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[hi + 1 :]:
        if len(r) > vi:
            try:
                agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':70s} {'n':>4s} {'avg_us':>10s} {'total_ms':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} {len(v):4d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/1e6:9.3f} {sum(v)/tot:6.3f}")


def raw(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pat and not re.search(pat, r[ki]):
            continue
        print("==", r[ki][:100])
        for i, h in enumerate(hdr):
            if KEYS.match(h):
                print(f"  {h:86s} {units[i]:16s} {r[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
