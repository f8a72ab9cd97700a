#!/usr/bin/env python
"""Per-source-line summary of an ncu report captured with --import-source on (kernels built with -lineinfo):
    python tools/ncu_source_lines.py report.ncu-rep [top_n]
Prints, for the hottest CUDA source lines, the share of executed warp instructions and of stall samples."""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, items = None, None, []
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif "Instructions Executed" in r:
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":  # a CUDA source line (its SASS rows follow with addresses)
        try:
            items.append((float(r[hdr.index("Instructions Executed")]), float(r[hdr.index("# Samples")]), cur_file, r[0], r[1].strip()))
        except ValueError:
            pass
ti, ts = sum(i[0] for i in items), sum(i[1] for i in items)
print(f"total warp instructions {ti:.0f}, stall samples {ts:.0f}")
for v, smp, f, ln, src in sorted(items, reverse=True)[:top]:
    print(f"{100 * v / ti:5.1f}% instr {100 * smp / max(ts, 1):5.1f}% samples  {f}:{ln}  {src[:110]}")
