#!/usr/bin/env python
"""Hottest CUDA source lines of one kernel of an .ncu-rep (stall samples aggregated per source line).

    python tools/ncu_lines.py report.ncu-rep KERNEL_REGEX [top_n]
"""
import csv, io, subprocess, sys
path, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
lines = []
for r in rows[h + 1:]:
    if len(r) > ii and r[0].strip().isdigit():
        num = lambda v: int(v) if v.strip().isdigit() else 0
        lines.append((num(r[si]), num(r[ii]), int(r[0]), r[1].strip()))
ts = sum(l[0] for l in lines) or 1
ti = sum(l[1] for l in lines) or 1
print(f"{ts} samples, {ti} warp instructions")
for s, n, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * s / ts:5.1f}% samples {100 * n / ti:5.1f}% instr  L{ln:<5d} {src[:150]}")
