#!/usr/bin/env python
"""Group the SASS of one kernel of an .ncu-rep into regions of equal execution count (basic-block-ish)."""
import csv, io, subprocess, sys
path, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.008
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[h]; si = hdr.index('Source'); ii = hdr.index('Instructions Executed'); sm = hdr.index('# Samples')
body = [(int(r[ii]), int(r[sm]), r[si].strip()) for r in rows[h + 1:] if len(r) > ii and r[ii].isdigit()]
tot = sum(b[0] for b in body); ts = sum(b[1] for b in body) or 1
groups = []
for n, s, src in body:
    if groups and abs(groups[-1][0] - n) <= 0.02 * max(n, 1):
        g = groups[-1]; g[1] += 1; g[2] += s; g[3] += n; g[4].append(src)
    else:
        groups.append([n, 1, s, n, [src]])
print(f"total {tot} warp instructions, {ts} samples")
for n, c, s, sumn, srcs in groups:
    if sumn > tot * thr or s > ts * thr:
        ops = {}
        for x in srcs:
            op = (x.split()[1] if x.startswith('@') else x.split()[0]).split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
        print(f'~{n:>8} x {c:>4} instrs = {100*sumn/tot:5.1f}% instr  {100*s/ts:5.1f}% samples  {top}')
