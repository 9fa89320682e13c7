#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum of one kernel of an .ncu-rep -> JSON (bench.py's roofline.traffic).

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep binned_raster profiles/raster_traffic.json
"""
import csv, io, json, subprocess, sys
path, pat, out_path = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = []
for r in rows[2:]:
    if pat in r[hdr.index("Kernel Name")]:
        tot = 0.0
        parts = {}
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            parts[key] = float(r[i]) * scale[units[i]]
            tot += parts[key]
        res.append({"kernel": r[hdr.index("Kernel Name")], "traffic_bytes": tot, **parts,
                    "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")])})
json.dump({"source": path, "launches": res, "traffic_bytes": sum(x["traffic_bytes"] for x in res) / max(len(res), 1)}, open(out_path, "w"), indent=1)
print(open(out_path).read())
