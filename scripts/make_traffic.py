#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture: per-launch DRAM bytes (read + write) of the kernels behind
bench.py's entry points.  usage: python scripts/make_traffic.py gpurun_out/<tag>/prof.ncu-rep profiles/traffic.json"""
import csv
import io
import json
import subprocess
import sys

ENTRY = {"k_deposit<2": "projection_T00_Tij_project", "k_geodesic<2": "kick_drift", "k_scatter": "rebin_sort", "k_prepare_tensor": "prepareFTsource_tensor"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
text = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(text)))
hdr, units = rows[0], rows[1]
out = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    for pat, entry in ENTRY.items():
        if pat in name:
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(m)
                tot += float(r[i]) * UNIT[units[i]]
            out[entry] = tot
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
