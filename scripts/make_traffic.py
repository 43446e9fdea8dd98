#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the bench's TIMED regime (the synthetic run >= 20 cycles in, as the
driver's warmup 5 / steps 20 run sees it): per-launch DRAM bytes (read + write), duration and DRAM throughput of the kernels
behind bench.py's particle entry points.  bench.py reports `roofline.traffic` from this file, scaled by the rank's share of
the particles, only for the configuration the capture was made on (config 3, 512^3).

usage: python scripts/make_traffic.py gpurun_out/<tag>/particles_timed_regime.ncu-rep profiles/traffic.json <capture label>"""
import csv
import io
import json
import subprocess
import sys

ENTRY = {"k_deposit<2": "projection_T00_Tij_project", "k_geodesic<2": "kick_drift", "k_scatter": "rebin_sort", "k_prepare_tensor": "prepareFTsource_tensor"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}
text = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(text)))
hdr, units = rows[0], rows[1]
out = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    for pat, entry in ENTRY.items():
        if pat in name:
            def val(m):
                i = hdr.index(m)
                return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
            rd, wr, ms = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")
            out[entry] = {"kernel": name.split("(")[0].replace("void ", "").replace("<unnamed>::", ""), "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                          "ncu_ms_per_launch": ms, "ncu_dram_gbs": (rd + wr) / (ms * 1e-3) / 1e9,
                          "regime": "timed run of bench.py: synthetic ICs evolved by 21 cycles (ncu -s 60 of the three particle kernels, warmup 21)",
                          "capture": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
