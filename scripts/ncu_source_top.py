#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` dump (per kernel, with that kernel's own header)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
k, hdr, data = None, None, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        k = r[1]; data[k] = {"hdr": None, "rows": []}; continue
    if r and r[0] == "Address":
        data[k]["hdr"] = r; continue
    if k and r:
        data[k]["rows"].append(r)
for k, d in data.items():
    hdr, v = d["hdr"], d["rows"]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si]) for r in v)
    agg = {c: sum(int(r[hdr.index(c)]) for r in v) for c in cols}
    s = sum(agg.values()) or 1
    print("=====", k[:70], "samples", tot, "SASS instrs", len(v), "warp instr executed", sum(int(r[ie]) for r in v))
    print("   stall mix:", {c: round(100 * t / s, 1) for c, t in sorted(agg.items(), key=lambda kv: -kv[1])[:7]})
    top = sorted(enumerate(v), key=lambda t: -int(t[1][si]))[:top_n]
    for idx, r in sorted(top):
        st = {c: int(r[hdr.index(c)]) for c in cols if int(r[hdr.index(c)]) > 0}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"{idx:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}%  x{int(r[ie]):9d}  {r[src].strip()[:64]:64s} {st}")
