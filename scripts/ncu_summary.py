#!/usr/bin/env python
"""Summarise ncu outputs into small text files that can be committed under profiles/.

  python scripts/ncu_summary.py launches <launches.csv>          -> per-kernel count / total / share table
  python scripts/ncu_summary.py full <prof.ncu-rep or raw.csv>   -> key metrics per profiled launch
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "")[:90]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ni, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(r[ui], 1e-6)
        tot[short(r[ni])] += v
        cnt[short(r[ni])] += 1
    total = sum(tot.values())
    print(f"# {len(rows) - 1} launches, {total:.3f} ms of kernel time (ncu: cold-cache, serialised -- compare shares, not absolutes)")
    print(f"{'kernel':92s} {'count':>6s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k:92s} {cnt[k]:6d} {tot[k]:10.3f} {tot[k] / cnt[k]:9.4f} {tot[k] / total:7.3f}")


def full(path):
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    else:
        text = open(path).read()
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"## {short(r[hdr.index('Kernel Name')])}   grid {r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''} block {r[hdr.index('Block Size')] if 'Block Size' in hdr else ''}")
        for m in KEY_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:86s} {r[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
