#!/usr/bin/env python
"""Condense ncu output into the small CSV/markdown summaries kept under profiles/.

    python tools/ncu_summary.py launches <launches.csv>            per-kernel mean duration and share of the step
    python tools/ncu_summary.py full <report.ncu-rep> [...]        one row per captured kernel, key metrics
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__sectors_read.sum", "dram__sectors_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_atom_dot_cas_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_atom_dot_cas_lookup_miss.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("kernel,launches,mean_us,share_pct")
    for k, v in agg.items():
        print(f"\"{k}\",{len(v)},{sum(v) / len(v) / 1e3:.1f},{sum(v) / tot * 100:.1f}")


def full(paths):
    w = csv.writer(sys.stdout)
    first = True
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr = rows[0]
        idx = [hdr.index(k) for k in KEEP if k in hdr]
        if first:
            w.writerow([hdr[i] for i in idx])
            w.writerow([rows[1][i] for i in idx])
            first = False
        for r in rows[2:]:
            w.writerow([r[i][:70] if hdr[i] == "Kernel Name" else r[i] for i in idx])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
