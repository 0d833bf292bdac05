#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report captured with --import-source on.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-name regex> [top N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern, "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, agg = None, {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 10 or r[0] in ("", "Line No"):
            continue
        try:
            ln, samples, inst, tinst = int(r[0]), int(r[6]), int(r[7]), int(r[8])
        except ValueError:
            continue
        a = agg.setdefault((cur, ln), [0, 0, 0, r[1]])
        a[0] += inst
        a[1] += tinst
        a[2] += samples
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print(f"warp instructions {tot}, stall samples {tots}")
    for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{f[:14]:14s}{ln:>5} {100 * a[0] / tot:5.1f}% inst  {a[1] / max(a[0], 1):5.1f} lanes  {100 * a[2] / tots:5.1f}% samples  {a[3].strip()[:100]}")


if __name__ == "__main__":
    main()
