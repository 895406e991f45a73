#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo and
--import-source on).  Usage: python tools/ncu_lines.py rep.ncu-rep kernel_regex [min_pct]"""
import csv
import subprocess
import sys

rep, rex = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + rex], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
f = None
agg = []
tot_i = tot_s = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0] != "":
        try:
            ie, s = int(r[7] or 0), int(r[6] or 0)
        except ValueError:
            continue
        agg.append((f, int(r[0]), r[1], ie, s))
        tot_i += ie
        tot_s += s
print(f"# warp instructions {tot_i}, stall samples {tot_s}")
for f, ln, src, ie, s in agg:
    if s > tot_s * min_pct / 100 or ie > tot_i * min_pct / 100:
        print(f"{f}:{ln} inst {100 * ie / max(tot_i, 1):4.1f}% samp {100 * s / max(tot_s, 1):4.1f}%  {src.strip()[:100]}")
