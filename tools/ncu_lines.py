#!/usr/bin/env python
"""Per-source-line warp-instruction counts / lane utilisation / stall samples from an .ncu-rep captured with
--import-source on (binary built with -lineinfo).  usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = next(r for r in rows if len(r) > 5 and r[0] == "Line No")
ci, ti, si = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
agg, fname = {}, None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if len(r) < 10 or r[0] in ("Line No", ""):
        continue
    try:
        agg[(fname, int(r[0]), r[1].strip()[:110])] = [int(r[ci]), int(r[ti]), int(r[si])]
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values())
tsamp = sum(v[2] for v in agg.values())
print(f"total warp instructions {tot}, samples {tsamp}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]:>10} {100 * v[0] / tot:5.1f}%  thr/inst {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / max(tsamp, 1):5.1f}%  {k[0]}:{k[1]}: {k[2]}")
