#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line:
   python tools/ncu_lines.py dump.csv [top]  -> file:line, warp-instructions executed, stall samples, source text"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "" and r[0].isdigit():
        ie = hdr.index("Instructions Executed")
        ss = hdr.index("# Samples")
        try:
            agg.append((cur_file, int(r[0]), int(r[ie]), int(r[ss]), r[1].strip()))
        except ValueError:
            pass
tot_i = sum(a[2] for a in agg)
tot_s = sum(a[3] for a in agg)
print(f"total warp-instructions {tot_i}, stall samples {tot_s}")
for a in sorted(agg, key=lambda x: -x[2])[:top]:
    print(f"{a[0]}:{a[1]:<5d} inst {a[2]:>12d} {100*a[2]/tot_i:5.1f}%  samples {a[3]:>7d} {100*a[3]/max(tot_s,1):5.1f}%  {a[4][:90]}")
