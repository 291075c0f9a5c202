#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
instructions executed, stall samples.  usage: ncu_lines.py dump.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = []
hdr = None
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
    if r[0].isdigit():  # a source line summary row
        try:
            inst = int(r[7]) if r[7] not in ("-", "") else 0
            samp = int(r[6]) if r[6] not in ("-", "") else 0
        except ValueError:
            continue
        agg.append((inst, samp, cur_file, int(r[0]), r[1].strip()))
tot_i = sum(a[0] for a in agg)
tot_s = sum(a[1] for a in agg)
print(f"total instructions {tot_i:.4g}, samples {tot_s}")
for inst, samp, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{inst/tot_i:6.1%} inst {samp/max(tot_s,1):6.1%} smp  {f}:{ln:<4d} {src[:110]}")
