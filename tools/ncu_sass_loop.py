#!/usr/bin/env python
"""Per-SASS-instruction view of a kernel's hottest loop from `ncu -i rep --page source --csv --print-source sass`:
executions, active threads, stall samples and the top stall reasons of every instruction between the first instruction that
matches <anchor> minus <before> and plus <after> lines.  usage: ncu_sass_loop.py dump.csv [anchor=DADD.RM] [before=16] [after=26]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
anchor = sys.argv[2] if len(sys.argv) > 2 else "DADD.RM"
before = int(sys.argv[3]) if len(sys.argv) > 3 else 16
after = int(sys.argv[4]) if len(sys.argv) > 4 else 26
print(rows[0][1][:120])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in body)
toti = sum(int(r[ix["Instructions Executed"]]) for r in body)
at = next(i for i, r in enumerate(body) if anchor in r[1])
lo, hi = max(0, at - before), min(len(body), at + after)
print(f"kernel: {len(body)} SASS instructions, {toti / 1e6:.1f} M warp instructions executed, {tot} stall samples")
S = I = 0
for r in body[lo:hi]:
    s, e = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
    S, I = S + s, I + e
    top = sorted(((int(r[ix[k]]), k) for k in reasons), reverse=True)[:3]
    print(f"{r[1].strip()[:56]:56s} exec {e / 1e6:7.2f}M thr {r[ix['Avg. Threads Executed']]:>5s}  smp {s:5d} {100 * s / tot:4.1f}%  "
          + " ".join(f"{k[6:]}:{v}" for v, k in top if v))
print(f"listed region: {100 * I / toti:.1f} % of the executed instructions, {100 * S / tot:.1f} % of the samples")
agg = sorted(((sum(int(r[ix[k]]) for r in body), k[6:]) for k in reasons), reverse=True)[:8]
print("stall samples by reason (whole kernel):", ", ".join(f"{k} {v}" for v, k in agg))
