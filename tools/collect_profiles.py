#!/usr/bin/env python
"""copies the summaries of a tools/run_round2.sh run from gpurun_out/r2_<tag>_* into profiles/r2_* and rebuilds
profiles/traffic.json (DRAM bytes + counters of the k_cost_batch_tile captures, stamped with the kernel-source hash).
usage: python tools/collect_profiles.py <tag>"""
import glob
import hashlib
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out", f"r2_{tag}_")
HASHED = ["dpose_b200/csrc/get_cost_batch_tile.cu", "dpose_b200/csrc/cost_math.cuh"]


def metrics(name):
    d = {}
    for ln in open(G + name + "_metrics.txt"):
        m = re.match(r"\s+(\S+)\s+([-\d.e+]+)\s*(\S*)", ln)
        if m:
            d.setdefault(m.group(1), (float(m.group(2)), m.group(3)))
    return d


def to_bytes(v):
    return v[0] * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(v[1], 1)


h = hashlib.sha256()
for f in HASHED:
    h.update(open(os.path.join(ROOT, f), "rb").read())
out = {"kernel_source_hash": h.hexdigest()[:16], "hashed_files": HASHED,
       "_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE k_cost_batch_tile launch (ncu --set full, "
                "profiles/r2_tile_<workload>_metrics.txt); bench.py copies the figure into roofline.traffic and flags it STALE "
                "when the hashed files have changed since the capture.  The map bytes are read by k_pack_tiles (100 MB + "
                "12.6 MB written per step at cfg5), the windows are then served from the L2-resident tile plane."}
for w in ("cfg5", "cfg2", "cfg3"):
    m = metrics("tile_" + w)
    t = m["gpu__time_duration.sum"]
    out[w] = int(to_bytes(m["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"]))
    out["ncu_" + w] = {"kernel_ms": t[0] if t[1] == "ms" else t[0] / 1e3,
                       "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0],
                       "warps_active_pct": m["sm__warps_active.avg.pct_of_peak_sustained_active"][0],
                       "thread_inst_per_inst": m["smsp__thread_inst_executed_per_inst_executed.ratio"][0],
                       "warp_instructions": m["smsp__inst_executed.sum"][0],
                       "fp64_pipe_pct": m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][0],
                       "alu_pipe_pct": m["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"][0],
                       "xu_pipe_pct": m["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"][0],
                       "limiter": "instruction issue + the two half-rate pipes (FP64 in the cell walk, ALU in phase 1)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
n = 0
for f in glob.glob(G + "*"):
    if f.endswith(".ncu-rep") or f.endswith("bench.err"):
        continue
    shutil.copy(f, os.path.join(ROOT, "profiles", os.path.basename(f).replace(f"r2_{tag}_", "r2_")))
    n += 1
print("copied", n, "files; traffic:", {k: v for k, v in out.items() if k in ("cfg5", "cfg2", "cfg3", "kernel_source_hash")})
