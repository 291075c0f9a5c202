#!/usr/bin/env python
"""writes DESIGN.md from tools/DESIGN.template.md, filling its R2_* placeholders from the bench lines kept under profiles/ (so that the document quotes the
committed evidence and nothing else).  usage: python tools/fill_design.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def line(name):
    for ln in open(os.path.join(P, name)):
        if ln.startswith("{"):
            d = json.loads(ln)
    return d


def lines(name):
    return [json.loads(ln) for ln in open(os.path.join(P, name)) if ln.startswith("{")]


def sci(v):
    m, e = f"{v:.2e}".split("e")
    return f"{m}e{int(e)}"


rep = {}
for w in ("cfg5", "cfg3", "cfg2"):
    d = line(f"r2_bench_{w}.json")
    W = w.upper()
    rep[f"R2_{W}_VALUE"] = sci(d["value"])
    rep[f"R2_{W}_FRAC"] = f"{100 * d['roofline']['frac']:.1f} %"
    rep[f"R2_{W}_MS"] = f"{d['ms_per_step']:.4f}" if d["ms_per_step"] < 0.1 else f"{d['ms_per_step']:.3f}"
    e = d["e2e"]
    rep[f"R2_{W}_E2E"] = f"{sci(e['value'])} ({e['ms_per_step']:.2f} ms" + (
        f"; both copies alone, at once: {e['pcie_bidirectional_ms_per_step']:.2f} ms)" if w == "cfg5" else ")")
    cb = d.get("cpu_baseline") or {}
    rep[f"R2_{W}_CPU"] = f"{sci(cb.get('value', 0))}" + (f" (C port {sci(cb['port_value'])})" if cb.get("port_value") else "")
for w in ("cfg2", "cfg3"):
    to = line(f"r2_bench_{w}.json").get("tile_order") or {}
    rep[f"R2_{w.upper()}_TILE_VALUE"] = sci(to["value"]) if to else "?"
d4 = line("r2_bench_cfg4.json")
r4 = line("r2_bench_cfg4_reference_arm.json")
rep["R2_CFG4_EVALS"] = str(d4["config"]["pose_evals_per_solve"])
rep["R2_CFG4_DEV"] = f"{1e3 * d4['ms_per_step']:.0f}"
rep["R2_CFG4_CALL"] = f"{1e3 * d4['e2e']['ms_per_step']:.0f}"
rep["R2_CFG4_VALUE"] = sci(d4["value"])
rep["R2_CFG4_CPU"] = f"{r4['ms_per_step']:.1f}"
sol = lines("r2_solver.jsonl")
s64 = [x for x in sol if x.get("what") == "solve_batch" and x.get("starts") == 65536][0]
rep["R2_CFG4_64K"] = f"{s64['call_us'] / 1e3:.2f}"
coop = [x for x in sol if x.get("what", "").startswith("batch launch") and not x["volatile_map"]]
rep["R2_COOP_TABLE"] = ", ".join(f"{x['poses']} poses {x['us_per_call']:.1f} µs" for x in coop)
side = lines("r2_side_kernels.jsonl")
let = [x for x in side if x.get("what") == "lethal_cells_within" and x.get("box", "").startswith("full")]
clean = [x for x in let if x["plane"].startswith("clean")][0]
dirty = [x for x in let if not x["plane"].startswith("clean")][0]
rep["R2_LETHAL_CLEAN_FRAC"] = f"{100 * clean['frac_of_hbm_peak']:.0f} %"
rep["R2_LETHAL_DIRTY_FRAC"] = f"{100 * dirty['frac_of_hbm_peak']:.0f} %"
rep["R2_LETHAL_CLEAN"] = f"{clean['device_us']:.0f}"
rep["R2_LETHAL_DIRTY"] = f"{dirty['device_us']:.0f}"
pre = [x for x in side if x.get("what") == "precompute"]
rep["R2_PRECOMPUTE_TABLE"] = "; ".join(
    f"{x['footprint'].split('_')[0]} {x['table'][0]}×{x['table'][1]}: {x['gpu_kernel_us']:.0f} µs device / {x['gpu_call_wall_us']:.0f} µs per call (cv2, one thread: {x['cpu_cv2_us']:.0f} µs)"
    for x in pre)
d1 = line("r2_bench_cfg5.json")
lo1 = d1.get("list_order") or {}
rep["R2_CFG5_LIST_MS"] = f"{lo1['ms_per_step']:.3f}" if lo1 else "?"
rep["R2_CFG5_LIST_VALUE"] = sci(lo1["value"]) if lo1 else "?"
rows = ["| GPUs | strong: poses/s (step) | efficiency | the job in list order, plain call (efficiency) | clean map | weak (list order): poses/s | e2e per-rank calls | e2e `_multi` (one process) | cost-only / sweep / sweep cost-only e2e |",
        "|---|---|---|---|---|---|---|---|---|"]
v1 = d1["value"]
v1c = (d1.get("clean_map") or {}).get("value")
for n, d in ((1, d1),) + tuple((n, line(f"r2_bench_cfg5_n{n}.json")) for n in (2, 4, 8) if os.path.exists(os.path.join(P, f"r2_bench_cfg5_n{n}.json"))):
    strong = d if d["scaling"] == "strong" else None
    other = d.get("other_scaling") or {}
    weak_v = other.get("value") if n > 1 else (lo1.get("value") or d["value"])  # the weak-scaling leg draws its poses per rank: list order
    cm = d.get("clean_map") or {}
    ev = d.get("e2e_variants") or {}
    em = d.get("e2e_multi") or {}
    lo = d.get("list_order") or {}
    rows.append(f"| {n} | {sci(d['value'])} ({d['ms_per_step']:.3f} ms) | {d['value'] / (n * v1):.2f} | "
                + (f"{sci(lo['value'])} ({lo['value'] / (n * lo1['value']):.2f})" if lo and lo1 else "—") + " | "
                + (f"{sci(cm['value'])} ({cm['value'] / (n * v1c):.2f})" if cm and v1c else "—") + " | "
                + (f"{sci(weak_v)} ({weak_v / (n * (lo1.get('value') or v1)):.3f})" if weak_v else "—") + " | "
                + f"{sci(d['e2e']['value'])} | " + (sci(em["value"]) if em.get("value") else "—") + " | "
                + " / ".join(sci(ev[k]["value"]) for k in ("cost_only", "sweep", "sweep_cost_only") if k in ev) + " |")
rep["R2_SCALING_TABLE"] = "\n\n" + "\n".join(rows) + "\n"
if os.path.exists(os.path.join(P, "r2_bench_cfg5_n8.json")):
    d8 = line("r2_bench_cfg5_n8.json")
    rep["R2_PACK_SHARE_N8"] = f"{100 * 0.052 / d8['ms_per_step']:.0f} %"
    rep["R2_N8_VALUE"] = sci(d8["value"])
    rep["R2_N8_EFF"] = f"{d8['value'] / (8 * v1):.2f}"
    rep["R2_N8_LIST_EFF"] = f"{(d8.get('list_order') or {}).get('value', 0) / (8 * lo1['value']):.2f}" if lo1 else "?"
t = json.load(open(os.path.join(P, "traffic.json")))["ncu_cfg5"]
rep["R2_TILE_COUNTERS"] = (f"ncu of this round's build: {t['kernel_ms']:.3f} ms per launch, issue slots {t['issue_active_pct']:.1f} % busy, "
                           f"{t['warps_active_pct']:.1f} % of the warp slots occupied, "
                           f"`smsp__thread_inst_executed_per_inst_executed` = {t['thread_inst_per_inst']:.2f} of 32 "
                           f"({100 * t['thread_inst_per_inst'] / 32:.0f} % of the lanes do work in the average instruction: the "
                           f"divergence of the cell walk), FP64 pipe {t['fp64_pipe_pct']:.0f} %, ALU pipe {t.get('alu_pipe_pct', 0):.0f} %, "
                           f"XU (find-leading-one, conversions) {t['xu_pipe_pct']:.0f} %, "
                           f"{t['warp_instructions'] / 1e7:.1f} warp instructions per pose")
path = os.path.join(ROOT, "DESIGN.md")
s = open(os.path.join(ROOT, "tools", "DESIGN.template.md")).read()  # the document with R2_* placeholders
missing = []
for k in sorted(rep, key=len, reverse=True):
    if k in s:
        s = s.replace(k, rep[k])
    else:
        missing.append(k)
open(path, "w").write(s)
import re
print("left:", sorted(set(re.findall(r"R2_[A-Z0-9_]+", s))), "unused:", missing)
