#!/usr/bin/env python
"""SASS evidence for the Blackwell-native claims: per kernel of libdpose_b200.so the counts of the instructions the
design relies on.  usage: python tools/sass_evidence.py > profiles/r2_sass_evidence.txt   (no GPU needed)

  LDG.E.256 / ST*.256 : 256-bit global accesses (new on sm_100): tile fetch, patch-record gather, pack loads
  UBLKCP              : cp.async.bulk (TMA bulk copy) staging the coefficient image
  SYNCS               : mbarrier arrive / try_wait
  ACQBULK / PREEXIT   : griddepcontrol.wait / launch_dependents (programmatic dependent launch)
  DFMA/DADD/DMUL      : FP64 arithmetic (the reference computes in double)
  TEX/TLD             : texture-unit gathers of the big-table modes
  SHFL                : warp shuffles (reductions, scans, broadcasts)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dpose_b200", "lib", "libdpose_b200.so")
PATTERNS = collections.OrderedDict([
    ("LDG.256", r"\bLDG\.E(\.\w+)*\.256"), ("STG.256", r"\bSTG\.E(\.\w+)*\.256"), ("LDG.128", r"\bLDG\.E(\.\w+)*\.128"),
    ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("ACQBULK", r"\bACQBULK"), ("PREEXIT", r"\bPREEXIT"),
    ("DFMA", r"\bDFMA"), ("DADD", r"\bDADD"), ("DMUL", r"\bDMUL"), ("TEX/TLD", r"\b(TEX|TLD)\b"), ("SHFL", r"\bSHFL"),
    ("LDS.128", r"\bLDS(\.\w+)*\.128"), ("ATOM/RED", r"\b(ATOMG|RED)\b"), ("BAR", r"\bBAR\.SYNC"),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        kernels[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4,}\*/", line) else 0
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                kernels[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, cubin architectures {arch}")
    print("# columns: " + " ".join(["instr"] + list(PATTERNS)))
    groups = collections.OrderedDict()
    for (mangled, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", "")).replace("dpose::", "")
        base = re.sub(r"<.*", "", short)
        groups.setdefault(base, []).append((short, c))
    for base, items in groups.items():
        tot = collections.Counter()
        for _, c in items:
            tot.update(c)
        print(f"\n{base}: {len(items)} instantiation(s); totals: " +
              " ".join(f"{k}={tot[k]}" for k in ["instructions"] + list(PATTERNS) if tot[k]))
        for short, c in items[:6]:
            print(f"  {short[:90]:90s} " + " ".join(f"{k}={c[k]}" for k in ["instructions"] + list(PATTERNS) if c[k]))
        if len(items) > 6:
            print(f"  ... {len(items) - 6} more instantiations")


if __name__ == "__main__":
    sys.exit(main())
