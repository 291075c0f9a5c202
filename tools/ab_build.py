#!/usr/bin/env python
"""builds A/B variants of libdpose_b200.so HERE (no GPU needed): dpose_b200/lib/ab/<name>.so per variant, then the default
build again.  usage: python tools/ab_build.py name1="-DDPOSE_TILE_CVT=1 ..." name2="..."   (tools/ab_run.sh runs them on the box)"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpose_b200 import build as B  # noqa: E402

ab = os.path.join(ROOT, "dpose_b200", "lib", "ab")
shutil.rmtree(ab, ignore_errors=True)
os.makedirs(ab)
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    os.environ["DPOSE_NVCC_EXTRA"] = flags
    B.build(force=True)
    shutil.copy(B.LIB, os.path.join(ab, name + ".so"))
    print("built", name, flags)
os.environ.pop("DPOSE_NVCC_EXTRA", None)
B.build(force=True)
shutil.copy(B.LIB, os.path.join(ab, "default.so"))
print("default build restored")
