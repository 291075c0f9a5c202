"""Builds libdpose_b200.so in-tree with nvcc for sm_100a (B200).

    python -m dpose_b200.build [--force]

The library links cudart statically, so it loads (and exports its symbols) on a box
without a GPU; every compute entry point then fails with DPOSE_ENODEV.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libdpose_b200.so")
SOURCES = ["api.cu", "precompute.cu", "lethal.cu", "get_cost.cu", "get_cost_batch_tile.cu", "get_cost_batch_coop.cu", "problem.cu", "solve.cu",
           "check_footprint.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "cost_math.cuh"), os.path.join(CSRC, "coop_eval.cuh"), os.path.join(HERE, "..", "include", "dpose_b200.h")]
# problem.cu mirrors scalar host arithmetic of the reference (x*x + y*y, regulariser sums): no FMA contraction there,
# so the constraint values are bit-identical to a plain x86-64 build of the reference
PER_FILE_FLAGS = {"problem.cu": ["-fmad=false"], "solve.cu": ["-fmad=false"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objdir = os.path.join(HERE, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CC/CXX wrappers without an OpenMP spec; nvcc only needs the distro g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    procs = []
    objs = []
    extra = os.environ.get("DPOSE_NVCC_EXTRA", "").split()  # e.g. -DDPOSE_TILE_CVT=0 for A/B builds
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + ccbin + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + extra + PER_FILE_FLAGS.get(s, []) + \
              ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc] + ccbin + ["-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                             "-o", LIB] + objs + ["-lpthread", "-ldl", "-lrt"]
    subprocess.run(link, check=True, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
