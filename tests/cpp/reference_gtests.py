"""Builds the REFERENCE'S OWN unit tests (dpose_core/test/*.cpp and the plugin's dpose_goal_tolerance/test/*.cpp together with
the plugin's unmodified source, all compiled from /root/reference) twice:

  on_reference : against the reference's own dpose_core sources (oracle/_ref objects)  -> oracle/_ref/tests/ref_<name>
  on_b200      : against THIS repository's drop-in headers (include/dpose_core/*.hpp) and libdpose_b200.so
                                                                                   -> tests/cpp/_build/b200_ref_<name>

Both use the stand-in Eigen / OpenCV / costmap_2d / GoogleTest headers of oracle/refshim (none of those libraries is
installed in this image; oracle/refshim/README.md).  The first build shows that the stand-ins carry the reference's
tests on the reference's code; the second is the drop-in proof: the same test sources, not a restatement, pass on the
CUDA implementation.  Binaries are git-ignored and travel prebuilt to the GPU box, where /root/reference is absent.
TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE_ROOT = os.environ.get("DPOSE_REFERENCE_ROOT", "/root/reference")
TEST_DIR = os.path.join(REFERENCE_ROOT, "dpose_core", "test")
SHIM = os.path.join(ROOT, "oracle", "refshim")
# name -> number of test cases GoogleTest would run (TEST/TEST_F count + TEST_P count x parameters)
SUITES = {"cost_data": 50, "jacobian_data": 28, "lethal_cells_within": 22, "pose_gradient": 2, "check_footprint": 87,
          # the plugin's own test (dpose_goal_tolerance/test/dpose_goal_tolerance.cpp, has its own main): built together with
          # the UNMODIFIED plugin source dpose_goal_tolerance/src/dpose_goal_tolerance.cpp
          "goal_tolerance": 21}
PLUGIN_SRC = os.path.join(REFERENCE_ROOT, "dpose_goal_tolerance", "src")


def _source(name):
    if name == "goal_tolerance":
        return os.path.join(REFERENCE_ROOT, "dpose_goal_tolerance", "test", "dpose_goal_tolerance.cpp")
    return os.path.join(TEST_DIR, name + ".cpp")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def reference_present():
    return all(os.path.exists(_source(n)) for n in SUITES)


def binary(kind, name):
    if kind == "on_reference":
        return os.path.join(ROOT, "oracle", "_ref", "tests", "ref_" + name)
    return os.path.join(ROOT, "tests", "cpp", "_build", "b200_ref_" + name)


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps if os.path.exists(d))


def _shim_headers():
    return [os.path.join(dp, f) for dp, _, fs in os.walk(SHIM) for f in fs]


def build_on_reference():
    """the reference's tests on the reference's sources; returns {name: path} (empty without /root/reference)"""
    if not reference_present():
        return {n: binary("on_reference", n) for n in SUITES if os.path.exists(binary("on_reference", n))}
    from oracle import refapi
    refapi.build()
    objs = [os.path.join(ROOT, "oracle", "_ref", o) for o in ("dpose_core.o", "dpose_costmap.o", "dpose_oracle.o")]
    out = {}
    for name in SUITES:
        src, exe = _source(name), binary("on_reference", name)
        plugin = name == "goal_tolerance"
        extra = ([os.path.join(ROOT, "oracle", "_ref", "dpose_goal_tolerance.o")] if plugin
                 else [os.path.join(SHIM, "gtest", "gtest_main.cpp")])
        if _stale(exe, [src] + objs + extra + _shim_headers()):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            subprocess.run([CXX, "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-I", SHIM,
                            "-I", os.path.join(REFERENCE_ROOT, "dpose_core", "src"), "-I", PLUGIN_SRC, src] + extra + objs +
                           ["-fopenmp", "-lm", "-o", exe], check=True)
        out[name] = exe
    return out


def build_on_b200():
    """the reference's tests on include/dpose_core + libdpose_b200.so; returns {name: path}"""
    if not reference_present():
        return {n: binary("on_b200", n) for n in SUITES if os.path.exists(binary("on_b200", n))}
    from dpose_b200 import build
    lib = build.build()
    include = os.path.join(ROOT, "include")
    headers = [os.path.join(dp, f) for dp, _, fs in os.walk(include) for f in fs] + _shim_headers()
    out = {}
    for name in SUITES:
        src, exe = _source(name), binary("on_b200", name)
        # the plugin's test links the plugin's own, unmodified translation unit — compiled here against the drop-in
        extra = ([os.path.join(PLUGIN_SRC, "dpose_goal_tolerance.cpp")] if name == "goal_tolerance"
                 else [os.path.join(SHIM, "gtest", "gtest_main.cpp")])
        if _stale(exe, [src] + extra + headers):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            # include/ first: <dpose_core/*.hpp> must resolve to the drop-in headers, everything else to the stand-ins
            subprocess.run([CXX, "-std=c++17", "-O2", "-w", "-I", include, "-I", SHIM, "-I", PLUGIN_SRC, src] + extra +
                           ["-L", os.path.dirname(lib), "-ldpose_b200", "-Wl,-rpath,$ORIGIN/../../../dpose_b200/lib", "-o", exe],
                           check=True)
        out[name] = exe
    return out


PERF_SOURCES = ("dpose_core.cpp", "dpose_costmap.cpp")  # dpose_core/perf/, one executable (dpose_core/CMakeLists.txt:76-79)


def perf_binary(kind):
    if kind == "on_reference":
        return os.path.join(ROOT, "oracle", "_ref", "perf", "ref_dpose_core_perf")
    return os.path.join(ROOT, "tests", "cpp", "_build", "b200_dpose_core_perf")


def build_perf(kind):
    """the reference's own Google-Benchmark file pair (perf/dpose_core.cpp + perf/dpose_costmap.cpp), unmodified, on the
    reference's sources or on the drop-in; returns the path or None"""
    exe = perf_binary(kind)
    srcs = [os.path.join(REFERENCE_ROOT, "dpose_core", "perf", f) for f in PERF_SOURCES]
    if not all(os.path.exists(f) for f in srcs):
        return exe if os.path.exists(exe) else None
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if kind == "on_reference":
        from oracle import refapi
        refapi.build()
        objs = [os.path.join(ROOT, "oracle", "_ref", o) for o in ("dpose_core.o", "dpose_costmap.o", "dpose_oracle.o")]
        if _stale(exe, srcs + objs + _shim_headers()):
            subprocess.run([CXX, "-std=c++17", "-O3", "-DNDEBUG", "-ffp-contract=off", "-w", "-I", SHIM,
                            "-I", os.path.join(REFERENCE_ROOT, "dpose_core", "src")] + srcs + objs +
                           ["-fopenmp", "-lm", "-o", exe], check=True)
    else:
        from dpose_b200 import build
        lib = build.build()
        include = os.path.join(ROOT, "include")
        headers = [os.path.join(dp, f) for dp, _, fs in os.walk(include) for f in fs] + _shim_headers()
        if _stale(exe, srcs + headers):
            subprocess.run([CXX, "-std=c++17", "-O3", "-DNDEBUG", "-w", "-I", include, "-I", SHIM] + srcs +
                           ["-L", os.path.dirname(lib), "-ldpose_b200", "-Wl,-rpath,$ORIGIN/../../../dpose_b200/lib",
                            "-o", exe], check=True)
    return exe


def run(exe, timeout=600):
    """-> (returncode, tests ran, tests failed, output)"""
    r = subprocess.run([exe], capture_output=True, text=True, timeout=timeout)
    ran = failed = -1
    for line in r.stdout.splitlines():
        if line.startswith("[==========]"):
            parts = line.replace(".", " ").split()
            ran, failed = int(parts[1]), int(parts[-1])
    return r.returncode, ran, failed, r.stdout[-4000:] + r.stderr[-2000:]


if __name__ == "__main__":
    print(build_on_reference())
    print(build_on_b200())
    print(build_perf("on_reference"), build_perf("on_b200"))
