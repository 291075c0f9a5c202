"""The reference-compatible C++ surface (include/dpose_core/*.hpp) compiled with g++ against
libdpose_b200.so: tests/cpp/test_dpose_core_shim.cpp restates the reference's gtests
(dpose_core/test/*.cpp) on it.  CPU run: host-side contracts only; GPU run: everything."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_dpose_core_shim.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "test_shim")


@pytest.fixture(scope="module")
def shim_binary():
    from dpose_b200 import build
    lib = build.build()
    deps = [SRC, lib] + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "include")) for f in fs]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        libdir = os.path.dirname(lib)
        # -std=c++14: the standard the reference builds with (dpose_core/CMakeLists.txt)
        subprocess.run(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        SRC, "-o", OUT, "-L", libdir, "-ldpose_b200", "-Wl,-rpath," + libdir], check=True)
    return OUT


def test_shim_compiles_and_host_contracts_hold(shim_binary):
    r = subprocess.run([shim_binary, "--cpu-only"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


@pytest.mark.gpu
def test_reference_gtests_restated_on_the_shim(shim_binary):
    r = subprocess.run([shim_binary], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
