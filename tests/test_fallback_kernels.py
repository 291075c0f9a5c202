"""The batched get_cost has three kernels (thread-per-pose tile kernel, warp-per-pose TMA kernel, generic window
scan); the dispatcher picks by table size.  DPOSE_BATCH_KERNEL forces one (read once per process), so each is
run in its own process here against the oracle on the same inputs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np, fixtures as fx, dpose_b200 as dp
from oracle import capi
m = fx.bernoulli_map(700, 500, 0.12, 31)
for name, n in (("rect_pad2", 6000), ("star40_pad2", 1500), ("triangle_pad5", 3000)):
    fp, pad = fx.FOOTPRINTS[name]
    poses = fx.random_poses(n, 700, 500, 32, margin=-5.0)
    out = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad)).get_cost_batch(dp.Costmap(m), poses)
    ref = capi.CostData(fp, pad).get_cost_batch(m, poses, shift=True, nthreads=capi.max_threads())
    assert fx.tol_ok(out, ref).all(), name
print("ok", dp.launch_count())
"""


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["tile", "tma", "generic"])
def test_each_batch_kernel_matches_the_oracle(kernel):
    env = dict(os.environ)
    env.pop("DPOSE_BATCH_KERNEL", None)
    if kernel != "tile":
        env["DPOSE_BATCH_KERNEL"] = kernel
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
