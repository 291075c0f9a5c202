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
ys, xs = np.nonzero(m == fx.LETHAL)
all_cells = dp.cell_vector_device.upload(np.stack([xs, ys], 1).astype(np.int32))
for name, n in (("rect_pad2", 6000), ("star40_pad2", 1500), ("triangle_pad5", 3000)):
    fp, pad = fx.FOOTPRINTS[name]
    pg, cm = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad)), dp.Costmap(m)
    # boxes inside the map: the reference's natural use (lethal_cells_within(rotated box) + get_cost) is the oracle
    poses = fx.random_poses(n, 700, 500, 32, margin=40.0)
    ref = capi.CostData(fp, pad).get_cost_batch(m, poses, shift=True, nthreads=capi.max_threads())
    assert fx.tol_ok(pg.get_cost_batch(cm, poses), ref).all(), name
    # poses on / beyond the border: lethal_cells_within clamps the box corners there (dpose_costmap.cpp:121-129) and
    # drops in-kernel cells, so the batched call (all lethal cells of the map) is compared with get_cost over all cells
    border = fx.random_poses(60, 700, 500, 33, margin=-5.0)
    border = border[(np.minimum(border[:, 0], 700 - border[:, 0]) < 30) | (np.minimum(border[:, 1], 500 - border[:, 1]) < 30)]
    out = pg.get_cost_batch(cm, border)
    for p, o in zip(border, out):
        c, J = pg.get_cost(p, all_cells)
        assert fx.tol_ok(o[0], c, rel=1e-11, abs_=1e-11) and fx.tol_ok(o[1:], J, rel=1e-9, abs_=1e-8).all(), (name, p)
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
