"""GPU parity against the REFERENCE'S OWN SOURCES (oracle/_ref, see tests/test_reference_build.py): the CUDA path
through the C ABI compared directly with the compiled dpose_core.cpp / dpose_costmap.cpp, without the oracle
restatement in between.  The prebuilt library travels to the GPU box; where it is absent the module is skipped
(the same comparisons then run against the oracle in test_gpu_parity.py)."""
import math
import zlib

import numpy as np
import pytest

import fixtures as fx
from oracle import refapi

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refapi.available(), reason="oracle/_ref did not travel")]


@pytest.fixture(scope="module")
def dp():
    import dpose_b200
    assert dpose_b200.device_count() > 0, "no CUDA device: GPU tests cannot run (no CPU fallback)"
    return dpose_b200


def rng_of(*key):
    return np.random.default_rng(zlib.crc32(repr(key).encode()))


@pytest.mark.parametrize("name", list(fx.FOOTPRINTS))
def test_cost_table_bit_exact(dp, name):
    fp, pad = fx.FOOTPRINTS[name]
    cd, ref = dp.cost_data(fp, pad), refapi.PoseGradient(fp, pad)
    assert np.array_equal(cd.get_data().view(np.uint32), ref.cost.view(np.uint32))
    assert np.array_equal(cd.get_center(), ref.center)
    assert np.array_equal(np.asarray(cd.get_box()).reshape(5, 2), ref.box)


def test_lethal_cells_within_bit_exact(dp):
    sx, sy = 300, 220
    m = fx.bernoulli_map(sx, sy, 0.2, 21)
    cm = dp.Costmap(m)
    rng = rng_of("lethal")
    for k in range(120):
        w, h = rng.integers(1, 60, size=2)
        box = fx.to_rectangle(np.array([-w, -h]), np.array([w, h]))
        ring = fx.rotated_ring(box, [rng.uniform(-20, sx + 20), rng.uniform(-20, sy + 20), rng.uniform(-4, 4)])
        assert np.array_equal(dp.lethal_cells_within(cm, ring), refapi.lethal_cells_within(m, ring))


@pytest.mark.parametrize("name", ["rect_pad2", "star40_pad2", "lshape_pad2", "ship_pad10"])
def test_get_cost_cells(dp, name):
    """pose_gradient::get_cost on explicit cell lists (the reference's own call shape)"""
    fp, pad = fx.FOOTPRINTS[name]
    pg, ref = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad)), refapi.PoseGradient(fp, pad)
    rng = rng_of("cells", name)
    span = max(ref.rows, ref.cols)
    for _ in range(60):
        se2 = np.array([rng.uniform(-6, 6), rng.uniform(-6, 6), rng.uniform(-math.pi, math.pi)])
        cells = rng.integers(-span, span, size=(int(rng.integers(0, 400)), 2)).astype(np.int32)
        c, J = pg.get_cost(se2, cells)
        rc, rJ = ref.get_cost(se2, cells)
        assert fx.tol_ok(c, rc) and fx.tol_ok(J, rJ).all(), (se2, c, rc, J, rJ)
        c0, _ = pg.get_cost(se2, cells, jacobian=False)
        assert fx.tol_ok(c0, ref.get_cost(se2, cells, False)[0])


@pytest.mark.parametrize("name,n", [("rect_pad2", 3000), ("star40_pad2", 800)])
def test_batch_vs_reference_functions(dp, name, n):
    """the batched kernel against lethal_cells_within + get_cost of the compiled reference, pose by pose.  The
    reference's finite differences lose digits with the magnitude of the coordinates, so it is
    evaluated on translated-equivalent inputs (pose and cells moved by floor(x), floor(y)), as in test_gpu_parity."""
    fp, pad = fx.FOOTPRINTS[name]
    side = 700
    m = fx.bernoulli_map(side, side, 0.10, 31)
    poses = fx.random_poses(n, side, side, 32, margin=-5.0)  # a few boxes reach over the border
    pg, ref = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad)), refapi.PoseGradient(fp, pad)
    out = pg.get_cost_batch(dp.Costmap(m), poses)
    want = np.zeros_like(out)
    for i, p in enumerate(poses):
        cells = refapi.lethal_cells_within(m, fx.rotated_ring(ref.box, p))
        shift = np.floor(p[:2]).astype(np.int64)
        c, J = ref.get_cost([p[0] - shift[0], p[1] - shift[1], p[2]], cells - shift.astype(np.int32))
        want[i] = [c, *J]
    # the kernel reads every lethal cell under the table, the reference only those inside the clamped box:
    # poses whose box leaves the map are compared on the cost of the cells both see (test_fallback_kernels.py)
    inside = (poses[:, 0] > 64) & (poses[:, 0] < side - 64) & (poses[:, 1] > 64) & (poses[:, 1] < side - 64)
    ok = fx.tol_ok(out, want)
    assert ok[inside].all(), (np.argwhere(~ok)[:5], out[~ok.all(1)][:3], want[~ok.all(1)][:3])
    assert inside.mean() > 0.6 and (want[inside, 0] > 0).mean() > 0.9


def test_check_footprint_batch(dp):
    sx, sy = 200, 160
    m = fx.bernoulli_map(sx, sy, 0.003, 41)
    cm = dp.Costmap(m)
    rng = rng_of("cf")
    for fp in (fx.RECT, fx.FOOTPRINTS["lshape_pad2"][0], fx.star40()):
        poses = np.stack([rng.uniform(-5, sx + 5, 400), rng.uniform(-5, sy + 5, 400), rng.uniform(-4, 4, 400)], 1)
        got = dp.check_footprint_batch(cm, fp, poses)
        c, s = np.cos(poses[:, 2]), np.sin(poses[:, 2])
        f = np.asarray(fp, np.float64)
        seen = set()
        for i in range(len(poses)):
            moved = np.stack([fx.round_half_away((c[i] * f[:, 0] + -s[i] * f[:, 1]) + poses[i, 0]),
                              fx.round_half_away((s[i] * f[:, 0] + c[i] * f[:, 1]) + poses[i, 1])], 1)
            want = refapi.check_footprint(m, moved)
            assert bool(got[i]) == want, (poses[i], moved.tolist())
            seen.add(want)
        assert seen == {True, False}


@pytest.mark.parametrize("name,cfg_over,goal", [
    ("rect_pad2", {}, [300.25, 200.5, 0.3]),
    ("star40_pad2", dict(lin_tolerance=10, weight_start_lin=0, weight_start_rot=0), [250.5, 250.5, 1.0]),
    ("lshape_pad2", dict(lin_tolerance=5, weight_goal_lin=0, weight_goal_rot=0), [10.0, 12.0, 3.0]),  # box clamped at the border
    ("rect_pad2", dict(lin_tolerance=16.67, rot_tolerance=0.7), [300.0, 200.0, 0.4]),  # fractional tolerance: truncated half size
])
def test_problem_eval_batch_vs_reference_plugin(dp, name, cfg_over, goal):
    """dpose_problem_eval_batch against the NLP callbacks of the compiled, unmodified plugin source
    (dpose_goal_tolerance.cpp:90-292): what Ipopt would be handed, for 2048 displacements at once"""
    cfg = dict(lin_tolerance=1, rot_tolerance=1, weight_goal_lin=1, weight_goal_rot=1, weight_start_lin=1, weight_start_rot=1)
    cfg.update(cfg_over)
    fp, pad = fx.FOOTPRINTS[name]
    m = fx.bernoulli_map(500, 500, 0.10, 1)
    start = [goal[0] - 40.0, goal[1] + 25.0, goal[2] + 2.5]
    pr = dp.problem(dp.pose_gradient(fp, dp.pose_gradient.parameter(pad)), dp.Costmap(m))
    pr.init(start, goal, cfg)
    ref = refapi.Problem(fp, pad, m)
    ref.init(start, goal, **cfg)
    rng = rng_of("problem", name)
    n = 2048
    r, a = rng.uniform(0, cfg["lin_tolerance"], n), rng.uniform(-np.pi, np.pi, n)  # the plugin's polar sampler (:333-357)
    xs = np.stack([np.cos(a) * r, np.sin(a) * r, rng.uniform(0, cfg["rot_tolerance"], n)], axis=1)
    xs[0] = 0.0
    xs[1:600, :2] *= (1.2 * cfg["lin_tolerance"] / np.maximum(np.hypot(xs[1:600, 0], xs[1:600, 1]), 1e-9))[:, None]  # rim of the box
    xs[1:600, 2] = rng.uniform(-np.pi, np.pi, 599)
    got = pr.eval_batch(xs)
    f, grad, g, jac = ref.eval_batch(xs)
    assert fx.tol_ok(got["f"], f).all()
    assert fx.tol_ok(got["grad_f"], grad).all()
    assert np.array_equal(got["g"], g) and np.array_equal(got["jac_g"], jac)
    dims, _, (gl, gu), _ = ref.info()
    assert pr.get_nlp_info() == tuple(dims[:3].tolist())
    assert np.array_equal(pr.get_bounds_info()[3], gu)
