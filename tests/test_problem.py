"""dpose_goal_tolerance::problem: the oracle restatement against an independent numpy
restatement and the reference's own rostest (dpose_goal_tolerance/test/dpose_goal_tolerance.cpp:52-112);
the batched GPU evaluation against the oracle."""
import math

import numpy as np
import pytest

import fixtures as fx
from oracle import capi

CFG_ALL = dict(lin_tolerance=1, rot_tolerance=1, weight_goal_lin=1, weight_goal_rot=1, weight_start_lin=1,
               weight_start_rot=1)


def line_map():  # test cpp:40-47: lethal row y = 100 on a 200 x 200 map
    m = np.zeros((200, 200), np.uint8)
    m[100, :] = fx.LETHAL
    return m


def np_normalize_angle(a):  # angles::normalize_angle
    r = math.fmod(a + math.pi, 2.0 * math.pi)
    return r + math.pi if r <= 0.0 else r - math.pi


def np_regs(start, goal, cfg):
    """pose_regularization set-up of problem::init (dpose_goal_tolerance.cpp:144-180), independently restated"""
    start, goal = np.asarray(start, float), np.asarray(goal, float)

    def make(wl, wr, s, g):  # :62-74
        w_lin = wl / max(float(np.sum((s[:2] - g[:2]) ** 2)), 1e-3)
        w_rot = wr / max(np_normalize_angle(g[2] - s[2]) ** 2, 1e-3)
        return g.copy(), np.array([w_lin, w_lin, w_rot])
    regs = []
    if cfg["weight_goal_lin"] or cfg["weight_goal_rot"]:
        regs.append(make(cfg["weight_goal_lin"], cfg["weight_goal_rot"],
                         goal + np.array([cfg["lin_tolerance"], 0, cfg["rot_tolerance"]]), goal))
    if cfg["weight_start_lin"] or cfg["weight_start_rot"]:
        s = start.copy()
        diff = start - goal
        diff[2] = np_normalize_angle(diff[2])
        ln, rn = float(np.hypot(diff[0], diff[1])), abs(diff[2])
        if ln != 0 and ln > cfg["lin_tolerance"]:
            s = goal + diff * cfg["lin_tolerance"] / ln
        if rn != 0 and rn > cfg["rot_tolerance"]:
            s[2] = goal[2] + diff[2] * cfg["rot_tolerance"] / rn
        else:
            s[2] = start[2]
        regs.append(make(cfg["weight_start_lin"], cfg["weight_start_rot"], goal, s))
    return regs


def np_eval(cd, cells, goal, regs, x):
    """on_new_x + eval_g + eval_jac_g (:96-110, :238-279) on top of the (separately pinned) oracle get_cost"""
    p = np.asarray(goal, float) + np.asarray(x, float)
    cost, J = cd.get_cost(p, cells)
    J = np.array(J, float)
    for g, norm in regs:
        d = p - g
        d[2] = np_normalize_angle(d[2])
        J += 2 * norm * d
        cost += float(np.dot(norm, d ** 2))
    return cost, J, np.array([x[0] ** 2 + x[1] ** 2, x[2] ** 2]), 2 * np.asarray(x, float)


def test_normalize_angle_known_values():
    for a, want in [(0.0, 0.0), (math.pi / 2, math.pi / 2), (3 * math.pi / 2, -math.pi / 2), (-3 * math.pi / 2, math.pi / 2),
                    (2 * math.pi + 0.25, 0.25), (-7.0, -7.0 + 2 * math.pi)]:
        assert capi.normalize_angle(a) == pytest.approx(want, abs=1e-15)
        assert capi.normalize_angle(a) == np_normalize_angle(a)


@pytest.mark.parametrize("yaw", [round(0.1 * k, 1) for k in range(20)])
def test_oracle_problem_passes_the_reference_rostest(yaw):  # test cpp:67-112
    fp, pad = fx.FOOTPRINTS["diamond_pad5"]
    cd = capi.CostData(fp, pad)
    pr = capi.Problem(cd, line_map())
    pr.init([99, 99, 0], [100, 100, yaw], **CFG_ALL)
    f, grad, _, _ = pr.eval_batch([[0, 0, 0]])
    assert f[0] != 0
    for i in range(3):
        d = np.zeros(3)
        d[i] = 1e-6
        fd = (pr.eval_batch([d])[0][0] - pr.eval_batch([-d])[0][0]) / 2e-6
        assert abs(fd - grad[0][i]) <= 1e-3


@pytest.mark.parametrize("cfg_over", [{}, dict(weight_goal_lin=0, weight_goal_rot=0), dict(weight_start_lin=0, weight_start_rot=0),
                                      dict(lin_tolerance=20, rot_tolerance=0.5, weight_goal_lin=3, weight_start_rot=0.2),
                                      dict(lin_tolerance=16.67)])
def test_oracle_problem_matches_independent_restatement(cfg_over):
    cfg = dict(CFG_ALL)
    cfg.update(cfg_over)
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    cd = capi.CostData(fp, pad)
    m = fx.bernoulli_map(400, 300, 0.10, 11)
    start, goal = [150.3, 120.9, 2.9], [200.25, 150.5, -2.8]
    pr = capi.Problem(cd, m)
    pr.init(start, goal, **cfg)
    # search box (:119-138): axis-aligned square +-(max|box| + |lin_tol|) around the goal, rounded
    h = float(int(float(np.abs(cd.box).max()) + abs(cfg["lin_tolerance"])))  # rectangle<int> << double: truncated
    want_box = fx.round_half_away(np.array([[-h, -h], [h, -h], [h, h], [-h, h], [-h, -h]]) + np.array(goal[:2]))
    assert np.array_equal(pr.search_box, want_box)
    cells = capi.lethal_cells_within(m, pr.search_box)
    assert pr.n_cells == len(cells)
    assert pr.bounds == (cfg["lin_tolerance"] ** 2, cfg["rot_tolerance"] ** 2)
    regs = np_regs(start, goal, cfg)
    rng = np.random.default_rng(0)
    xs = rng.uniform(-1, 1, (50, 3)) * np.array([cfg["lin_tolerance"], cfg["lin_tolerance"], 3.0])
    f, grad, g, jac = pr.eval_batch(xs)
    for i, x in enumerate(xs):
        c, J, gg, jj = np_eval(cd, cells, goal, regs, x)
        assert f[i] == pytest.approx(c, rel=1e-12, abs=1e-12)
        assert np.allclose(grad[i], J, rtol=1e-12, atol=1e-10)
        assert np.array_equal(g[i], gg) and np.array_equal(jac[i], jj)


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dp():
    import dpose_b200 as dp
    if dp.device_count() == 0:
        pytest.skip("no CUDA device")
    return dp


@pytest.mark.gpu
@pytest.mark.parametrize("name,cfg_over,goal", [
    ("rect_pad2", {}, [1000.25, 1000.5, 0.3]),
    ("rect_pad2", dict(lin_tolerance=20, rot_tolerance=1.0), [700.0, 1300.0, -2.0]),
    ("star40_pad2", dict(lin_tolerance=10, weight_start_lin=0, weight_start_rot=0), [500.5, 500.5, 1.0]),
    ("lshape_pad2", dict(lin_tolerance=5, weight_goal_lin=0, weight_goal_rot=0), [10.0, 12.0, 3.0]),  # box clamped at the border
    ("rect_pad2", dict(lin_tolerance=16.67, rot_tolerance=0.7), [700.0, 1300.0, 0.4]),  # fractional tolerance: truncated half size
])
def test_problem_eval_batch_vs_oracle(dp, name, cfg_over, goal):
    cfg = dict(CFG_ALL)
    cfg.update(cfg_over)
    fp, pad = fx.FOOTPRINTS[name]
    m = fx.bernoulli_map(2000, 2000, 0.10, 1)
    start = [goal[0] - 40.0, goal[1] + 25.0, goal[2] + 2.5]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    pr = dp.problem(pg, dp.Costmap(m))
    pr.init(start, goal, cfg)
    oc = capi.Problem(capi.CostData(fp, pad), m)
    oc.init(start, goal, **cfg)
    assert np.array_equal(pr.search_box, oc.search_box)
    assert pr.get_nlp_info() == (3, 2, 3)
    xl, xu, gl, gu = pr.get_bounds_info()
    assert np.all(xl == -1e6) and np.all(xu == 1e6) and np.all(gl == 0) and tuple(gu) == oc.bounds
    rng = np.random.default_rng(3)
    n = 4096
    r, a = rng.uniform(0, cfg["lin_tolerance"], n), rng.uniform(-np.pi, np.pi, n)  # the plugin's polar sampler (:333-357)
    xs = np.stack([np.cos(a) * r, np.sin(a) * r, rng.uniform(0, cfg["rot_tolerance"], n)], axis=1)
    xs[0] = 0.0
    got = pr.eval_batch(xs)
    f, grad, g, jac = oc.eval_batch(xs, nthreads=capi.max_threads())
    assert fx.tol_ok(got["f"], f).all()
    assert fx.tol_ok(got["grad_f"], grad).all()
    assert np.array_equal(got["g"], g) and np.array_equal(got["jac_g"], jac)
    # single-x TNLP callbacks
    assert fx.tol_ok(pr.eval_f(xs[5]), f[5]) and fx.tol_ok(pr.eval_grad_f(xs[5]), grad[5]).all()
    assert np.array_equal(pr.eval_g(xs[5]), g[5]) and np.array_equal(pr.eval_jac_g(xs[5]), jac[5])


@pytest.mark.gpu
@pytest.mark.parametrize("yaw", [round(0.1 * k, 1) for k in range(0, 20, 3)])
def test_problem_gradient_rostest_on_gpu(dp, yaw):  # test cpp:67-112 through the batched evaluation
    fp, pad = fx.FOOTPRINTS["diamond_pad5"]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    pr = dp.problem(pg, dp.Costmap(line_map()))
    pr.init([99, 99, 0], [100, 100, yaw], CFG_ALL)
    xs = np.zeros((7, 3))
    for i in range(3):
        xs[1 + 2 * i, i] = 1e-6
        xs[2 + 2 * i, i] = -1e-6
    out = pr.eval_batch(xs)
    assert out["f"][0] != 0
    for i in range(3):
        fd = (out["f"][1 + 2 * i] - out["f"][2 + 2 * i]) / 2e-6
        assert abs(fd - out["grad_f"][0][i]) <= 1e-3
