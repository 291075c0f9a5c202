"""check_footprint: the oracle restatement against the reference's own tests
(dpose_core/test/check_footprint.cpp) and a brute-force even-odd/outline check; the batched GPU version against
the oracle."""
import math

import numpy as np
import pytest

import fixtures as fx
from oracle import capi

# test cpp:137-147 (Eigen comma init fills row-wise: first all x, then all y)
SQUARE = np.array(list(zip([10, 10, -12, -5], [-10, 8, 9, -7])), dtype=np.int32)
HAND = np.array(list(zip([20, 20, 10, 10, 30, 30, 11, 11, 32, 32, 11, 11, 31, 31, 5, 10, 9, -5, -6],
                         [12, 11, 10, 8, 8, 6, 6, 3, 2, -1, -2, -4, -5, -7, -6, -10, -11, -6, 9])), dtype=np.int32)


def transformed(fp, x, y, th):
    c, s = math.cos(th), math.sin(th)
    f = np.asarray(fp, np.float64)
    return np.stack([fx.round_half_away((c * f[:, 0] + -s * f[:, 1]) + x), fx.round_half_away((s * f[:, 0] + c * f[:, 1]) + y)], 1)


def ros_outline(poly):
    """costmap_2d::Costmap2D::polygonOutlineCells restated: raytraceLine along every edge incl. the closing one"""
    cells = []
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        cells += [tuple(c) for c in capi.raytrace(a, b).tolist()]
        cells.append((int(b[0]), int(b[1])))  # ROS' raytraceLine marks the end cell too
    return cells


def ros_convex_fill(outline):
    """convexFillCells restated: per x column, fill between the smallest and largest outline y"""
    cols = {}
    for x, y in outline:
        lo, hi = cols.get(x, (y, y))
        cols[x] = (min(lo, y), max(hi, y))
    return {(x, y) for x, (lo, hi) in cols.items() for y in range(lo, hi + 1)}


# ---- is_inside (test cpp:23-88) ---------------------------------------------------------------------------
def test_is_inside():
    assert capi.is_inside(30, 40, np.empty((0, 2), np.int32))
    fit = np.array([[0, 0], [29, 0], [29, 39], [0, 39]], np.int32)
    assert capi.is_inside(30, 40, fit)
    # Eigen linear (column-major) index into the 2 x 4 matrix: index k -> vertex k // 2, coordinate k % 2
    for idx, off in [(0, -1), (1, -1), (2, 1), (3, -1), (4, 1), (5, 1), (6, -1), (7, 1)]:
        p = fit.copy()
        p[idx // 2, idx % 2] += off
        assert not capi.is_inside(30, 40, p)


# ---- x_bresenham (test cpp:92-160) ----------------------------------------------------------------------------
@pytest.mark.parametrize("angle", [round(0.1 * k, 1) for k in range(1, 14)])
def test_x_bresenham_follows_raytrace(angle):
    end = np.array([int(math.cos(angle) * 100), int(math.sin(angle) * 100)])  # Eigen << double -> int truncates
    assert end[1] > 0
    br = capi.XBresenham([0, 0], end)
    curr_y = -1
    for x, y in capi.raytrace([0, 0], end).tolist():
        if y != curr_y:
            assert x == br.get_x()
            br.advance_x()
            curr_y = y


def test_x_bresenham_order_invariance():
    assert capi.XBresenham([0, 0], [10, 12]).get_x() == capi.XBresenham([10, 12], [0, 0]).get_x()


# ---- check_footprint (test cpp:162-212) --------------------------------------------------------------------------
@pytest.mark.parametrize("yaw", [round(0.1 * k, 1) for k in range(0, 63)])
def test_check_footprint_regression_vs_ros_fill(yaw):
    fp = transformed(SQUARE, 50, 50, yaw)
    outline = ros_outline(fp)
    area = ros_convex_fill(outline)
    assert area
    m = np.full((100, 100), fx.LETHAL, np.uint8)
    for x, y in area:
        m[y, x] = 0
    assert capi.check_footprint(m, fp)
    for x, y in set(outline):
        m[y, x] = fx.LETHAL
        assert not capi.check_footprint(m, fp), (x, y)
        m[y, x] = 0


def test_check_footprint_small_and_outside():
    m = np.zeros((20, 30), np.uint8)
    m[5, 7] = fx.LETHAL
    assert capi.check_footprint(m, np.empty((0, 2), np.int32))          # empty footprint (:402-404)
    assert capi.check_footprint(m, [[3, 3]]) and not capi.check_footprint(m, [[7, 5]])  # point (:406-408)
    assert not capi.check_footprint(m, [[5, 5], [10, 5]]) and capi.check_footprint(m, [[5, 6], [10, 6]])  # line
    assert not capi.check_footprint(m, [[0, 0], [30, 0], [5, 5]])       # vertex outside the map (:464-465)
    assert capi.check_footprint(m, [[0, 0], [6, 0], [6, 4], [0, 4]])    # clear rectangle
    assert not capi.check_footprint(m, [[2, 2], [12, 2], [12, 9], [2, 9]])  # lethal cell strictly inside
    assert not capi.check_footprint(m, [[2, 2], [12, 2], [12, 9], [2, 9], [2, 2]])  # same, closed


@pytest.mark.parametrize("name", ["hand", "star40", "lshape"])
def test_check_footprint_non_convex_interior(name):
    """every cell strictly inside the polygon (even-odd, by exact point-in-polygon on cell centres) and off the
    outline must be seen; cells far outside must not"""
    fp = {"hand": HAND, "star40": fx.star40(), "lshape": fx.cells_m(fx.LSHAPE_M, 0.05)}[name]
    rng = np.random.default_rng(2)
    for th in rng.uniform(-math.pi, math.pi, 6):
        poly = transformed(fp, 60, 60, th)
        free = np.zeros((120, 120), np.uint8)
        assert capi.check_footprint(free, poly)
        outline = set(ros_outline(poly))
        xs, ys = poly[:, 0].astype(float), poly[:, 1].astype(float)
        n = len(poly)
        for _ in range(60):
            x, y = int(rng.integers(0, 120)), int(rng.integers(0, 120))
            m = free.copy()
            m[y, x] = fx.LETHAL
            got = capi.check_footprint(m, poly)
            if (x, y) in outline:
                assert not got
                continue
            inside = False  # crossing number on the cell centre
            for i in range(n):
                j = (i + 1) % n
                if (ys[i] > y) != (ys[j] > y) and x < (xs[j] - xs[i]) * (y - ys[i]) / (ys[j] - ys[i]) + xs[i]:
                    inside = not inside
            # cells within one cell of the outline are discretisation-dependent: only assert the clear cases
            near = any(abs(x - ox) <= 1 and abs(y - oy) <= 1 for ox, oy in outline)
            if not near:
                assert got == (not inside), (x, y, inside)


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dp():
    import dpose_b200 as dp
    if dp.device_count() == 0:
        pytest.skip("no CUDA device")
    return dp


@pytest.mark.gpu
def test_gpu_check_footprint_reference_regression(dp):  # test cpp:162-212 on the GPU path
    for yaw in [round(0.1 * k, 1) for k in range(0, 63, 4)]:
        fp = transformed(SQUARE, 50, 50, yaw)
        outline = ros_outline(fp)
        m = np.full((100, 100), fx.LETHAL, np.uint8)
        for x, y in ros_convex_fill(outline):
            m[y, x] = 0
        assert dp.check_footprint(m, fp)
        for x, y in list(set(outline))[::5]:
            m[y, x] = fx.LETHAL
            assert not dp.check_footprint(m, fp), (x, y)
            m[y, x] = 0
    m = np.zeros((20, 30), np.uint8)
    m[5, 7] = fx.LETHAL
    for fp in ([], [[3, 3]], [[7, 5]], [[5, 5], [10, 5]], [[5, 6], [10, 6]], [[0, 0], [30, 0], [5, 5]],
               [[0, 0], [6, 0], [6, 4], [0, 4]], [[2, 2], [12, 2], [12, 9], [2, 9]], [[2, 2], [12, 2], [12, 9], [2, 9], [2, 2]]):
        fp = np.array(fp, np.int32).reshape(-1, 2)
        assert dp.check_footprint(m, fp) == capi.check_footprint(m, fp), fp.tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("name,density", [("rect", 0.002), ("hand", 0.001), ("star40", 0.002), ("lshape", 0.004), ("square", 0.01)])
def test_gpu_check_footprint_batch_vs_oracle(dp, name, density):
    fp = {"rect": fx.RECT, "hand": HAND, "star40": fx.star40(), "lshape": fx.cells_m(fx.LSHAPE_M, 0.05), "square": SQUARE}[name]
    m = fx.bernoulli_map(600, 500, density, 21)
    poses = fx.random_poses(20000, 600, 500, 22, margin=-20.0)  # some candidates stick out of the map
    got = dp.check_footprint_batch(dp.Costmap(m), fp, poses)
    want = capi.check_footprint_batch(m, fp, poses, nthreads=capi.max_threads())
    assert 0.02 < want.mean() < 0.98  # the case is not trivial either way
    assert np.array_equal(got, want), int((got != want).sum())


def circle(steps, radius, cx=0, cy=0):
    """the reference's dense benchmark footprint (perf/dpose_costmap.cpp:103-113): rounded points of a circle, with
    duplicate vertices and zero-length edges once steps exceeds the number of outline cells"""
    ang = (math.pi * 2 * np.arange(steps)) / steps
    return np.stack([fx.round_half_away(np.cos(ang) * radius) + cx, fx.round_half_away(np.sin(ang) * radius) + cy], 1).astype(np.int32)


@pytest.mark.parametrize("steps", [65, 100, 1000])
def test_check_footprint_many_vertices_oracle(steps):
    """footprints beyond 64 vertices, as in the reference's benchmark: free disc, then one lethal cell inside / on the
    outline / outside"""
    fp = circle(steps, 50, 100, 100)
    m = np.zeros((200, 200), np.uint8)
    assert capi.check_footprint(m, fp)
    for (x, y), want in (((100, 100), False), ((130, 120), False), ((150, 100), False), ((100, 50), False),
                         ((160, 160), True), ((151, 100), True)):
        m[y, x] = fx.LETHAL
        assert capi.check_footprint(m, fp) == want, (x, y)
        m[y, x] = 0


@pytest.mark.gpu
@pytest.mark.parametrize("steps", [64, 65, 100, 1000])
def test_gpu_check_footprint_many_vertices(dp, steps):
    fp = circle(steps, 50, 100, 100)
    m = np.zeros((200, 200), np.uint8)
    assert dp.check_footprint(m, fp)
    rng = np.random.default_rng(steps)
    for x, y in [(100, 100), (150, 100), (151, 100), (100, 50), (160, 160)] + rng.integers(40, 160, size=(25, 2)).tolist():
        m[y, x] = fx.LETHAL
        assert dp.check_footprint(m, fp) == capi.check_footprint(m, fp), (x, y)
        m[y, x] = 0
    # batched, in the robot frame: the workspace path processes poses in chunks
    fp0 = circle(steps, 20)
    big = fx.bernoulli_map(400, 300, 0.0006, 5)
    poses = fx.random_poses(3000, 400, 300, 6, margin=-10.0)
    got = dp.check_footprint_batch(dp.Costmap(big), fp0, poses)
    want = capi.check_footprint_batch(big, fp0, poses, nthreads=capi.max_threads())
    assert 0.05 < want.mean() < 0.95
    assert np.array_equal(got, want), int((got != want).sum())


@pytest.mark.gpu
def test_gpu_check_footprint_chunked_workspace(dp):
    """a 1 MB workspace budget forces the >64-vertex thread-per-pose kernel to process its 700 poses in 9 chunks
    (fresh process: the budget is read once per process)"""
    import os
    import subprocess
    import sys
    tests_dir = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import numpy as np, fixtures as fx, dpose_b200 as dp\n"
        "from oracle import capi\n"
        "from test_check_footprint import circle\n"
        "fp = circle(200, 20)\n"
        "m = fx.bernoulli_map(400, 300, 0.0006, 5)\n"
        "poses = fx.random_poses(700, 400, 300, 6, margin=-10.0)\n"
        "got = dp.check_footprint_batch(dp.Costmap(m), fp, poses)\n"
        "want = capi.check_footprint_batch(m, fp, poses, nthreads=2)\n"
        "assert 0.05 < want.mean() < 0.95 and np.array_equal(got, want), int((got != want).sum())\n"
        "print('chunked ok')\n" % (os.path.dirname(tests_dir), tests_dir))
    env = dict(os.environ, DPOSE_CHECK_WS_MB="1", DPOSE_CHECK_COOP_MAX="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "chunked ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_gpu_check_footprint_random_polygon_soak(dp):
    """random footprints (convex-ish, self-intersecting, closed, with repeated vertices and horizontal runs), each
    checked at a few hundred random poses through the CTA-per-pose kernel and, in a second batch large enough, the
    thread-per-pose kernel: both must reproduce the oracle exactly"""
    rng = np.random.default_rng(2026)
    m = fx.bernoulli_map(500, 400, 0.0015, 77)
    cm = dp.Costmap(m)
    mismatches = 0
    seen = set()
    for k in range(160):
        n = int(rng.integers(3, 24))
        if k % 2 == 0:
            ang = np.sort(rng.uniform(0, 2 * math.pi, n))
            rad = rng.uniform(2, 40, n)
            fp = np.stack([np.round(rad * np.cos(ang)), np.round(rad * np.sin(ang))], 1)
        else:
            fp = rng.integers(-30, 31, size=(n, 2)).astype(np.float64)
        if k % 5 == 0:
            fp[n // 2] = fp[n // 2 - 1]  # repeated vertex
        if k % 7 == 0:
            fp[-1, 1] = fp[0, 1]  # horizontal closing edge
        if k % 11 == 0:
            fp = np.vstack([fp, fp[:1]])  # closed ring
        fp = fp.astype(np.int32)
        if len(np.unique(fp[:, 1])) < 2:
            continue  # every edge horizontal: undefined upstream (see tests/test_reference_build.py)
        n_poses = 300 if k % 4 else 40000  # 40000 > 32768: thread-per-pose kernel
        poses = np.stack([rng.uniform(-10, 510, n_poses), rng.uniform(-10, 410, n_poses), rng.uniform(-4, 4, n_poses)], 1)
        got = dp.check_footprint_batch(cm, fp, poses)
        want = capi.check_footprint_batch(m, fp, poses, nthreads=capi.max_threads())
        mismatches += int((got != want).sum())
        seen.update(np.unique(want).tolist())
        assert mismatches == 0, (k, fp.tolist(), poses[np.argwhere(got != want).ravel()[:3]].tolist())
    assert seen == {0, 1}
