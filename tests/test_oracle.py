"""Pins the CPU oracle (oracle/dpose_oracle.c): against the cv2-generated goldens, against
cv2 live when importable, and against the reference's own unit tests restated
(dpose_core/test/*.cpp)."""
import math

import numpy as np
import pytest

import fixtures as fx
from oracle import capi


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---------------------------------------------------------------- tables ----
@pytest.mark.parametrize("name", list(fx.FOOTPRINTS))
def test_table_matches_cv2_golden_bit_exact(golden_tables, name):
    fp, pad = fx.FOOTPRINTS[name]
    cd = capi.CostData(fp, pad)
    g = golden_tables
    assert np.array_equal(cd.outline, g[name + ".outline"])
    assert np.array_equal(cd.mask, g[name + ".mask"])
    assert np.array_equal(cd.edt_sq, g[name + ".edt_sq"])
    assert np.array_equal(bits(cd.pre_blur), bits(g[name + ".pre_blur"]))
    assert np.array_equal(bits(cd.cost), bits(g[name + ".cost"]))
    assert np.array_equal(cd.center, g[name + ".center"])
    assert np.array_equal(cd.box, g[name + ".box"])
    assert np.float32(cd.min_outer) == g[name + ".min_outer"]


def test_survey_known_answers():
    # SURVEY.md §8(c) known-answer vectors
    cd = capi.CostData(fx.SHIP, 10)
    assert (cd.rows, cd.cols) == (61, 131) and tuple(cd.center) == (70, 30)
    assert abs(cd.min_outer - (-0.30232432)) < 1e-7
    cells = np.array([[x, y] for x in range(20) for y in range(20)])
    for th, cost in ((0.1, 4828.132297528709), (0.7, 6207.677289522671), (1.4, 4742.112729824741)):
        c, _ = cd.get_cost([0, 0, th], cells)
        assert abs(c - cost) < 1e-8
    cd = capi.CostData(fx.ARROW, 3)
    assert (cd.rows, cd.cols) == (47, 117) and tuple(cd.center) == (63, 23)
    c, J = cd.get_cost([0, 0, 0], np.array([[x, y] for x in range(10) for y in range(10)]))
    assert abs(c - 1603.027720451355) < 1e-8
    assert np.allclose(J, [0, 80.0, 360.0], atol=1e-3)
    cd = capi.CostData(fx.RECT, 2)
    assert (cd.rows, cd.cols) == (17, 25) and tuple(cd.center) == (12, 8)
    assert abs(cd.min_outer - (-0.02828427)) < 1e-7


def test_oracle_vs_cv2_live_random_polygons():
    cv2 = pytest.importorskip("cv2")  # noqa: F841
    from oracle import cv_pipeline
    rng = np.random.default_rng(123)
    n_checked = 0
    for _ in range(300):
        n = int(rng.integers(3, 15))
        r = int(rng.integers(3, 61))
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        rad = rng.uniform(0.2, 1.0, n) * r
        pts = np.stack([np.round(rad * np.cos(ang)), np.round(rad * np.sin(ang))], 1).astype(np.int32)
        if rng.random() < 0.3:
            rng.shuffle(pts)  # self-intersecting
        pts += rng.integers(-20, 20, 2).astype(np.int32)
        if len(np.unique(pts, axis=0)) < 3:
            continue
        pad = int(rng.integers(0, 8))
        cv = cv_pipeline.cv_cost_data(pts, pad)
        cd = capi.CostData(pts, pad)
        assert np.array_equal(cv["outline"], cd.outline)
        assert np.array_equal(cv["mask"], cd.mask)
        assert np.array_equal(bits(cv["edt"]), bits(np.sqrt(cd.edt_sq.astype(np.float32))))
        assert np.array_equal(bits(cv["cost"]), bits(cd.cost))
        n_checked += 1
    assert n_checked > 250


# -------------------------------------------- reference test/cost_data.cpp ----
@pytest.mark.parametrize("shift", [(10, 0), (4, 5), (-4, -5), (0, -9)])
def test_translated_triangle(shift):  # test/cost_data.cpp:23-62
    o = capi.CostData(fx.TRIANGLE, 0)
    t = capi.CostData(fx.TRIANGLE + np.array(shift, dtype=np.int32), 0)
    assert (o.rows, o.cols) == (t.rows, t.cols)
    assert np.array_equal(o.box + np.array(shift), t.box)
    assert np.array_equal(o.center - np.array(shift), t.center)


@pytest.mark.parametrize("pad", [1, 5, 10])
def test_padded_triangle(pad):  # test/cost_data.cpp:68-115
    o = capi.CostData(fx.TRIANGLE, 0)
    p = capi.CostData(fx.TRIANGLE, pad)
    assert (o.rows + 2 * pad, o.cols + 2 * pad) == (p.rows, p.cols)
    assert np.array_equal(o.box.min(0) - pad, p.box.min(0))
    assert np.array_equal(o.box.max(0) + pad, p.box.max(0))
    assert np.array_equal(o.center + pad, p.center)


@pytest.mark.parametrize("angle", [round(0.1 * i, 1) for i in range(1, 30)])
def test_rotated_triangle(angle):  # test/cost_data.cpp:121-154
    pad = 5
    c, s = math.cos(angle), math.sin(angle)
    rot = np.array([[c, -s], [s, c]])
    tri = fx.round_half_away(fx.TRIANGLE.astype(np.float64) @ rot.T)
    d = capi.CostData(tri, pad)
    diff = tri.max(0) - tri.min(0)
    assert d.cols == diff[0] + 1 + 2 * pad and d.rows == diff[1] + 1 + 2 * pad
    assert np.array_equal(tri.min(0) - pad, d.box.min(0))
    assert np.array_equal(tri.max(0) + 1 + pad, d.box.max(0))


# ---------------------------------------- reference test/jacobian_data.cpp ----
@pytest.mark.parametrize("theta", [round(0.1 * i, 1) for i in range(1, 15)])
def test_jacobian_self_consistency(theta):  # test/jacobian_data.cpp:43-78 (+ theta)
    cd = capi.CostData(fx.SHIP, 10)
    cells = np.array([[x, y] for x in range(20) for y in range(20)])
    _, J = cd.get_cost([0, 0, theta], cells)
    for a in range(3):
        off = np.zeros(3)
        off[a] = 1e-6
        lo, _ = cd.get_cost(np.array([0, 0, theta]) - off, cells, want_jacobian=False)
        hi, _ = cd.get_cost(np.array([0, 0, theta]) + off, cells, want_jacobian=False)
        diff = (hi - lo) / 2e-6
        assert abs((diff - J[a]) / (diff if diff else 1.0)) <= 1e-3


# ----------------------------------------------------- get_cost goldens ----
def test_get_cost_matches_numpy_golden(golden_get_cost, golden_tables):
    g = golden_get_cost
    cache = {}
    for i in range(int(g["n"])):
        name = str(g[f"{i}.table"])
        if name not in cache:
            cache[name] = capi.CostData(*fx.FOOTPRINTS[name])
        cost, J = cache[name].get_cost(g[f"{i}.pose"], g[f"{i}.cells"])
        # same algorithm in two languages: agreement far inside the north-star tolerance
        assert abs(cost - float(g[f"{i}.cost"])) <= 1e-9 * max(1.0, abs(cost))
        assert np.all(fx.tol_ok(J, g[f"{i}.J"], rel=1e-6, abs_=5e-7)), (i, J, g[f"{i}.J"])
        tc, tJ = cache[name].get_cost_truth(g[f"{i}.pose"], g[f"{i}.cells"])
        assert abs(tc - cost) <= 1e-9 * max(1.0, abs(cost))
        assert np.all(fx.tol_ok(tJ, J)), (i, tJ, J)


def test_get_cost_empty_and_outside():
    cd = capi.CostData(fx.RECT, 2)
    c, J = cd.get_cost([0, 0, 0], np.empty((0, 2), np.int32))
    assert c == 0 and np.all(J == 0)
    c, J = cd.get_cost([0, 0, 0.3], np.array([[500, 500], [-400, 3]]))
    assert c == 0 and np.all(J == 0)


def test_rounding_tie_follows_std_round():
    # dpose_core.cpp:177: k = 2.5 must select k_upper = 3 (half away from zero)
    cd = capi.CostData(fx.RECT, 2)
    # pose (0.5, 0.5, 0): k = center + p - t -> integer + 0.5 on both axes
    cells = np.array([[-9, -5]], np.int32)  # k = (12-9-0.5, 8-5-0.5) = (2.5, 2.5)
    c, _ = cd.get_cost([0.5, 0.5, 0.0], cells)
    # upper = (3,3): c_rel = 0 -> value = T(lo.y=2, lo.x=2)
    assert c == float(cd.cost[2, 2])


# ----------------------------------- reference test/lethal_cells_within.cpp ----
def test_lethal_dot():  # :29-43
    m = np.zeros((200, 200), np.uint8)
    m[100, 100] = fx.LETHAL
    cells = capi.lethal_cells_within(m, fx.to_rectangle((99, 99), (101, 101)))
    assert cells.tolist() == [[100, 100]]


def test_lethal_line():  # :45-57
    m = np.zeros((200, 200), np.uint8)
    m[100, 0:200:20] = fx.LETHAL
    cells = capi.lethal_cells_within(m, fx.to_rectangle((80, 99), (121, 101)))
    assert len(cells) == 3


def test_lethal_only_254_counts():
    m = np.zeros((50, 50), np.uint8)
    m[10, 10], m[10, 11], m[10, 12] = 253, 254, 255
    cells = capi.lethal_cells_within(m, fx.to_rectangle((0, 0), (49, 49)))
    assert cells.tolist() == [[11, 10]]


def test_lethal_empty_map():
    assert len(capi.lethal_cells_within(np.zeros((0, 0), np.uint8), fx.to_rectangle(5, 5))) == 0


def _convex_fill_cells_ros(ring):
    """costmap_2d::Costmap2D::convexFillCells restated from its published algorithm
    (polygonOutlineCells via raytraceLine, sort by x, fill each column min_y..max_y)."""
    outline = []
    for i in range(4):
        outline += [tuple(c) for c in capi.raytrace(ring[i], ring[i + 1]).tolist()]
    cols = {}
    for x, y in outline:
        lo, hi = cols.get(x, (y, y))
        cols[x] = (min(lo, y), max(hi, y))
    return {(x, y) for x, (lo, hi) in cols.items() for y in range(lo, hi + 1)}


@pytest.mark.parametrize("angle", [round(0.1 * i, 1) for i in range(20)])
def test_lethal_regression_vs_convex_fill(angle):  # :61-115
    m = np.full((200, 200), fx.LETHAL, np.uint8)
    ring = fx.rotated_ring(fx.to_rectangle((-20, -10), (30, 20)), (100.0, 100.0, angle))
    cells = capi.lethal_cells_within(m, ring)
    assert len(cells) > 0
    ours = {tuple(c) for c in cells.tolist()}
    assert len(ours) == len(cells)
    assert ours == _convex_fill_cells_ros(ring)


def test_lethal_goldens(golden_lethal):
    g = golden_lethal
    maps = {"bern200": fx.bernoulli_map(200, 200, 0.10, 0), "full200": np.full((200, 200), fx.LETHAL, np.uint8)}
    for i in range(int(g["n"])):
        cells = capi.lethal_cells_within(maps[str(g[f"{i}.map"])], g[f"{i}.ring"])
        assert np.array_equal(cells, g[f"{i}.cells"]), i


def test_raytrace_matches_ros_semantics():
    # end exclusive, max(|dx|,|dy|) cells, 8-connected
    for b, e in (((0, 0), (5, 2)), ((3, 3), (-4, 9)), ((2, 2), (2, 9)), ((1, 1), (6, 6)), ((0, 0), (0, 0))):
        ray = capi.raytrace(b, e)
        assert len(ray) == max(abs(e[0] - b[0]), abs(e[1] - b[1]))
        if len(ray):
            assert tuple(ray[0]) == b
            assert np.all(np.abs(np.diff(ray, axis=0)).max(axis=1) == 1) if len(ray) > 1 else True


# -------------------------------------------- batch (window == list) ----------
@pytest.mark.parametrize("name", ["rect_pad2", "star40_pad2", "lshape_pad2"])
def test_batch_window_equals_all_cells(golden_batch, name):
    """lethal_cells_within(rotated box) + get_cost == get_cost over ALL lethal cells, as
    long as the rotated box lies inside the map.  (When it does not, the reference clamps
    the box CORNERS, dpose_costmap.cpp:121-129, which shears the rectangle and drops cells;
    the batched API is defined over all lethal cells, so those poses are excluded here and
    covered by the GPU tests against the all-cells golden.)"""
    m = fx.bernoulli_map(256, 192, 0.10, int(golden_batch["map_seed"]))
    cd = capi.CostData(*fx.FOOTPRINTS[name])
    poses = golden_batch[name + ".poses"]
    inside = np.array([(lambda r: r[:, 0].min() >= 0 and r[:, 1].min() >= 0 and r[:, 0].max() <= 255
                        and r[:, 1].max() <= 191)(fx.rotated_ring(cd.box, p)) for p in poses])
    assert inside.sum() >= 8
    out = cd.get_cost_batch(m, poses)
    ref = golden_batch[name + ".out"]
    assert np.all(fx.tol_ok(out[inside], ref[inside], rel=1e-6, abs_=5e-7))
    out_mt = cd.get_cost_batch(m, poses, nthreads=4)
    assert np.array_equal(out, out_mt)
    tr = cd.get_cost_batch_truth(m, poses)
    assert np.all(fx.tol_ok(tr[inside], ref[inside]))
    sh = cd.get_cost_batch(m, poses, shift=True)
    assert np.all(fx.tol_ok(sh[inside], ref[inside]))


def test_make_footprint():  # dpose_costmap.cpp:40-55
    assert np.array_equal(capi.make_footprint(fx.RECT_M, 0.05), fx.RECT)
    assert capi.make_footprint([[0.125, -0.125]], 0.05).tolist() == [[3, -3]]  # 2.5 -> 3
    with pytest.raises(RuntimeError):
        capi.make_footprint(fx.RECT_M, 0.0)
