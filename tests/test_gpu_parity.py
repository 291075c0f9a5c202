"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU
oracle, the cv2-generated goldens and the reference's own unit tests restated.

Bars (BASELINE.json north_star): squared EDT and lethal-cell lists bit-exact (tables too);
cost and Jacobian within |gpu - ref| <= 1e-6 + 1e-5 |ref|.
"""
import math

import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dp():
    import dpose_b200
    assert dpose_b200.device_count() > 0, "no CUDA device: GPU tests cannot run (no CPU fallback)"
    return dpose_b200


@pytest.fixture(scope="module")
def capi():
    from oracle import capi
    return capi


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


TILE_MIN = 40000  # batches of at least this many poses run the thread-per-pose kernel (k_cost_batch_tile) in every build;
                  # batches up to 148 x 1024 lanes / 32 lanes per pose = 4736 poses the cooperative kernel (k_cost_batch_coop)


def batch_both(pg, cm, poses, rel=1e-12, **kw):
    """the same poses through BOTH batched kernels: as they are (cooperative kernel, G lanes per pose) and padded with
    far-away poses (empty windows) up to the thread-per-pose kernel's batch size.  The two differ only by the summation
    order; returns the thread-per-pose result."""
    poses = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(-1, 3))
    n = len(poses)
    assert n < TILE_MIN
    coop = pg.get_cost_batch(cm, poses, **kw)
    padded = np.tile(np.array([[1e8, -1e8, 0.5]]), (TILE_MIN, 1))  # far off every map: empty windows
    padded[:n] = poses
    full = pg.get_cost_batch(cm, padded, **kw)
    assert np.all(full[n:] == 0)
    tile = full[:n]
    # summation order only: the error scales with the sum of the terms' magnitudes — the cost of the row (all terms are
    # positive there) bounds it, a Jacobian entry can be small by cancellation
    scale = np.maximum(1.0, np.abs(tile).max(axis=1, keepdims=True))
    assert np.all(np.abs(coop - tile) <= rel * scale), (np.abs(coop - tile) / scale).max()
    return tile


# ------------------------------------------------------------- precompute ----
@pytest.mark.parametrize("name", list(fx.FOOTPRINTS))
def test_tables_bit_exact_vs_cv2_golden(dp, golden_tables, name):
    fp, pad = fx.FOOTPRINTS[name]
    cd = dp.cost_data(fp, pad)
    g = golden_tables
    d2, outline, mask, pre = cd.stages()
    assert np.array_equal(outline, g[name + ".outline"])
    assert np.array_equal(mask, g[name + ".mask"])
    assert np.array_equal(d2, g[name + ".edt_sq"])  # squared EDT bit-exact
    assert np.array_equal(bits(pre), bits(g[name + ".pre_blur"]))
    assert np.array_equal(bits(cd.get_data()), bits(g[name + ".cost"]))
    assert np.array_equal(cd.get_center(), g[name + ".center"])
    assert np.array_equal(cd.get_box(), g[name + ".box"])


def test_tables_bit_exact_vs_oracle_random_polygons(dp, capi):
    rng = np.random.default_rng(77)
    checked = 0
    for _ in range(120):
        n = int(rng.integers(3, 41))
        r = int(rng.integers(3, 90))
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        rad = rng.uniform(0.2, 1.0, n) * r
        pts = np.stack([np.round(rad * np.cos(ang)), np.round(rad * np.sin(ang))], 1).astype(np.int32)
        if rng.random() < 0.3:
            rng.shuffle(pts)
        pts += rng.integers(-30, 30, 2).astype(np.int32)
        pad = int(rng.integers(0, 12))
        o = capi.CostData(pts, pad)
        cd = dp.cost_data(pts, pad)
        d2, outline, mask, pre = cd.stages()
        assert np.array_equal(outline, o.outline)
        assert np.array_equal(mask, o.mask)
        assert np.array_equal(d2, o.edt_sq)
        assert np.array_equal(bits(pre), bits(o.pre_blur))
        assert np.array_equal(bits(cd.get_data()), bits(o.cost))
        assert np.array_equal(cd.get_center(), o.center) and np.array_equal(cd.get_box(), o.box)
        checked += 1
    assert checked == 120


@pytest.mark.parametrize("scale,pad", [(1.0, 7), (0.45, 3), (1.5, 10)])
def test_large_table_exact(dp, capi, scale, pad):
    # tables larger than any fixture: 325 x 215 and 500 x 325 cells run the Meijster lower-envelope row pass (wider than
    # 160 columns), 150 x 97 the multi-tile warp-shuffle scan; squared distances bit-exact either way
    fp = np.round(np.array([[150, 0], [40, 90], [-120, 60], [-160, -10], [-20, -100], [90, -70]]) * scale).astype(np.int32)
    o = capi.CostData(fp, pad)
    cd = dp.cost_data(fp, pad)
    d2, outline, mask, pre = cd.stages()
    assert np.array_equal(d2, o.edt_sq) and np.array_equal(mask, o.mask)
    assert np.array_equal(bits(cd.get_data()), bits(o.cost))
    assert np.array_equal(bits(dp.cost_data(fp, pad).get_data()), bits(o.cost))  # a second handle: pooled scratch reused


@pytest.mark.parametrize("shift", [(10, 0), (4, 5), (-4, -5), (0, -9)])
def test_translated_triangle(dp, shift):  # dpose_core/test/cost_data.cpp:23-62
    o = dp.cost_data(fx.TRIANGLE, 0)
    t = dp.cost_data(fx.TRIANGLE + np.array(shift, dtype=np.int32), 0)
    assert o.get_data().shape == t.get_data().shape
    assert np.array_equal(o.get_box() + np.array(shift), t.get_box())
    assert np.array_equal(o.get_center() - np.array(shift), t.get_center())
    assert np.array_equal(bits(o.get_data()), bits(t.get_data()))  # stronger than the reference's "lazy compare"


@pytest.mark.parametrize("pad", [1, 5, 10])
def test_padded_triangle(dp, pad):  # dpose_core/test/cost_data.cpp:68-115
    o, p = dp.cost_data(fx.TRIANGLE, 0), dp.cost_data(fx.TRIANGLE, pad)
    assert (o.rows + 2 * pad, o.cols + 2 * pad) == (p.rows, p.cols)
    assert np.array_equal(o.get_box().min(0) - pad, p.get_box().min(0))
    assert np.array_equal(o.get_box().max(0) + pad, p.get_box().max(0))
    assert np.array_equal(o.get_center() + pad, p.get_center())


@pytest.mark.parametrize("angle", [round(0.1 * i, 1) for i in range(1, 30)])
def test_rotated_triangle(dp, angle):  # dpose_core/test/cost_data.cpp:121-154
    pad = 5
    c, s = math.cos(angle), math.sin(angle)
    tri = fx.round_half_away(fx.TRIANGLE.astype(np.float64) @ np.array([[c, -s], [s, c]]).T)
    d = dp.cost_data(tri, pad)
    diff = tri.max(0) - tri.min(0)
    assert d.cols == diff[0] + 1 + 2 * pad and d.rows == diff[1] + 1 + 2 * pad
    assert np.array_equal(tri.min(0) - pad, d.get_box().min(0))
    assert np.array_equal(tri.max(0) + 1 + pad, d.get_box().max(0))


# --------------------------------------------------------------- get_cost ----
def test_get_cost_cells_vs_golden_and_oracle(dp, capi, golden_get_cost):
    g = golden_get_cost
    pgs, ocs = {}, {}
    for i in range(int(g["n"])):
        name = str(g[f"{i}.table"])
        if name not in pgs:
            fp, pad = fx.FOOTPRINTS[name]
            pgs[name] = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
            ocs[name] = capi.CostData(fp, pad)
        pose, cells = g[f"{i}.pose"], g[f"{i}.cells"]
        cost, J = pgs[name].get_cost(pose, cells)
        assert fx.tol_ok(cost, g[f"{i}.cost"]), (i, cost, g[f"{i}.cost"])
        assert np.all(fx.tol_ok(J, g[f"{i}.J"])), (i, J, g[f"{i}.J"])
        oc, oJ = ocs[name].get_cost(pose, cells)
        assert fx.tol_ok(cost, oc) and np.all(fx.tol_ok(J, oJ))
        # far tighter against the analytic long-double evaluation
        tc, tJ = ocs[name].get_cost_truth(pose, cells)
        assert abs(cost - tc) <= 1e-10 * max(1.0, abs(tc))
        assert np.all(np.abs(J - tJ) <= 1e-9 * np.maximum(1.0, np.abs(tJ)))
        # without Jacobian: same cost
        c2, J2 = pgs[name].get_cost(pose, cells, jacobian=False)
        assert c2 == cost and J2 is None
        # device-resident cell vector
        c3, J3 = pgs[name].get_cost(pose, dp.cell_vector_device.upload(cells))
        assert c3 == cost and np.array_equal(J3, J)


@pytest.mark.parametrize("theta", [round(0.1 * i, 1) for i in range(1, 15)])
def test_jacobian_self_consistency(dp, theta):  # dpose_core/test/jacobian_data.cpp:43-78 (+theta)
    pg = dp.pose_gradient(fx.SHIP, dp.pose_gradient.parameter(10))
    cells = np.array([[x, y] for x in range(20) for y in range(20)], np.int32)
    _, J = pg.get_cost([0, 0, theta], cells)
    for a in range(3):
        off = np.zeros(3)
        off[a] = 1e-6
        lo, _ = pg.get_cost(np.array([0, 0, theta]) - off, cells, jacobian=False)
        hi, _ = pg.get_cost(np.array([0, 0, theta]) + off, cells, jacobian=False)
        diff = (hi - lo) / 2e-6
        assert abs((diff - J[a]) / (diff if diff else 1.0)) <= 1e-3


def test_get_cost_edge_cases(dp):
    pg = dp.pose_gradient(fx.RECT)
    c, J = pg.get_cost([0, 0, 0], np.empty((0, 2), np.int32))
    assert c == 0 and np.all(J == 0)
    c, J = pg.get_cost([0, 0, 0.3], np.array([[500, 500], [-400, 3], [2 ** 31 - 1, -2 ** 31]]))
    assert c == 0 and np.all(J == 0)
    # rounding tie: k = 2.5 selects k_upper = 3 (std::round), value T(2, 2) (dpose_core.cpp:177)
    c, _ = pg.get_cost([0.5, 0.5, 0.0], np.array([[-9, -5]], np.int32))
    assert c == float(pg.get_data().get_data()[2, 2])
    # a big list: multi-CTA reduction path, deterministic
    rng = np.random.default_rng(3)
    cells = rng.integers(-14, 15, size=(300000, 2)).astype(np.int32)
    r1, r2 = pg.get_cost([0.2, -0.3, 0.4], cells), pg.get_cost([0.2, -0.3, 0.4], cells)
    assert r1[0] == r2[0] and np.array_equal(r1[1], r2[1])


# ------------------------------------------------------------ lethal cells ----
def test_lethal_reference_tests(dp):  # dpose_core/test/lethal_cells_within.cpp:29-57
    m = np.zeros((200, 200), np.uint8)
    m[100, 100] = fx.LETHAL
    assert dp.lethal_cells_within(m, fx.to_rectangle((99, 99), (101, 101))).tolist() == [[100, 100]]
    m = np.zeros((200, 200), np.uint8)
    m[100, 0:200:20] = fx.LETHAL
    assert len(dp.lethal_cells_within(m, fx.to_rectangle((80, 99), (121, 101)))) == 3
    m = np.zeros((50, 50), np.uint8)
    m[10, 10], m[10, 11], m[10, 12] = 253, 254, 255
    assert dp.lethal_cells_within(m, fx.to_rectangle((0, 0), (49, 49))).tolist() == [[11, 10]]
    assert len(dp.lethal_cells_within(np.zeros((200, 200), np.uint8), fx.to_rectangle(198, 198))) == 0
    assert len(dp.lethal_cells_within(np.zeros((0, 0), np.uint8), fx.to_rectangle(5, 5))) == 0


def test_lethal_goldens_bit_exact(dp, golden_lethal):
    g = golden_lethal
    maps = {"bern200": dp.Costmap(fx.bernoulli_map(200, 200, 0.10, 0)),
            "full200": dp.Costmap(np.full((200, 200), fx.LETHAL, np.uint8))}
    for i in range(int(g["n"])):
        cells = dp.lethal_cells_within(maps[str(g[f"{i}.map"])], g[f"{i}.ring"])
        assert np.array_equal(cells, g[f"{i}.cells"]), i


def test_lethal_random_boxes_vs_oracle(dp, capi):
    rng = np.random.default_rng(9)
    m = fx.bernoulli_map(1777, 1234, 0.10, 21)  # width not a multiple of 16: pitch padding
    cm = dp.Costmap(m)
    for k in range(60):
        w, h = rng.integers(1, 900, 2)
        ring = fx.rotated_ring(fx.to_rectangle((-w // 2, -h // 2), (w // 2, h // 2)),
                               (rng.uniform(-100, 1900), rng.uniform(-100, 1400), rng.uniform(0, 6.3)))
        got = dp.lethal_cells_within(cm, ring)
        ref = capi.lethal_cells_within(m, ring)
        assert np.array_equal(got, ref), (k, ring.tolist(), len(got), len(ref))
    # device-resident result equals the host one
    ring = fx.to_rectangle((3, 5), (1500, 1100))
    dev = dp.lethal_cells_within(cm, ring, on_device=True)
    assert np.array_equal(dev.numpy(), capi.lethal_cells_within(m, ring))


def test_lethal_large_boxes_capacity_estimates(dp, capi):
    """boxes of more than 4 M cells size their result by what the previous extraction on the map returned (+25 %): a
    first call (no estimate), a call whose estimate is far too small (the write pass is repeated with the exact size),
    one whose estimate is too large, and a volatile map — all bit-exact against the oracle"""
    side = 3000
    rng = np.random.default_rng(17)
    m = np.zeros((side, side), np.uint8)
    m[:, : side // 2][rng.random((side, side // 2)) < 0.01] = fx.LETHAL
    m[:, side // 2:][rng.random((side, side - side // 2)) < 0.30] = fx.LETHAL
    m[::7, ::5] = 253  # near misses: only 254 is lethal
    cm = dp.Costmap(m)
    left = fx.to_rectangle((0, 0), (side // 2 - 1, side - 1))          # 4.5 M cells, 1 % lethal
    full = fx.to_rectangle((0, 0), (side - 1, side - 1))                # 9 M cells, mostly the dense half
    tilted = fx.rotated_ring(fx.to_rectangle((-1300, -1100), (1300, 1100)), (1500.0, 1500.0, 0.6))
    for ring in (left, full, left, tilted, full):
        got = dp.lethal_cells_within(cm, ring)
        want = capi.lethal_cells_within(m, ring)
        assert got.shape == want.shape and np.array_equal(got, want)
    cm.set_volatile(True)
    assert np.array_equal(dp.lethal_cells_within(cm, tilted), capi.lethal_cells_within(m, tilted))
    m2 = m.copy()
    m2[1000:1200, 100:900] = fx.LETHAL
    cm.set_volatile(False)
    cm.update(m2[1000:1200, 100:900], 100, 1000)  # dirty rows only: the plane is rebuilt for rows 1000 .. 1199
    assert np.array_equal(dp.lethal_cells_within(cm, full), capi.lethal_cells_within(m2, full))


def test_lethal_volatile_map_rereads_the_box_rows(dp, capi):
    """extraction on a VOLATILE map (a producer on the device rewrites the library's copy behind its back) re-reads the
    rows of the box it is asked for — and sees the new bytes there; the next full-map extraction sees them everywhere"""
    torch = pytest.importorskip("torch")
    m = fx.bernoulli_map(1200, 900, 0.10, 41)
    cm = dp.Costmap(m)
    cm.set_volatile(True)
    full = fx.to_rectangle((0, 0), (1199, 899))
    assert np.array_equal(dp.lethal_cells_within(cm, full), capi.lethal_cells_within(m, full))
    m2 = fx.bernoulli_map(1200, 900, 0.20, 42)
    ptr, pitch = cm.device_ptr()

    class Raw:  # the library's buffer as a torch tensor (CUDA array interface)
        __cuda_array_interface__ = {"shape": (900, pitch), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}
    torch.as_tensor(Raw(), device="cuda")[:, :1200].copy_(torch.from_numpy(m2))
    torch.cuda.synchronize()
    for ring in (fx.to_rectangle((300, 411), (401, 512)),                                         # the plugin's search box: one launch
                 fx.rotated_ring(fx.to_rectangle((-500, -300), (500, 300)), (600.0, 450.0, 0.4)),  # count / scan / write
                 fx.to_rectangle((0, 0), (1199, 0)), full):
        got, want = dp.lethal_cells_within(cm, ring), capi.lethal_cells_within(m2, ring)
        assert got.shape == want.shape and np.array_equal(got, want)


def test_lethal_perf_bench_workload(dp, capi):  # dpose_core/perf/dpose_core.cpp:73-90
    m = np.zeros((200, 200), np.uint8)
    m.reshape(-1)[::2] = fx.LETHAL
    got = dp.lethal_cells_within(m, fx.to_rectangle(198, 198))
    assert np.array_equal(got, capi.lethal_cells_within(m, fx.to_rectangle(198, 198)))


def test_map_update_window(dp, capi):
    m = fx.bernoulli_map(300, 200, 0.10, 4)
    cm = dp.Costmap(m)
    patch = fx.bernoulli_map(37, 21, 0.5, 5)
    m2 = m.copy()
    m2[50:71, 100:137] = patch
    cm.update(patch, 100, 50)
    ring = fx.to_rectangle((0, 0), (299, 199))
    assert np.array_equal(dp.lethal_cells_within(cm, ring), capi.lethal_cells_within(m2, ring))


# ---------------------------------------------------------- batched get_cost ----
@pytest.mark.parametrize("name", ["rect_pad2", "star40_pad2", "lshape_pad2"])
def test_batch_vs_all_cells_golden(dp, golden_batch, name):
    m = fx.bernoulli_map(256, 192, 0.10, int(golden_batch["map_seed"]))
    fp, pad = fx.FOOTPRINTS[name]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    cm = dp.Costmap(m)
    out = batch_both(pg, cm, golden_batch[name + ".poses"])
    assert np.all(fx.tol_ok(out, golden_batch[name + ".out"]))
    # equals the list-based GPU path over all lethal cells of the map
    ys, xs = np.nonzero(m == fx.LETHAL)
    cells = np.stack([xs, ys], 1).astype(np.int32)
    for p, o in zip(golden_batch[name + ".poses"][:6], out[:6]):
        c, J = pg.get_cost(p, cells)
        assert fx.tol_ok(o[0], c, rel=1e-12, abs_=1e-12) and np.all(fx.tol_ok(o[1:], J, rel=1e-11, abs_=1e-11))


@pytest.mark.parametrize("name,n", [("rect_pad2", 20000), ("star40_pad2", 4000), ("ship_pad10", 500)])
def test_batch_config2_vs_oracle(dp, capi, name, n):
    """BASELINE config 2 shape (2000x2000, 10 % lethal), subsampled to what the oracle does in
    seconds: 100 % of poses inside tolerance against the literal reference restatement."""
    m = fx.bernoulli_map(2000, 2000, 0.10, 1)
    poses = fx.random_poses(n, 2000, 2000, 2, margin=100.0 if name == "ship_pad10" else 32.0)
    fp, pad = fx.FOOTPRINTS[name]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    cm = dp.Costmap(m)
    out = batch_both(pg, cm, poses)
    oc = capi.CostData(fp, pad)
    nt = capi.max_threads()
    # gate: 100 % against the reference evaluated on translated-equivalent inputs
    shifted = oc.get_cost_batch(m, poses, shift=True, nthreads=nt)
    ok = fx.tol_ok(out, shifted)
    assert ok.all(), (np.argwhere(~ok)[:5], out[~ok.all(1)][:3], shifted[~ok.all(1)][:3])
    # against the long-double analytic evaluation: 4 orders of magnitude inside the tolerance
    truth = oc.get_cost_batch_truth(m, poses, nthreads=nt)
    assert np.all(np.abs(out - truth) <= 1e-9 * np.maximum(1.0, np.abs(truth)))
    # literal reference (finite differences at coordinates ~1e3): its own round-off occasionally
    # reaches the tolerance on a near-zero J_theta of a many-cell table; every such element must be
    # the reference's deviation from the truth, not ours
    ref = oc.get_cost_batch(m, poses, nthreads=nt)
    bad = ~fx.tol_ok(out, ref)
    assert bad.mean() < 1e-3, bad.mean()
    assert np.all(np.abs(out[bad] - truth[bad]) <= 0.01 * np.abs(ref[bad] - truth[bad]) + 1e-12)
    assert fx.tol_ok(out[:, 0], ref[:, 0]).all()  # the cost itself never needs the protocol
    assert (ref[:, 0] > 0).mean() > 0.9  # the workload is not trivially empty
    # determinism and the no-Jacobian variant
    out2 = batch_both(pg, cm, poses)
    assert np.array_equal(out, out2)
    out3 = batch_both(pg, cm, poses, jacobian=False)
    assert np.array_equal(out3[:, 0], out[:, 0]) and np.all(out3[:, 1:] == 0)
    c1, c2 = pg.get_cost_batch(cm, poses), pg.get_cost_batch(cm, poses[::-1].copy())[::-1]
    assert np.array_equal(c1, c2)  # cooperative kernel: deterministic, independent of the batch position


def test_batch_large_coordinates_protocol(dp, capi):
    """10000-cell coordinates: the reference's FD round-off reaches the tolerance,
    so the gate is against the reference on translated-equivalent inputs; the literal pass rate is
    informational and every literal failure is adjudicated against long double."""
    size = 10000
    rng = np.random.default_rng(4)
    m = np.zeros((size, size), np.uint8)
    # lethal only inside a band to keep host generation cheap
    band = (rng.random((600, size)) < 0.10)
    m[4700:5300][band] = fx.LETHAL
    n = 3000
    poses = np.stack([rng.uniform(32, size - 32, n), rng.uniform(4750, 5250, n), rng.uniform(-math.pi, math.pi, n)], 1)
    pg = dp.pose_gradient(fx.RECT)
    cm = dp.Costmap(m)
    out = batch_both(pg, cm, poses)
    oc = capi.CostData(fx.RECT, 2)
    shifted = oc.get_cost_batch(m, poses, shift=True, nthreads=capi.max_threads())
    assert fx.tol_ok(out, shifted).all()
    literal = oc.get_cost_batch(m, poses, nthreads=capi.max_threads())
    truth = oc.get_cost_batch_truth(m, poses, nthreads=capi.max_threads())
    bad = ~fx.tol_ok(out, literal)
    assert bad.mean() < 0.02
    # where the literal reference disagrees, the GPU is the one next to the truth
    assert np.all(np.abs(out[bad] - truth[bad]) <= 0.01 * np.abs(literal[bad] - truth[bad]) + 1e-12)


def test_batch_border_and_degenerate_poses(dp):
    m = fx.bernoulli_map(64, 48, 0.3, 8)
    pg = dp.pose_gradient(fx.RECT)
    cm = dp.Costmap(m)
    ys, xs = np.nonzero(m == fx.LETHAL)
    cells = np.stack([xs, ys], 1).astype(np.int32)
    poses = np.array([[0, 0, 0.1], [63.9, 47.9, 2.0], [-5.0, 20.0, -1.0], [70.0, 20.0, 0.5], [32, -8, 3.0],
                      [1e6, 1e6, 0.0], [-1e7, 5, 1.0], [31.5, 23.5, 0.0], [31.25, 23.75, math.pi / 2]])
    # (31.5, 23.5, 0): every k is an exact .5 tie and the arithmetic is exact (c = 1, s = 0), so all
    # paths must agree with std::round.  Exact ties under a rotation (e.g. theta = pi/2 at half-integer
    # positions, where cos is 6e-17 rather than 0) depend on the rounding order of k = C + R^T (p - t) and
    # differ between ANY two evaluation orders, including the reference's own; they are not tested.
    out = batch_both(pg, cm, poses)
    for p, o in zip(poses, out):
        c, J = pg.get_cost(p, cells)
        assert fx.tol_ok(o[0], c, rel=1e-12, abs_=1e-12) and np.all(fx.tol_ok(o[1:], J, rel=1e-11, abs_=1e-11)), p
    assert np.all(out[5] == 0) and np.all(out[6] == 0)
    assert len(pg.get_cost_batch(cm, np.empty((0, 3)))) == 0
    full = dp.Costmap(np.full((64, 48), fx.LETHAL, np.uint8))  # every window cell lethal
    yy, xx = np.mgrid[0:64, 0:48]
    allc = np.stack([xx.ravel(), yy.ravel()], 1).astype(np.int32)
    o = batch_both(pg, full, poses[:5])
    for p, r in zip(poses[:5], o):
        c, J = pg.get_cost(p, allc)
        assert fx.tol_ok(r[0], c, rel=1e-12, abs_=1e-12) and np.all(fx.tol_ok(r[1:], J, rel=1e-10, abs_=1e-10))


@pytest.mark.parametrize("name", ["rect_pad2", "star40_pad2", "triangle_pad5"])
def test_batch_clip_invariant_on_a_fully_lethal_map(dp, name):
    """The thread-per-pose kernel has no per-cell bounds test: cells outside the kernel must land on zero
    patches.  Every window cell is lethal here and the angles sit on / next to the thresholds where a strip
    of the kernel rectangle stops being clipped (|cos| or |sin| ~ 0.5 / window side) — any cell that
    escaped the table would read garbage and show up against the list kernel (which does test bounds)."""
    fp, pad = fx.FOOTPRINTS[name]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    side = 320
    full = dp.Costmap(np.full((side, side), fx.LETHAL, np.uint8))
    yy, xx = np.mgrid[0:side, 0:side]
    allc = dp.cell_vector_device.upload(np.stack([xx.ravel(), yy.ravel()], 1).astype(np.int32))
    rng = np.random.default_rng(12)
    eps = [0.0, 1e-12, 1e-6, 1e-3, 0.005, 0.008, 0.0089, 0.009, 0.0178, 0.0179, 0.018, 0.0199, 0.02, 0.0201, 0.03, 0.05]
    thetas = [k * math.pi / 2 + sgn * e for k in range(-2, 3) for e in eps for sgn in (-1, 1)]
    poses = np.array([[rng.uniform(100, 220), rng.uniform(100, 220), th] for th in thetas] +
                     [[rng.uniform(-10, 330), rng.uniform(-10, 330), rng.uniform(-4, 4)] for _ in range(200)])
    out = batch_both(pg, full, poses)
    for p, o in zip(poses, out):
        c, J = pg.get_cost(p, allc)
        assert fx.tol_ok(o[0], c, rel=1e-11, abs_=1e-11) and np.all(fx.tol_ok(o[1:], J, rel=1e-9, abs_=1e-8)), p


@pytest.mark.parametrize("half", [(150, 150), (240, 100), (252, 252)])
def test_batch_clip_invariant_large_tables(dp, half):
    """The fp32 round-off of the thread-per-pose kernel's window clip grows with the square of the table side (|t| up
    to 2 win_max^2 before the cancellation): its margin scales with it (clip_margin_for) and dpose_pg_create caps tables
    at 512 x 512.  Tables of 300 .. 510 cells, a fully lethal map, angles on both sides of clip_thr = 0.5 / win_max
    where a strip starts to be clipped: a dropped valid cell or a cell that escaped the zero ring shows up against the
    list kernel (which tests bounds per cell) and against the cooperative kernel (FP64 clip)."""
    hx, hy = half
    fp = np.array([[hx, hy], [-hx, hy], [-hx, -hy], [hx, -hy]], np.int32)
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(2))
    cd = pg.get_data()
    win_max = math.floor(math.hypot(cd.cols - 3, cd.rows - 3)) + 2
    thr = 0.5 / win_max
    side = 2 * max(hx, hy) + 400
    full = dp.Costmap(np.full((side, side), fx.LETHAL, np.uint8))
    yy, xx = np.mgrid[0:side, 0:side]
    allc = dp.cell_vector_device.upload(np.stack([xx.ravel(), yy.ravel()], 1).astype(np.int32))
    rng = np.random.default_rng(5)
    eps = [thr * k for k in (0.0, 0.5, 0.999, 1.0, 1.001, 1.01, 1.1, 1.5, 2.0, 4.0)] + [5.2e-4, 1e-3]
    thetas = [k * math.pi / 2 + sgn * e for k in (-1, 0, 1, 2) for e in eps for sgn in (-1, 1)] + [0.3, 0.78, 2.1, -2.9]
    c0 = side / 2
    poses = np.array([[c0 + rng.uniform(-30, 30), c0 + rng.uniform(-30, 30), th] for th in thetas])
    out = batch_both(pg, full, poses, rel=1e-10)  # 1e5 .. 2.5e5 terms per sum
    for p, o in zip(poses, out):
        c, J = pg.get_cost(p, allc)
        # 1e5 .. 2.6e5 cells per sum: a Jacobian entry can vanish by symmetry while its terms are ~ cost / cell x lever arm
        # each, so the rounding of the two summation orders is measured against cost x table side
        j_abs = 1e-14 * abs(c) * max(cd.rows, cd.cols)
        assert fx.tol_ok(o[0], c, rel=1e-11, abs_=1e-9) and np.all(fx.tol_ok(o[1:], J, rel=1e-9, abs_=j_abs)), (p, o, c, J)
    with pytest.raises(RuntimeError):  # beyond 512 x 512 cells the table is refused
        dp.pose_gradient(np.array([[300, 10], [-300, 10], [-300, -10], [300, -10]], np.int32))


@pytest.mark.parametrize("n,lanes", [(3000, 32), (9000, 16), (18000, 8), (19000, 1)])
def test_batch_cooperative_kernel_every_group_size(dp, capi, n, lanes):
    """batch sizes on both sides of the kernel switch (up to 4736 poses: cooperative kernel with 32 lanes per pose; above:
    thread-per-pose kernel on CTAs of 128 ... 1024 threads — an A/B build with a smaller DPOSE_COOP_MIN_GROUP runs the
    16- and 8-lane instantiations on the same sizes): against the oracle on translated-equivalent inputs, for a table
    whose windows fit one chunk and one whose windows span several"""
    m = fx.bernoulli_map(1200, 1000, 0.10, 21)
    cm = dp.Costmap(m)
    for name in ("rect_pad2", "ship_pad10"):
        k = n if name == "rect_pad2" else max(n // 20, 1000)
        fp, pad = fx.FOOTPRINTS[name]
        pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
        poses = fx.random_poses(n, 1200, 1000, 22, margin=-20.0 if name == "rect_pad2" else 80.0)
        out = pg.get_cost_batch(cm, poses)
        sub = np.random.default_rng(n).choice(n, k, replace=False)
        inside = (np.minimum(poses[sub, 0], 1200 - poses[sub, 0]) > 80) & (np.minimum(poses[sub, 1], 1000 - poses[sub, 1]) > 80)
        sub = sub[inside]
        want = capi.CostData(fp, pad).get_cost_batch(m, poses[sub], shift=True, nthreads=capi.max_threads())
        assert fx.tol_ok(out[sub], want).all() and (want[:, 0] > 0).mean() > 0.9
        again = pg.get_cost_batch(cm, poses[::-1].copy())[::-1]
        assert np.array_equal(out, again)  # independent of the batch position


def test_batch_ragged_sizes_bit_equal(dp):
    """the thread-per-pose kernel deals 32-pose groups to its warps through a per-CTA counter: batch sizes on both sides of
    every boundary of that deal (one group per CTA, one group per warp, ragged last groups, shares that are not whole
    groups, empty shares of the last CTAs) must reproduce, bit for bit, the prefix of one large batch — a pose's result
    does not depend on the batch size, its position or the warp that evaluated it"""
    m = fx.bernoulli_map(900, 700, 0.10, 5)
    pg = dp.pose_gradient(fx.RECT)
    cm = dp.Costmap(m)
    big = 148 * 1024 * 2 + 4129
    poses = fx.random_poses(big, 900, 700, 77, margin=20.0)
    ref = pg.get_cost_batch(cm, poses)
    assert (ref[:, 0] > 0).mean() > 0.8
    # (up to 148 * 32 = 4736 poses the cooperative kernel runs: another summation order)
    # 131 073 and 2 x 131 072 + 4736: the host-buffer call stages 131 072-pose chunks, and the kernel follows the size of the CALL —
    # a last chunk of 1 or 4736 poses must not fall to the cooperative kernel
    for n in (4737, 4738, 4769, 9473, 20001, 148 * 128 + 31, 131072, 131073, 148 * 1024 - 1, 148 * 1024, 148 * 1024 + 1,
              148 * 1024 + 32 * 148 + 5, 2 * 131072 + 4736, big - 1):
        out = pg.get_cost_batch(cm, poses[:n])
        assert np.array_equal(out, ref[:n]), n
    tail = pg.get_cost_batch(cm, poses[-5000:])  # other positions, other warps
    assert np.array_equal(tail, ref[-5000:])


@pytest.mark.parametrize("nx,ny,nt", [(40, 30, 7), (300, 200, 5), (1, 1, 1), (1200, 700, 3)])
def test_sweep_and_cost_only_entry_points(dp, nx, ny, nt):
    """the lattice entry point generates its poses inside the kernel (x0 + dx ix: one rounded product, one rounded sum,
    the same bits as numpy): results must equal the pose-list call on the same lattice bit for bit, through both
    kernels (8400 poses: cooperative, 300000 / 2.5e6: thread per pose, the last one chunked through the host pipeline);
    the cost-only variants return column 0"""
    m = fx.bernoulli_map(700, 500, 0.08, 3)
    pg = dp.pose_gradient(fx.RECT)
    cm = dp.Costmap(m)
    x, y, th = (20.25, 0.55, nx), (-3.0, 0.73, ny), (-3.0, 0.9, nt)
    poses = dp.sweep_poses(x, y, th)
    assert poses.shape == (nx * ny * nt, 3)
    want = pg.get_cost_batch(cm, poses)
    got = pg.get_cost_sweep(cm, x, y, th)
    assert got.shape == (nt, ny, nx, 4) and np.array_equal(got.reshape(-1, 4), want)
    assert (want[:, 0] > 0).mean() > 0.3 or nx * ny * nt == 1
    costs = pg.get_cost_sweep(cm, x, y, th, cost_only=True)
    assert costs.shape == (nt, ny, nx) and np.array_equal(costs.ravel(), want[:, 0])
    assert np.array_equal(pg.get_cost_batch_costs(cm, poses), want[:, 0])
    nojac = pg.get_cost_sweep(cm, x, y, th, jacobian=False)
    assert np.array_equal(nojac[..., 0].ravel(), want[:, 0]) and np.all(nojac[..., 1:] == 0)
    # index order: x fastest, then y, then theta
    k = (nt - 1, ny // 2, nx // 3)
    p = [x[0] + x[1] * k[2], y[0] + y[1] * k[1], th[0] + th[1] * k[0]]
    assert fx.tol_ok(got[k], pg.get_cost_batch(cm, [p])[0], rel=1e-12, abs_=1e-12).all()


def test_sweep_device_entry_and_argument_checks(dp):
    import ctypes as C
    from dpose_b200._lib import lib
    m = fx.bernoulli_map(300, 200, 0.1, 4)
    pg, cm = dp.pose_gradient(fx.RECT), dp.Costmap(m)
    x, y, th = (30.0, 1.0, 64), (30.0, 1.5, 32), (0.0, 0.4, 4)
    want = pg.get_cost_sweep(cm, x, y, th)
    torch = pytest.importorskip("torch")
    d_out = torch.zeros((4, 32, 64, 4), dtype=torch.float64, device="cuda")
    pg.get_cost_sweep(cm, x, y, th, d_out=d_out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want)
    assert lib().dpose_pg_get_cost_sweep(pg._data._h, cm._h, None, None, 1, 0) == 1  # DPOSE_EINVAL
    with pytest.raises(dp.DposeError):
        pg.get_cost_sweep(cm, (0, 1, 1 << 30), (0, 1, 1 << 30), (0, 1, 4), d_out=d_out.data_ptr())  # 2^62 poses: refused


def test_batch_sees_map_updates(dp):
    """the batched kernel reads a derived bit plane: it must follow dpose_map_update and the volatile mode"""
    m = fx.bernoulli_map(300, 200, 0.10, 4)
    poses = fx.random_poses(3000, 300, 200, 6, margin=0.0)
    pg = dp.pose_gradient(fx.RECT)
    cm = dp.Costmap(m)
    before = pg.get_cost_batch(cm, poses)
    patch = fx.bernoulli_map(120, 90, 0.6, 5)
    m2 = m.copy()
    m2[50:140, 100:220] = patch
    cm.update(patch, 100, 50)
    after = pg.get_cost_batch(cm, poses)
    fresh = pg.get_cost_batch(dp.Costmap(m2), poses)
    assert np.array_equal(after, fresh) and not np.array_equal(after, before)
    # several dirty windows before one batch call: the lazy plane rebuild covers the union of their rows
    for (x0, y0, w, h, seed) in [(3, 1, 40, 7, 7), (250, 190, 50, 10, 8), (0, 96, 300, 1, 9)]:
        patch = fx.bernoulli_map(w, h, 0.5, seed)
        m2[y0:y0 + h, x0:x0 + w] = patch
        cm.update(patch, x0, y0)
    assert np.array_equal(pg.get_cost_batch(cm, poses), pg.get_cost_batch(dp.Costmap(m2), poses))
    fresh = pg.get_cost_batch(dp.Costmap(m2), poses)
    cm.set_volatile(True)
    assert np.array_equal(pg.get_cost_batch(cm, poses), fresh)
    assert np.array_equal(pg.get_cost_batch(cm, poses), fresh)


def test_batch_device_pointer_entry(dp):
    torch = pytest.importorskip("torch")
    m = fx.bernoulli_map(512, 512, 0.10, 1)
    poses = fx.random_poses(5000, 512, 512, 2)
    pg = dp.pose_gradient(fx.RECT)
    cm = dp.Costmap(m)
    ref = pg.get_cost_batch(cm, poses)
    d_p = torch.from_numpy(poses).cuda()
    d_o = torch.empty((len(poses), 4), dtype=torch.float64, device="cuda")
    pg.get_cost_batch_device(cm, d_p.data_ptr(), len(poses), d_o.data_ptr(),
                             stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_o.cpu().numpy(), ref)
    # buffers that are only 8-byte aligned (the fast kernel stores 32-byte results when it can)
    d_p2 = torch.empty(len(poses) * 3 + 1, dtype=torch.float64, device="cuda")
    d_o2 = torch.empty(len(poses) * 4 + 1, dtype=torch.float64, device="cuda")
    d_p2[1:] = d_p.reshape(-1)
    pg.get_cost_batch_device(cm, d_p2.data_ptr() + 8, len(poses), d_o2.data_ptr() + 8,
                             stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_o2[1:].reshape(-1, 4).cpu().numpy(), ref)
    # map built from a device buffer
    cm2 = dp.Costmap(device_ptr=torch.from_numpy(m).cuda().data_ptr(), size_x=512, size_y=512)
    assert np.array_equal(pg.get_cost_batch(cm2, poses), ref)
    assert pg.window_bytes(cm, poses) > 400 * len(poses)


@pytest.mark.parametrize("n", [300, 20000])  # cooperative kernel / thread-per-pose kernel
def test_batch_row_band(dp, n):
    """dpose_pg_get_cost_batch_device_rows: only lethal cells of the band's rows count — the plain call on a map whose
    other rows are free (to rounding where the band cuts a pose's window: the kernel coordinates are formed from the
    window's origin, which the cut moves; bit for bit everywhere else); poses whose reach lies inside the band get the
    bits of the plain call on the whole map; and of a VOLATILE map only the band's rows are re-read in front of the launch"""
    torch = pytest.importorskip("torch")
    from dpose_b200.dist import reach_of_box, row_band
    m = fx.bernoulli_map(640, 500, 0.10, 21)
    pg = dp.pose_gradient(fx.RECT)
    reach = pg.reach()
    assert reach == reach_of_box(pg.get_data().get_box())
    poses = fx.random_poses(n, 640, 500, 22, margin=-10.0)
    d_p = torch.from_numpy(poses).cuda()
    d_o = torch.empty((n, 4), dtype=torch.float64, device="cuda")

    def run(cm, rows=None):
        d_o.fill_(-1.0)
        pg.get_cost_batch_device(cm, d_p.data_ptr(), n, d_o.data_ptr(), rows=rows)
        torch.cuda.synchronize()
        return d_o.cpu().numpy()
    cm = dp.Costmap(m)
    whole = run(cm)
    for lo, hi in ((0, 500), (137, 301), (0, 9), (491, 500), (250, 251)):
        cut = np.zeros_like(m)
        cut[lo:hi] = m[lo:hi]
        got, want = run(cm, (lo, hi)), run(dp.Costmap(cut))
        uncut = (poses[:, 1] - reach >= lo) & (poses[:, 1] + reach <= hi - 1) | (poses[:, 1] + reach < lo) | (poses[:, 1] - reach > hi - 1)
        assert np.array_equal(got[uncut], want[uncut]), (lo, hi)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-9), (lo, hi)
        assert (got[(poses[:, 1] + reach < lo) | (poses[:, 1] - reach > hi - 1)] == 0).all()
    stripe = (poses[:, 1] >= 200.0) & (poses[:, 1] <= 330.5)
    band = row_band(200.0, 330.5, reach, 500)
    assert band == (200 - reach, 332 + reach)
    assert stripe.sum() > n // 5 and np.array_equal(run(cm, band)[stripe], whole[stripe])
    for bad in ((5, 5), (7, 3), (0, 501)):
        with pytest.raises(dp.DposeError):
            pg.get_cost_batch_device(cm, d_p.data_ptr(), n, d_o.data_ptr(), rows=bad)
    # a volatile map (a producer on the device rewrites the library's copy behind its back): the band call re-reads the
    # band's rows, and only them
    vm = dp.Costmap(m)
    vm.set_volatile(True)
    assert np.array_equal(run(vm), whole)
    m2 = fx.bernoulli_map(640, 500, 0.15, 23)
    ptr, pitch = vm.device_ptr()
    assert pitch == 640

    class Raw:  # the library's buffer as a torch tensor (CUDA array interface)
        __cuda_array_interface__ = {"shape": (500, 640), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}
    torch.as_tensor(Raw(), device="cuda").copy_(torch.from_numpy(m2))
    torch.cuda.synchronize()
    cut = np.zeros_like(m2)
    cut[band[0]:band[1]] = m2[band[0]:band[1]]
    got, want = run(vm, band), run(dp.Costmap(cut))
    assert np.array_equal(got[stripe], want[stripe]) and np.allclose(got, want, rtol=1e-9, atol=1e-9)
    assert np.array_equal(got[stripe], run(dp.Costmap(m2))[stripe]) and not np.array_equal(got[stripe], whole[stripe])
    assert np.array_equal(run(vm), run(dp.Costmap(m2)))  # the next plain call re-reads every row


@pytest.mark.parametrize("world", [2, 8])
def test_tile_ordered_stripes_equal_plain_call(dp, world):
    """the strong-scaling recipe of bench.py / INTEGRATION.md 4a on one device, rank by rank: order the pose set by map tile
    (dpose_order_poses_by_tile), cut it into contiguous slices (dist.shard_bounds), evaluate every slice on ITS band of a
    volatile map (dist.row_band, dpose_pg_get_cost_batch_device_rows), scatter the results back — bit for bit the plain
    call on the whole list (a pose's result depends neither on its position in the batch nor on rows outside its reach)"""
    torch = pytest.importorskip("torch")
    from dpose_b200.dist import row_band, shard_bounds
    m = fx.bernoulli_map(1500, 1100, 0.10, 31)
    n = 60000 * world // 2
    poses = fx.random_poses(n, 1500, 1100, 32, margin=-5.0)  # windows across the border included
    pg = dp.pose_gradient(fx.RECT)
    cm = dp.Costmap(m)
    want = pg.get_cost_batch(cm, poses)
    cm.set_volatile(True)
    perm = dp.order_poses_by_tile(poses, 1500, 1100)
    ordered = np.ascontiguousarray(poses[perm])
    d_p = torch.from_numpy(ordered).cuda()
    d_o = torch.full((n, 4), -1.0, dtype=torch.float64, device="cuda")
    rows_read = 0
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        band = row_band(ordered[lo:hi, 1].min(), ordered[lo:hi, 1].max(), pg.reach(), 1100)
        rows_read += band[1] - band[0]
        pg.get_cost_batch_device(cm, d_p.data_ptr() + 24 * lo, hi - lo, d_o.data_ptr() + 32 * lo, rows=band)
    torch.cuda.synchronize()
    got = np.empty_like(want)
    got[perm] = d_o.cpu().numpy()
    assert np.array_equal(got, want)
    assert rows_read < 1100 + world * (2 * pg.reach() + 10)  # the bands tile the map: all ranks together read it about once


def test_batch_multi_gpu_shards(dp):
    """dpose_pg_get_cost_batch_multi: one process, one host buffer, the pose batch sharded contiguously over all GPUs of
    the box.  Bit-equal to the per-device calls on the same shards; for batches whose shards run the thread-per-pose
    kernel also bit-equal to the single-GPU call on the whole batch (that kernel's sums do not depend on the batch
    size; the cooperative kernel of small shards picks its lanes per pose by the shard size, so there the comparison
    is to rounding)."""
    from dpose_b200.core import get_cost_batch_multi
    from dpose_b200.dist import shard_bounds
    n = min(dp.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    m = fx.bernoulli_map(512, 512, 0.10, 1)
    pgs = [dp.pose_gradient(fx.RECT, device=d) for d in range(n)]
    maps = [dp.Costmap(m, device=d) for d in range(n)]
    for count in (10001, 45000 * n + 7):
        poses = fx.random_poses(count, 512, 512, 2)
        got = get_cost_batch_multi(pgs, maps, poses)
        for d in range(n):
            lo, hi = shard_bounds(count, d, n)
            assert np.array_equal(got[lo:hi], pgs[d].get_cost_batch(maps[d], poses[lo:hi])), (count, d)
        ref = pgs[0].get_cost_batch(maps[0], poses)
        if count > 40000 * n:
            assert np.array_equal(got, ref)
        else:
            scale = np.maximum(1.0, np.abs(ref).max(axis=1, keepdims=True))
            assert np.all(np.abs(got - ref) <= 1e-12 * scale)
    print(f"multi-GPU shards on {n} devices: bit-equal")


def test_full_size_properties_cfg5(dp):
    """BASELINE config 5 at full size (10000 x 10000 map, 1e7 poses), checked through size-independent
    properties of the path rather than against the (too slow) CPU oracle:
      * batch-position independence: a permuted batch gives bit-identical per-pose results
      * additivity over the obstacle set: cost / J over the map = over its even-x cells + over its odd-x cells
      * angle periodicity: theta and theta + 2 pi agree
      * the strong-scaling recipe of bench.py: the job ordered by map tile, cut into 8 contiguous slices, every slice
        evaluated on its band of rows of the (volatile) map — bit-identical to the plain call
      * EVERY pose agrees with the oracle (shifted-input protocol)"""
    from oracle import capi
    side, n = 10000, 10_000_000
    rng = np.random.default_rng(4)
    m = np.zeros((side, side), np.uint8)
    for r0 in range(0, side, 1000):  # in slabs: keeps the float64 scratch at 80 MB
        m[r0:r0 + 1000] = np.where(rng.random((1000, side)) < 0.10, fx.LETHAL, 0)
    poses = np.empty((n, 3))
    poses[:, 0] = rng.uniform(32, side - 32, n)
    poses[:, 1] = rng.uniform(32, side - 32, n)
    poses[:, 2] = rng.uniform(-math.pi, math.pi, n)
    pg = dp.pose_gradient(fx.RECT)
    cm = dp.Costmap(m)
    out = pg.get_cost_batch(cm, poses)
    assert np.isfinite(out).all() and (out[:, 0] >= 0).all() and (out[:, 0] > 0).mean() > 0.99
    # batch-position independence (bit-exact)
    perm = rng.permutation(n)
    out_p = pg.get_cost_batch(cm, poses[perm])
    assert np.array_equal(out_p, out[perm])
    del out_p
    # additivity over a partition of the obstacle set
    even = m.copy()
    even[:, 1::2] = 0
    odd = m.copy()
    odd[:, 0::2] = 0
    parts = pg.get_cost_batch(dp.Costmap(even), poses) + pg.get_cost_batch(dp.Costmap(odd), poses)
    del even, odd
    assert np.all(np.abs(parts - out) <= 1e-9 + 1e-11 * np.abs(out))
    del parts
    # periodicity in theta (sin / cos of theta + 2 pi differ in the last bits only)
    sub = slice(0, 1_000_000)
    turned = poses[sub].copy()
    turned[:, 2] += 2.0 * math.pi
    assert np.all(fx.tol_ok(pg.get_cost_batch(cm, turned), out[sub], rel=1e-9, abs_=1e-8))
    del turned
    # the job ordered by map tile, 8 contiguous slices, each on its band of the volatile map (device entry point)
    torch = pytest.importorskip("torch")
    from dpose_b200.dist import row_band, shard_bounds
    order = dp.order_poses_by_tile(poses, side, side)
    ordered = np.ascontiguousarray(poses[order])
    d_p = torch.from_numpy(ordered).cuda()
    d_o = torch.full((n, 4), -1.0, dtype=torch.float64, device="cuda")
    cm.set_volatile(True)
    rows_read = 0
    for r in range(8):
        lo, hi = shard_bounds(n, r, 8)
        band = row_band(ordered[lo:hi, 1].min(), ordered[lo:hi, 1].max(), pg.reach(), side)
        rows_read += band[1] - band[0]
        pg.get_cost_batch_device(cm, d_p.data_ptr() + 24 * lo, hi - lo, d_o.data_ptr() + 32 * lo, rows=band)
    torch.cuda.synchronize()
    cm.set_volatile(False)
    assert np.array_equal(d_o.cpu().numpy(), out[order]) and rows_read < side + 8 * (2 * pg.reach() + 10)
    del d_p, d_o, ordered, order
    # EVERY pose against the oracle (translated-equivalent inputs): the CPU port does ~5e6 poses/s
    # on the GPU box's cores, so the full batch takes a few seconds
    oc = capi.CostData(fx.RECT, 2)
    nt = capi.max_threads()
    worst = 0.0
    for lo in range(0, n, 2_000_000):
        ref = oc.get_cost_batch(m, poses[lo:lo + 2_000_000], shift=True, nthreads=nt)
        got = out[lo:lo + 2_000_000]
        assert fx.tol_ok(got, ref).all(), lo
        worst = max(worst, float(np.max(np.abs(got - ref) / (1e-6 + 1e-5 * np.abs(ref)))))
    assert worst < 0.2  # the largest deviation uses under a fifth of the tolerance
    # literal reference (no shift) on a slice: at coordinates ~1e4 its own FD round-off reaches the tolerance on
    # ~1 % of the poses; each of those must be the reference's deviation from the long-double truth
    sl = slice(0, 200_000)
    lit = oc.get_cost_batch(m, poses[sl], nthreads=nt)
    truth = oc.get_cost_batch_truth(m, poses[sl], nthreads=nt)
    bad = ~fx.tol_ok(out[sl], lit)
    assert bad.any(axis=1).mean() < 0.05
    assert np.all(np.abs(out[sl][bad] - truth[bad]) <= 0.01 * np.abs(lit[bad] - truth[bad]) + 1e-12)
    assert fx.tol_ok(out[sl][:, 0], lit[:, 0]).all()  # the cost itself never needs the protocol


def test_batch_random_footprints_all_table_modes(dp, capi):
    """random footprints from a few cells to 150 cells across: the tile kernel's 8-replica, 4-replica, 2-replica (two planes
    of halves), single-replica (texture) and no-shared-memory (both halves through the texture unit) modes, multi-chunk
    windows included, against the oracle on translated-equivalent inputs"""
    rng = np.random.default_rng(99)
    m = fx.bernoulli_map(900, 800, 0.05, 13)
    cm = dp.Costmap(m)
    sizes = []
    for k, radius in enumerate([3, 6, 10, 16, 24, 35, 50, 75]):
        n = int(rng.integers(3, 10))
        ang = np.sort(rng.uniform(0, 2 * math.pi, n))
        rad = rng.uniform(0.5, 1.0, n) * radius
        fp = np.stack([np.round(rad * np.cos(ang)), np.round(rad * np.sin(ang))], 1).astype(np.int32)
        if len(np.unique(fp[:, 0])) < 2 or len(np.unique(fp[:, 1])) < 2:
            continue
        pad = int(rng.integers(1, 5))
        pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
        oc = capi.CostData(fp, pad)
        sizes.append((oc.rows, oc.cols))
        poses = fx.random_poses(1200 if radius < 40 else 300, 900, 800, 100 + k, margin=float(2 * radius + 10))
        out = batch_both(pg, cm, poses)
        want = oc.get_cost_batch(m, poses, shift=True, nthreads=capi.max_threads())
        ok = fx.tol_ok(out, want)
        assert ok.all(), (fp.tolist(), pad, np.argwhere(~ok)[:4], out[~ok.all(1)][:2], want[~ok.all(1)][:2])
        assert (want[:, 0] > 0).mean() > 0.5
    assert min(r * c for r, c in sizes) < 400 and max(r * c for r, c in sizes) > 8000  # from 8 replicas to no shared-memory copy at all
    assert any(760 <= r * c <= 1519 for r, c in sizes)  # the 2-replica mode (star40's 35 x 55 table runs it on 768 threads)
