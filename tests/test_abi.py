"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports
every symbol include/dpose_b200.h declares, validates arguments like the reference, and
refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import fixtures as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from dpose_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "dpose_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dpose_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(built)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    from dpose_b200 import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_host_side_functions_match_oracle(built):
    import dpose_b200 as dp
    from oracle import capi
    rng = np.random.default_rng(5)
    for _ in range(200):
        b, e = rng.integers(-50, 50, 2), rng.integers(-50, 50, 2)
        assert np.array_equal(dp.raytrace(b, e), capi.raytrace(b, e))
    assert np.array_equal(dp.make_footprint(fx.RECT_M, 0.05), fx.RECT)
    assert np.array_equal(dp.make_footprint(fx.LSHAPE_M, 0.05), capi.make_footprint(fx.LSHAPE_M, 0.05))
    assert np.array_equal(dp.to_rectangle(3, 4), fx.to_rectangle(3, 4))
    assert np.array_equal(dp.to_rectangle((1, 2), (3, 4)), fx.to_rectangle((1, 2), (3, 4)))


def test_order_poses_by_tile(built):
    """dpose_order_poses_by_tile (host arithmetic: runs without a GPU): the stable order by the 32 x 8-cell tile of the
    pose's position, positions outside the map at the border tiles, NaN last; contiguous slices of the ordered set are
    stripes of the map"""
    import dpose_b200 as dp
    from dpose_b200.dist import shard_bounds
    rng = np.random.default_rng(7)
    for size_x, size_y, n in ((300, 260, 5000), (33, 9, 400), (1, 1, 10), (2000, 2000, 100000)):
        poses = np.stack([rng.uniform(-20, size_x + 20, n), rng.uniform(-20, size_y + 20, n), rng.uniform(-3, 3, n)], axis=1)
        if n >= 400:
            poses[rng.integers(0, n, 7), rng.integers(0, 2, 7)] = np.nan
        perm = dp.order_poses_by_tile(poses, size_x, size_y)
        ntx = (size_x + 31) // 32
        x, y = np.clip(poses[:, 0], 0, size_x - 1), np.clip(poses[:, 1], 0, size_y - 1)
        key = np.floor(y / 8) * ntx + np.floor(x / 32)
        key[np.isnan(poses[:, 0]) | np.isnan(poses[:, 1])] = np.inf
        assert np.array_equal(perm, np.argsort(key, kind="stable"))
    inside = fx.random_poses(100000, 2000, 2000, 3)
    ordered = inside[dp.order_poses_by_tile(inside, 2000, 2000)]
    spans = [ordered[slice(*shard_bounds(len(ordered), r, 8)), 1] for r in range(8)]
    assert all(s.max() - s.min() < 2000 / 8 + 16 for s in spans)            # each slice: an eighth of the rows (+ a tile row)
    assert all(a.max() < b.min() + 8 for a, b in zip(spans, spans[1:]))     # in map order
    with pytest.raises(dp.DposeError):
        dp.order_poses_by_tile(inside, 0, 5)


def test_world_cell_conversions(built):  # dpose_goal_tolerance.cpp:294-330 (Costmap2D::worldToMap / mapToWorld)
    import dpose_b200 as dp
    origin, res, size = (-10.0, 5.0), 0.05, (400, 300)
    world = np.array([[-10.0, 5.0, 0.1], [-9.951, 5.149, -3.0], [9.99, 19.99, 1.0], [10.0, 6.0, 0.0], [-10.01, 6.0, 0.0],
                      [0.0, 20.0, 0.0]])
    cells, ok = dp.to_cell(world, origin, res, size)
    assert ok.tolist() == [True, True, True, False, False, False]
    assert cells[:3].tolist() == [[0.0, 0.0, 0.1], [0.0, 2.0, -3.0], [399.0, 299.0, 1.0]]
    back, ok2 = dp.to_msg(np.array([[0.0, 0.0, 0.1], [12.4, 7.5, 2.0], [-0.1, 3.0, 0.0]]), origin, res)
    assert ok2.tolist() == [True, True, False]
    assert np.allclose(back[0], [-10.0 + 0.025, 5.0 + 0.025, 0.1]) and np.allclose(back[1], [-10.0 + 12.5 * res, 5.0 + 8.5 * res, 2.0])
    # a pose converted to its cell and back lands on that cell's centre
    w = np.array([[3.333, 12.777, 0.5]])
    c, _ = dp.to_cell(w, origin, res, size)
    b, _ = dp.to_msg(c, origin, res)
    assert np.all(np.abs(b[0, :2] - w[0, :2]) <= res / 2 + 1e-12)


def test_error_contract(built):
    import dpose_b200 as dp
    with pytest.raises(RuntimeError, match="resolution must be positive"):  # dpose_costmap.cpp:44-45
        dp.make_footprint(fx.RECT_M, 0.0)
    with pytest.raises(RuntimeError, match="resolution must be positive"):
        dp.make_footprint(fx.RECT_M, -1.0)
    for n in (0, 1, 2):  # dpose_core/test/pose_gradient.cpp:9-17
        with pytest.raises(RuntimeError, match="at least three points"):
            dp.pose_gradient(np.zeros((n, 2), np.int32), dp.pose_gradient.parameter())


def test_no_cpu_fallback(built):
    import dpose_b200 as dp
    if dp.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(dp.DposeError) as ei:
        dp.cost_data(fx.RECT, 2)
    assert ei.value.code == 4  # DPOSE_ENODEV
    with pytest.raises(dp.DposeError):
        dp.Costmap(np.zeros((8, 8), np.uint8))


def test_product_does_not_reference_oracle():
    """the product path must never route through oracle/"""
    bad = []
    for base in ("dpose_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|dpose_oracle|refshim|libdpose_ref|refapi", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_product_library_does_not_link_the_checkers(built):
    """libdpose_b200.so depends on neither the oracle nor the compiled reference (ldd), and exports none of their symbols"""
    import subprocess
    from dpose_b200 import build
    so = build.build()
    deps = subprocess.run(["ldd", so], capture_output=True, text=True, check=True).stdout
    assert "dpose_oracle" not in deps and "dpose_ref" not in deps, deps
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    assert not re.search(r"\b(oracle_|ref_)\w+", syms)


def test_argument_validation_without_a_device(built):
    """NULL / out-of-range arguments are rejected with DPOSE_EINVAL before any CUDA call, on any box"""
    from dpose_b200 import _lib
    L = _lib.lib()
    assert L.dpose_pg_create(None, 4, 2, 0, ctypes.byref(ctypes.c_void_p())) == _lib.EINVAL
    fp = fx.RECT.copy()
    assert L.dpose_pg_create(fp.ctypes.data, 4, 2, 0, None) == _lib.EINVAL
    assert L.dpose_pg_create(fp.ctypes.data, 5000, 2, 0, ctypes.byref(ctypes.c_void_p())) == _lib.EINVAL  # vertex cap
    assert L.dpose_pg_create(fp.ctypes.data, 4, 100000, 0, ctypes.byref(ctypes.c_void_p())) == _lib.EINVAL  # padding cap
    assert b"footprint" in L.dpose_last_error() or b"padding" in L.dpose_last_error()
    assert L.dpose_map_create(None, 8, 8, 8, 0, ctypes.byref(ctypes.c_void_p())) == _lib.EINVAL
    m = np.zeros((8, 8), np.uint8)
    assert L.dpose_map_create(m.ctypes.data, 8, 8, 4, 0, ctypes.byref(ctypes.c_void_p())) == _lib.EINVAL  # pitch < size_x
    assert L.dpose_pg_get_cost_batch(None, None, None, 0, None, 1) == _lib.EINVAL
    assert L.dpose_problem_create(None, None, ctypes.byref(ctypes.c_void_p())) == _lib.EINVAL
    ok = ctypes.c_int(0)
    assert L.dpose_check_footprint(None, fp.ctypes.data, 4, 254, ctypes.byref(ok)) == _lib.EINVAL
    assert L.dpose_world_to_cell(None, 3, 0.0, 0.0, 0.05, 10, 10, None, None) == _lib.EINVAL
    w = np.zeros((1, 3))
    assert L.dpose_world_to_cell(w.ctypes.data, 1, 0.0, 0.0, 0.0, 10, 10, w.ctypes.data, None) == _lib.EINVAL  # res <= 0
    n = ctypes.c_size_t(0)
    assert L.dpose_raytrace(None, None, None, 0, ctypes.byref(n)) == _lib.EINVAL
    # destroying NULL handles is a no-op
    L.dpose_pg_destroy(None)
    L.dpose_map_destroy(None)
    L.dpose_cells_destroy(None)
    L.dpose_problem_destroy(None)
    L.dpose_free(None)
    L.dpose_host_free(None)
