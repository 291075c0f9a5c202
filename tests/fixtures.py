"""Shared inputs: the reference's own fixture footprints and the BASELINE.md configs."""
import math

import numpy as np

LETHAL = 254

# dpose_core/test/jacobian_data.cpp:16-21
SHIP = np.array([[50, 0], [40, -20], [-30, -20], [-60, 0], [-30, 20], [40, 20]], dtype=np.int32)
# dpose_core/perf/dpose_core.cpp:25-30
ARROW = np.array([[50, 0], [40, -20], [-60, -20], [-30, 0], [-60, 20], [40, 20]], dtype=np.int32)
# dpose_core/test/cost_data.cpp:40-43 (Eigen comma init fills row-wise: x = 2,0,-2 ; y = 0,10,0)
TRIANGLE = np.array([[2, 0], [0, 10], [-2, 0]], dtype=np.int32)
# BASELINE config 1/2/5: 1.0 x 0.6 m rectangle at 0.05 m
RECT_M = np.array([[0.5, 0.3], [-0.5, 0.3], [-0.5, -0.3], [0.5, -0.3]])
RECT = np.array([[10, 6], [-10, 6], [-10, -6], [10, -6]], dtype=np.int32)
# dpose_goal_tolerance/test/dpose_goal_tolerance.cpp:33-34 at 0.05 m
DIAMOND_M = np.array([[0.5, 0.0], [0.0, 0.3], [-0.2, 0.0], [0.0, -0.3]])
# dpose_goal_tolerance/example/example.yaml:3-11 at 0.05 m (example_map.yaml)
LSHAPE_M = np.array([[0.5, 0.1], [-0.5, -0.1], [-0.4, -0.8], [-0.3, -0.75], [-0.3, -0.25], [0.5, -0.1]])


def star40():
    """BASELINE config 3: 40-vertex non-convex gear, 0.5 x 0.3 m half extents at 0.02 m."""
    pts = []
    for i in range(40):
        a = 2.0 * math.pi * i / 40
        r = 1.0 if i % 2 == 0 else 0.6
        pts.append([round(25 * r * math.cos(a)), round(15 * r * math.sin(a))])
    return np.array(pts, dtype=np.int32)


def round_half_away(v):
    v = np.asarray(v, dtype=np.float64)
    return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int32)


def cells_m(xy_m, res):
    return round_half_away(np.asarray(xy_m) / res)


FOOTPRINTS = {
    # name: (cells, padding)
    "ship_pad10": (SHIP, 10),
    "arrow_pad3": (ARROW, 3),
    "triangle_pad0": (TRIANGLE, 0),
    "triangle_pad5": (TRIANGLE, 5),
    "rect_pad2": (RECT, 2),
    "diamond_pad5": (cells_m(DIAMOND_M, 0.05), 5),
    "lshape_pad2": (cells_m(LSHAPE_M, 0.05), 2),
    "star40_pad2": (star40(), 2),
}


def bernoulli_map(size_x, size_y, density, seed):
    """uint8 costmap, row-major [y, x]; lethal (254) with i.i.d. probability `density`."""
    rng = np.random.default_rng(seed)
    return np.where(rng.random((size_y, size_x)) < density, LETHAL, 0).astype(np.uint8)


def random_poses(n, size_x, size_y, seed, margin=32.0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(margin, size_x - margin, n)
    y = rng.uniform(margin, size_y - margin, n)
    th = rng.uniform(-math.pi, math.pi, n)
    return np.stack([x, y, th], axis=1)


def rotated_ring(box, pose):
    """round(T(x,y) R(theta) * ring) as dpose_core/test/lethal_cells_within.cpp:82-86."""
    c, s = math.cos(pose[2]), math.sin(pose[2])
    b = np.asarray(box, dtype=np.float64).reshape(5, 2)
    x = (c * b[:, 0] + -s * b[:, 1]) + pose[0]
    y = (s * b[:, 0] + c * b[:, 1]) + pose[1]
    return np.stack([round_half_away(x), round_half_away(y)], axis=1)


def to_rectangle(a, b):
    """dpose_core.hpp:45-69: to_rectangle(w, h) or to_rectangle(min, max) -> (5,2) ring."""
    if np.ndim(a) == 0:
        lo, hi = (0, 0), (int(a), int(b))
    else:
        lo, hi = a, b
    return np.array([[lo[0], lo[1]], [hi[0], lo[1]], [hi[0], hi[1]], [lo[0], hi[1]], [lo[0], lo[1]]],
                    dtype=np.int32)


def tol_ok(got, ref, rel=1e-5, abs_=1e-6):
    """north_star tolerance: |got-ref| <= 1e-6 + 1e-5 |ref|."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref) <= abs_ + rel * np.abs(ref)
