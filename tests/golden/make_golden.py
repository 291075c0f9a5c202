"""Generates tests/golden/*.npz.  Run from the repo root IN THE BUILD CONTAINER:

    python tests/golden/make_golden.py

Table goldens come from OpenCV itself (cv2 4.13.0, IPP off) running the reference's
pipeline (oracle/cv_pipeline.py, dpose_core.cpp:55-139).  cost/Jacobian goldens come from
the float64 numpy restatement of dpose_core.cpp:200-260 evaluated on those cv2 tables;
lethal-cell goldens from a brute-force python restatement written here.  The C oracle and
the CUDA path are both tested against these files; nothing here is read from
/root/reference at run time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixtures as fx  # noqa: E402
from oracle import cv_pipeline  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def brute_d2(outline):
    zy, zx = np.nonzero(outline == 0)
    ys, xs = np.mgrid[0:outline.shape[0], 0:outline.shape[1]]
    d2 = (ys[..., None] - zy) ** 2 + (xs[..., None] - zx) ** 2
    return d2.min(axis=-1).astype(np.int32)


def py_raytrace(b, e):
    """pure-python dpose_costmap.cpp:57-104"""
    draw = [e[0] - b[0], e[1] - b[1]]
    dab = [abs(draw[0]), abs(draw[1])]
    major = 1 if dab[1] > dab[0] else 0
    minor = 1 if dab[1] < dab[0] else 0
    if major == minor:
        minor = 1 - major
    den, add = dab[major], dab[minor]
    num = den // 2
    sg = [(d > 0) - (d < 0) for d in draw]
    inc_minor, inc_major = list(sg), list(sg)
    inc_minor[major] = 0
    inc_major[minor] = 0
    cur = [b[0], b[1]]
    out = []
    for _ in range(den):
        out.append(tuple(cur))
        num += add
        if num >= den:
            num -= den
            cur[0] += inc_minor[0]
            cur[1] += inc_minor[1]
        cur[0] += inc_major[0]
        cur[1] += inc_major[1]
    return out


def py_lethal_cells_within(m, ring):
    """pure-python dpose_costmap.cpp:106-184 with (y, x)-sorted output"""
    sy, sx = m.shape
    if sx == 0 or sy == 0:
        return np.empty((0, 2), np.int32)
    ring = [(min(max(int(x), 0), sx - 1), min(max(int(y), 0), sy - 1)) for x, y in ring]
    rays = {}
    for i in range(4):
        for (x, y) in py_raytrace(ring[i], ring[i + 1]):
            lo, hi = rays.get(y, (x, x))
            rays[y] = (min(lo, x), max(hi, x))
    out = []
    for y in sorted(rays):
        lo, hi = rays[y]
        for x in range(lo, hi + 1):
            if m[y, x] == fx.LETHAL:
                out.append((x, y))
    return np.array(out, dtype=np.int32).reshape(-1, 2)


def main():
    # ---- cost tables ------------------------------------------------------
    tables = {}
    for name, (fp, pad) in fx.FOOTPRINTS.items():
        cv = cv_pipeline.cv_cost_data(fp, pad)
        d2 = brute_d2(cv["outline"])
        assert np.array_equal(np.rint(cv["edt"].astype(np.float64) ** 2).astype(np.int32), d2)
        tables[name + ".footprint"] = fp
        tables[name + ".padding"] = np.int32(pad)
        tables[name + ".cost"] = cv["cost"]
        tables[name + ".pre_blur"] = cv["pre_blur"]
        tables[name + ".edt_sq"] = d2
        tables[name + ".outline"] = cv["outline"]
        tables[name + ".mask"] = cv["mask"]
        tables[name + ".center"] = cv["center"]
        tables[name + ".box"] = cv["box"].astype(np.int32)
        tables[name + ".min_outer"] = np.float32(cv["min_outer"])
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **tables)

    # ---- get_cost known answers (list-based) --------------------------------
    gc = {}
    rng = np.random.default_rng(11)
    block20 = np.array([[x, y] for x in range(20) for y in range(20)], dtype=np.int32)
    block10 = np.array([[x, y] for x in range(10) for y in range(10)], dtype=np.int32)
    cases = []
    # dpose_core/test/jacobian_data.cpp:26-41: ship pad 10, 20x20 block, theta 0.1..1.4
    for th in np.arange(0.1, 1.5, 0.1):
        cases.append(("ship_pad10", (0.0, 0.0, float(th)), block20))
    # dpose_core/perf/dpose_core.cpp:45-57
    cases.append(("arrow_pad3", (0.0, 0.0, 0.0), block10))
    # random poses / random cell clouds on every table
    for name in fx.FOOTPRINTS:
        for _ in range(6):
            pose = (float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5)), float(rng.uniform(-3.2, 3.2)))
            cells = rng.integers(-70, 70, size=(int(rng.integers(1, 400)), 2)).astype(np.int32)
            cases.append((name, pose, cells))
    # translated far from the origin (config 2 scale)
    for name in ("rect_pad2", "star40_pad2"):
        for _ in range(4):
            t = rng.uniform(100, 1900, 2)
            pose = (float(t[0]), float(t[1]), float(rng.uniform(-3.2, 3.2)))
            cells = (np.floor(t) + rng.integers(-40, 40, size=(300, 2))).astype(np.int32)
            cases.append((name, pose, cells))
    for i, (name, pose, cells) in enumerate(cases):
        cost, J = cv_pipeline.np_get_cost(tables[name + ".cost"], tables[name + ".center"], pose, cells)
        gc[f"{i}.table"] = np.array(name)
        gc[f"{i}.pose"] = np.array(pose, dtype=np.float64)
        gc[f"{i}.cells"] = cells
        gc[f"{i}.cost"] = np.float64(cost)
        gc[f"{i}.J"] = J
    gc["n"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(OUT, "get_cost.npz"), **gc)

    # ---- lethal cells ------------------------------------------------------
    lc = {}
    k = 0
    m200 = fx.bernoulli_map(200, 200, 0.10, 0)
    full = np.full((200, 200), fx.LETHAL, np.uint8)
    base = fx.to_rectangle((-20, -10), (30, 20))  # test/lethal_cells_within.cpp:66
    for th in np.arange(0.0, 2.0, 0.1):
        ring = fx.rotated_ring(base, (100.0, 100.0, float(th)))
        for mname, m in (("bern200", m200), ("full200", full)):
            lc[f"{k}.map"] = np.array(mname)
            lc[f"{k}.ring"] = ring
            lc[f"{k}.cells"] = py_lethal_cells_within(m, ring)
            k += 1
    # clamping quirks: partially / fully outside boxes
    for ring in (fx.to_rectangle((-30, -30), (10, 12)), fx.to_rectangle((190, 150), (260, 230)),
                 fx.to_rectangle((300, 300), (340, 320)), fx.to_rectangle(198, 198),
                 fx.rotated_ring(base, (5.0, 190.0, 0.7))):
        for mname, m in (("bern200", m200), ("full200", full)):
            lc[f"{k}.map"] = np.array(mname)
            lc[f"{k}.ring"] = np.asarray(ring, dtype=np.int32)
            lc[f"{k}.cells"] = py_lethal_cells_within(m, ring)
            k += 1
    lc["n"] = np.int32(k)
    np.savez_compressed(os.path.join(OUT, "lethal.npz"), **lc)

    # ---- batched get_cost on a small map (window semantics = all lethal cells) ----
    bt = {}
    m = fx.bernoulli_map(256, 192, 0.10, 7)
    ys, xs = np.nonzero(m == fx.LETHAL)
    all_cells = np.stack([xs, ys], axis=1).astype(np.int32)
    for name in ("rect_pad2", "star40_pad2", "lshape_pad2"):
        poses = fx.random_poses(24, 256, 192, 13, margin=-20.0)  # some windows leave the map
        out = np.zeros((len(poses), 4))
        for i, p in enumerate(poses):
            c, J = cv_pipeline.np_get_cost(tables[name + ".cost"], tables[name + ".center"], p, all_cells)
            out[i, 0] = c
            out[i, 1:] = J
        bt[name + ".poses"] = poses
        bt[name + ".out"] = out
    bt["map_seed"] = np.int32(7)
    np.savez_compressed(os.path.join(OUT, "batch.npz"), **bt)
    print("golden written:", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
