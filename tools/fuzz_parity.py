#!/usr/bin/env python
"""parity soak: random footprints (3 - 12 vertices, 3 - 70 cells, convex and not), paddings, map sizes (ragged pitches, tiny
maps), densities and pose sets — through the precompute, the extraction, the list kernel, BOTH batched kernels, the row-band
call and the ordered-stripes recipe, each against the CPU oracle.  One JSON line per configuration, a summary at the end.
usage (under gpurun): python tools/fuzz_parity.py [configs] [seed] [big]"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import fixtures as fx  # noqa: E402
import dpose_b200 as dp  # noqa: E402
from dpose_b200.dist import row_band, shard_bounds  # noqa: E402
from oracle import capi  # noqa: E402

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 40
BIG = len(sys.argv) > 3 and sys.argv[3] == "big"  # tables of 200 - 480 cells a side (the cap is 512): texture / no-shared-memory modes, Meijster row pass
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
nt = capi.max_threads()
TILE_MIN = 148 * 32 + 1
fails, t_start = [], time.time()


def check(tag, cond, detail=""):
    if not cond:
        fails.append((tag, detail))
    return bool(cond)


for k in range(n_cfg):
    radius = float(rng.choice([100, 160, 230] if BIG else [3, 5, 8, 12, 18, 26, 40, 55, 70]))
    nv = int(rng.integers(3, 13))
    ang = np.sort(rng.uniform(0, 2 * math.pi, nv))
    if rng.random() < 0.3:
        ang = rng.permutation(ang)  # self-intersecting outlines too (even-odd fill)
    rad = rng.uniform(0.35, 1.0, nv) * radius
    fp = np.stack([np.round(rad * np.cos(ang)), np.round(rad * np.sin(ang))], 1).astype(np.int32)
    if len(np.unique(fp[:, 0])) < 2 or len(np.unique(fp[:, 1])) < 2:
        continue
    pad = int(rng.integers(0, 7))
    sx, sy = (int(rng.integers(900, 2400)), int(rng.integers(900, 2000))) if BIG else (int(rng.integers(40, 1400)), int(rng.integers(40, 1100)))
    dens = float(rng.choice([0.0, 0.01, 0.05, 0.10] if BIG else [0.0, 0.01, 0.05, 0.10, 0.3, 1.0]))
    m = fx.bernoulli_map(sx, sy, dens, int(rng.integers(1 << 30)))
    if rng.random() < 0.3:
        m[rng.random(m.shape) < 0.05] = rng.choice([253, 255, 1, 128])  # non-lethal costs must not count
    rec = {"cfg": k, "vertices": nv, "radius": radius, "pad": pad, "map": [sx, sy], "density": dens}
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    oc = capi.CostData(fp, pad)
    rec["table"] = [oc.rows, oc.cols]
    ok = check("table", np.array_equal(pg.get_data().get_data().view(np.uint32), oc.cost.view(np.uint32)), rec)
    ok &= check("edt", np.array_equal(pg.get_data().stages()[0], oc.edt_sq), rec)
    cm = dp.Costmap(m)
    # extraction: a rotated box somewhere (partly outside), the whole map
    for ring in (fx.rotated_ring(fx.to_rectangle((-int(radius) - 9, -int(radius) - 4), (int(radius) + 7, int(radius) + 11)),
                                 (rng.uniform(-10, sx + 10), rng.uniform(-10, sy + 10), rng.uniform(-3, 3))),
                 fx.to_rectangle(sx - 1, sy - 1)):
        got, want = dp.lethal_cells_within(cm, ring), capi.lethal_cells_within(m, ring)
        ok &= check("lethal", got.shape == want.shape and np.array_equal(got, want), rec)
    # poses inside (the oracle's per-pose box is clamped at the border) -> both batched kernels against the oracle
    reach = pg.reach()
    if sx > 2 * reach + 4 and sy > 2 * reach + 4:
        n = int(rng.choice([1, 37, 300] if BIG else [1, 37, 700, 5000]))
        poses = fx.random_poses(n, sx, sy, int(rng.integers(1 << 30)), margin=float(reach + 1))
        want = oc.get_cost_batch(m, poses, shift=True, nthreads=nt)
        coop = pg.get_cost_batch(cm, poses) if n < TILE_MIN else None
        padded = np.tile(np.array([[1e8, -1e8, 0.5]]), (max(TILE_MIN, n), 1))
        padded[:n] = poses
        tile = pg.get_cost_batch(cm, padded)[:n]
        def agrees(got, tag):
            """within the tolerance of the oracle (shifted-input protocol); where not, the deviation must be the REFERENCE's own
            finite-difference round-off (sums of 1e4 cells: 1e-6 absolute is inside its noise) — adjudicated against the
            long-double evaluation of the same sums, as tests/test_gpu_parity.py::test_full_size_properties_cfg5 does"""
            good = fx.tol_ok(got, want)
            if good.all():
                return True
            bad = np.flatnonzero(~good.all(axis=1))
            truth = oc.get_cost_batch_truth(m, poses[bad], nthreads=nt)
            mine, theirs = np.abs(got[bad] - truth), np.abs(want[bad] - truth)
            fine = np.all(mine <= 0.01 * theirs + 1e-9 * np.maximum(1.0, np.abs(truth)), axis=1) & fx.tol_ok(got[bad, 0], want[bad, 0])
            rec.setdefault("adjudicated", {})[tag] = {"poses": int(bad.size), "of": int(n), "reference_noise": int(fine.sum())}
            if not fine.all():
                i = bad[np.flatnonzero(~fine)[0]]
                rec["failure"] = {"kernel": tag, "pose": poses[i].tolist(), "got": got[i].tolist(), "oracle": want[i].tolist(),
                                  "truth": truth[np.flatnonzero(~fine)[0]].tolist(), "footprint": fp.tolist()}
            return bool(fine.all())
        ok &= check("batch tile kernel", agrees(tile, "tile"), rec)
        if coop is not None:
            ok &= check("batch coop kernel", agrees(coop, "coop"), rec)
            scale = np.maximum(1.0, np.abs(tile).max(axis=1, keepdims=True))
            ok &= check("coop vs tile", np.all(np.abs(coop - tile) <= 1e-11 * scale), rec)
        # one pose through the list kernel over its own extraction
        p0 = poses[0]
        cells = dp.lethal_cells_within(cm, fx.rotated_ring(pg.get_data().get_box(), p0))
        c, J = pg.get_cost(p0, cells)
        oc_c, oc_J = oc.get_cost(p0, cells)
        ok &= check("list kernel", fx.tol_ok(c, oc_c) and fx.tol_ok(J, oc_J, rel=1e-5, abs_=1e-5 if max(sx, sy) > 500 else 1e-6).all(), rec)
    # poses anywhere (windows across the border, off the map): finite, the ordered-stripes recipe on a volatile map = the plain call
    n = int(rng.choice([6000, 20000, 60000]))
    poses = np.stack([rng.uniform(-30, sx + 30, n), rng.uniform(-30, sy + 30, n), rng.uniform(-7, 7, n)], 1)
    plain = pg.get_cost_batch(cm, poses)
    ok &= check("finite", np.isfinite(plain).all() and (plain[:, 0] >= 0).all(), rec)
    world = int(rng.choice([1, 2, 3, 8]))
    order = dp.order_poses_by_tile(poses, sx, sy)
    ordered = np.ascontiguousarray(poses[order])
    d_p = torch.from_numpy(ordered).cuda()
    d_o = torch.full((n, 4), -1.0, dtype=torch.float64, device="cuda")
    cm.set_volatile(True)
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        if hi > lo:
            band = row_band(min(max(ordered[lo:hi, 1].min(), 0), sy - 1), min(max(ordered[lo:hi, 1].max(), 0), sy - 1), reach, sy)
            pg.get_cost_batch_device(cm, d_p.data_ptr() + 24 * lo, hi - lo, d_o.data_ptr() + 32 * lo, rows=band)
    torch.cuda.synchronize()
    cm.set_volatile(False)
    got = np.empty_like(plain)
    got[order] = d_o.cpu().numpy()
    # shards of <= 4736 poses run the cooperative kernel (another summation order): to rounding there, bit for bit otherwise
    exact = min(hi - lo for lo, hi in (shard_bounds(n, r, world) for r in range(world))) > 4736
    scale = np.maximum(1.0, np.abs(plain).max(axis=1, keepdims=True))
    ok &= check("ordered stripes", np.array_equal(got, plain) if exact else np.all(np.abs(got - plain) <= 1e-11 * scale), rec)
    rec["ok"] = bool(ok)
    print(json.dumps(rec), flush=True)
print(json.dumps({"configs": n_cfg, "failures": len(fails), "first_failures": [(t, str(d)[:200]) for t, d in fails[:5]],
                  "seconds": round(time.time() - t_start, 1)}))
sys.exit(1 if fails else 0)
