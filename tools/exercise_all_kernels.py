#!/usr/bin/env python
"""every kernel of the library once at small sizes (no torch: host-buffer entry points only), each result compared with
the oracle — a one-minute parity sweep for A/B builds, in particular the bounds-checking build (-DDPOSE_TILE_CHECK=1,
tools/run_check_build.sh).  Written for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), which is closed
on this GPU pool; the traps of the check build and the oracle comparison stand in for it.
usage (under gpurun): python tools/exercise_all_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import fixtures as fx  # noqa: E402
import dpose_b200 as dp  # noqa: E402
from oracle import capi  # noqa: E402

done = []


def ok(name, cond):
    assert cond, name
    done.append(name)


m = fx.bernoulli_map(300, 260, 0.10, 11)
cm = dp.Costmap(m)
for name, (fp, pad) in (("rect", (fx.RECT, 2)), ("star40", fx.FOOTPRINTS["star40_pad2"]), ("ship", fx.FOOTPRINTS["ship_pad10"])):
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))  # k_precompute (shuffle row pass; Meijster for the wide ship table)
    oc = capi.CostData(fp, pad)
    ok(name + " table", np.array_equal(pg.get_data().get_data().view(np.uint32), oc.cost.view(np.uint32)))
    ok(name + " edt", np.array_equal(pg.get_data().stages()[0], oc.edt_sq))
    # thread-per-pose kernel (8 replicas / two planes / texture mode by table size; 128-thread CTAs), cooperative kernel (200 poses)
    margin = 85.0 if name == "ship" else 36.0  # inside: the oracle's per-pose box is clamped at the border, the batch kernels' is not
    for n in (6000, 200):
        poses = fx.random_poses(n, 300, 260, 3 + n, margin=margin)
        out = pg.get_cost_batch(cm, poses)
        ok(f"{name} batch {n}", fx.tol_ok(out, oc.get_cost_batch(m, poses, nthreads=capi.max_threads())).all())
        edge = pg.get_cost_batch(cm, fx.random_poses(n, 300, 260, 5 + n, margin=-40.0))  # windows across the border and off the map
        ok(f"{name} border batch {n}", np.isfinite(edge).all())
        costs = pg.get_cost_batch_costs(cm, poses)
        ok(f"{name} costs {n}", fx.tol_ok(costs, out[:, 0], rel=1e-9, abs_=1e-9).all())
    if name == "rect":
        poses = fx.random_poses(150000, 300, 260, 5, margin=margin)  # 1024-thread CTAs, two staging chunks
        ok("rect batch 150000", fx.tol_ok(pg.get_cost_batch(cm, poses), oc.get_cost_batch(m, poses, nthreads=capi.max_threads())).all())
        sw = pg.get_cost_sweep(cm, (20.0, 1.5, 90), (20.0, 2.5, 60), (-1.0, 0.5, 4))  # k_sweep_poses
        ok("sweep", np.array_equal(sw.reshape(-1, 4), pg.get_cost_batch(cm, dp.sweep_poses((20.0, 1.5, 90), (20.0, 2.5, 60), (-1.0, 0.5, 4)))))
        # list kernel + extraction (small box: one launch; full map: count / scan / write)
        ring = fx.rotated_ring(pg.get_data().get_box(), (150.0, 130.0, 0.3))
        cells = dp.lethal_cells_within(cm, ring)
        ok("lethal small", np.array_equal(cells, capi.lethal_cells_within(m, ring)))
        full = dp.lethal_cells_within(cm, fx.to_rectangle(299, 259))
        ok("lethal full", np.array_equal(full, capi.lethal_cells_within(m, fx.to_rectangle(299, 259))))
        c, J = pg.get_cost([150.0, 130.0, 0.3], cells)
        rc, rJ = oc.get_cost([150.0, 130.0, 0.3], cells)
        ok("get_cost list", fx.tol_ok(c, rc) and fx.tol_ok(J, rJ).all())
        c, J = pg.get_cost([150.0, 130.0, 0.3], full)  # > 4096 cells: the grid-strided form
        rc, rJ = oc.get_cost([150.0, 130.0, 0.3], full)
        ok("get_cost long list", fx.tol_ok(c, rc) and fx.tol_ok(J, rJ).all())
        # dirty rows and a volatile map: k_pack_tiles on a part and on the whole plane
        m2 = m.copy()
        m2[40:90, 100:180] = fx.bernoulli_map(80, 50, 0.3, 7)
        cm.update(m2[40:90, 100:180], 100, 40)
        poses = fx.random_poses(6000, 300, 260, 9, margin=margin)
        ref2 = oc.get_cost_batch(m2, poses, nthreads=capi.max_threads())
        ok("batch after update", fx.tol_ok(pg.get_cost_batch(cm, poses), ref2).all())
        cm.set_volatile(True)
        ok("batch volatile", fx.tol_ok(pg.get_cost_batch(cm, poses), ref2).all())
        ok("lethal volatile", np.array_equal(dp.lethal_cells_within(cm, fx.to_rectangle(299, 259)), capi.lethal_cells_within(m2, fx.to_rectangle(299, 259))))
        cm.set_volatile(False)
        cm.update(m[40:90, 100:180], 100, 40)
        # the multi-start solve: k_problem_*, k_solve, k_check_footprint*, k_solve_select
        cfg = dict(lin_tolerance=8.0, rot_tolerance=0.8, weight_goal_lin=1, weight_goal_rot=1, weight_start_lin=1, weight_start_rot=1)
        pr = dp.problem(pg, cm)
        pr.init([140.0, 120.0, 0.2], [150.0, 130.0, 0.3], cfg)
        op = capi.Problem(oc, m)
        op.init([140.0, 120.0, 0.2], [150.0, 130.0, 0.3], **cfg)
        x0 = dp.sample_offsets(1, 8.0, 0.8, 256)
        ev = pr.eval_batch(x0)
        f_ref, g_ref, c_ref, _ = op.eval_batch(x0)
        ok("problem eval", fx.tol_ok(ev["f"], f_ref).all() and fx.tol_ok(ev["grad_f"], g_ref).all() and np.array_equal(ev["g"], c_ref))
        sol = pr.solve_batch(x0, trace=True)
        bad = 0
        for i in range(256):
            rec = sol["trace"][i, :sol["evals"][i]]
            f_ref, g_ref, _, _ = op.eval_batch(rec[:, :3])
            bad += int(not (fx.tol_ok(rec[:, 3], f_ref).all() and fx.tol_ok(rec[:, 4:7], g_ref).all()))
        ok("solve iterates", bad == 0)
    # check_footprint: CTA-per-pose (small batch) and thread-per-pose (large batch)
    sparse = fx.bernoulli_map(300, 260, 0.002, 13)
    for n in (300, 40000):
        poses = fx.random_poses(n, 300, 260, 17 + n, margin=-10.0)
        got = dp.check_footprint_batch(dp.Costmap(sparse), fp, poses)
        ok(f"{name} check_footprint {n}", np.array_equal(got, capi.check_footprint_batch(sparse, fp, poses, nthreads=capi.max_threads())))
print(f"exercise_all_kernels ok: {len(done)} checks, {dp.launch_count()} launches")
