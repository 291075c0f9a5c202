#!/usr/bin/env python
"""parity soak of the caller-side rows: random footprints, maps, goals, tolerances and weights through
dpose_problem_eval_batch, dpose_problem_solve_batch and dpose_check_footprint_batch, against the oracle (itself pinned to the
compiled plugin, tests/test_reference_build.py): f / grad f within tolerance, g and the search box bit-exact, every solver
step replayed bit for bit by the restated rule, every final pose's flag equal to check_footprint, the ranking recomputed.
usage (under gpurun): python tools/fuzz_solver.py [configs] [seed]"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import fixtures as fx  # noqa: E402
import dpose_b200 as dp  # noqa: E402
from oracle import capi  # noqa: E402

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 11)
nt = capi.max_threads()
fails, t_start = [], time.time()


def check(tag, cond, rec):
    if not cond:
        fails.append((tag, dict(rec)))
    return bool(cond)


done = 0
for k in range(n_cfg):
    radius = float(rng.choice([3, 5, 8, 12, 18, 26]))
    nv = int(rng.integers(3, 10))
    ang = np.sort(rng.uniform(0, 2 * math.pi, nv))
    rad = rng.uniform(0.4, 1.0, nv) * radius
    fp = np.stack([np.round(rad * np.cos(ang)), np.round(rad * np.sin(ang))], 1).astype(np.int32)
    if len(np.unique(fp[:, 0])) < 2 or len(np.unique(fp[:, 1])) < 2:
        continue
    pad = int(rng.integers(1, 5))
    sx, sy = int(rng.integers(120, 900)), int(rng.integers(120, 700))
    dens = float(rng.choice([0.0, 0.002, 0.01, 0.05, 0.15]))
    m = fx.bernoulli_map(sx, sy, dens, int(rng.integers(1 << 30)))
    lin, rot = float(rng.choice([1.0, 5.0, 16.67, 20.0, 40.5])), float(rng.choice([0.1, 0.7, 1.0, 3.0]))
    cfg = dict(lin_tolerance=lin, rot_tolerance=rot, weight_goal_lin=float(rng.choice([0, 1, 3.5])), weight_goal_rot=float(rng.choice([0, 1])),
               weight_start_lin=float(rng.choice([0, 1, 0.25])), weight_start_rot=float(rng.choice([0, 1])))
    goal = [float(rng.uniform(0, sx - 1)), float(rng.uniform(0, sy - 1)), float(rng.uniform(-3.1, 3.1))]  # near the border too: clamped box
    start = [float(rng.uniform(0, sx - 1)), float(rng.uniform(0, sy - 1)), float(rng.uniform(-3.1, 3.1))]
    n = int(rng.choice([1, 33, 500, 4096]))
    rec = {"cfg": k, "vertices": nv, "radius": radius, "pad": pad, "map": [sx, sy], "density": dens, "lin": lin, "rot": rot, "starts": n}
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    cm = dp.Costmap(m)
    pr = dp.problem(pg, cm)
    pr.init(start, goal, cfg)
    oc = capi.CostData(fp, pad)
    op = capi.Problem(oc, m)
    op.init(start, goal, **cfg)
    ok = check("search box", np.array_equal(pr.search_box, op.search_box), rec)
    x0 = dp.sample_offsets(int(rng.integers(1 << 30)), lin, rot, n)
    ev = pr.eval_batch(x0)
    f_ref, g_ref, c_ref, j_ref = op.eval_batch(x0, nthreads=nt)
    ok &= check("eval f", fx.tol_ok(ev["f"], f_ref).all(), rec)
    ok &= check("eval grad", fx.tol_ok(ev["grad_f"], g_ref, rel=1e-5, abs_=2e-6).all(), rec)
    ok &= check("eval g / jac g", np.array_equal(ev["g"], c_ref) and np.array_equal(ev["jac_g"], j_ref), rec)
    sol = pr.solve_batch(x0, trace=True)
    x, f, st, evals, tr = sol["x"], sol["f"], sol["status"], sol["evals"], sol["trace"]
    scfg = capi.solver_cfg(oc.box, lin, rot)
    idx = np.arange(n) if n <= 300 else rng.choice(n, 300, replace=False)
    recs = np.concatenate([tr[i, :evals[i]] for i in idx])
    fr, gr, cg, _ = op.eval_batch(recs[:, :3], nthreads=nt)
    ok &= check("iterates f", fx.tol_ok(recs[:, 3], fr).all(), rec)
    ok &= check("iterates grad", fx.tol_ok(recs[:, 4:7], gr, rel=1e-5, abs_=2e-6).all(), rec)
    ok &= check("iterates g", np.array_equal(recs[:, 7:9], cg), rec)
    ok &= check("feasible", np.all(recs[:, 7] <= lin ** 2 * (1 + 1e-12)) and np.all(recs[:, 8] <= rot ** 2 * (1 + 1e-12)), rec)
    steps_ok = True
    for i in idx:
        r_ = tr[i, :evals[i]]
        trials, statuses, s = capi.solver_replay(scfg, x0[i], zip(r_[:, 3], r_[:, 4:7]))
        steps_ok &= bool(np.array_equal(trials, r_[:, :3]) and np.array_equal(np.array(s.x[:]), x[i]) and s.f == f[i]
                         and s.status == (st[i] & 3) and s.evals == evals[i])
    ok &= check("step rule replay", steps_ok, rec)
    poses = x + np.asarray(goal)
    free = capi.check_footprint_batch(m, fp, poses, nthreads=nt).astype(bool)
    ok &= check("footprint flags", np.array_equal(free, (st & dp.SOLVE_FOOTPRINT_FREE) != 0), rec)
    acc = ((st & 3) == 1) & free
    sm = sol["summary"]
    rank_ok = np.array_equal(acc, (st & dp.SOLVE_ACCEPTED) != 0) and sm["n_accepted"] == int(acc.sum())
    if acc.any():
        rank_ok &= sm["first"] == int(np.nonzero(acc)[0][0]) and sm["best"] == int(np.nonzero(acc)[0][np.argmin(f[acc])])
    else:
        rank_ok &= sm["first"] == -1 and sm["best"] == -1
    ok &= check("ranking", bool(rank_ok), rec)
    # check_footprint on poses anywhere (off the map included), both kernels (CTA per pose / thread per pose)
    for nb in (int(rng.integers(1, 400)), 40000):
        pb = np.stack([rng.uniform(-20, sx + 20, nb), rng.uniform(-20, sy + 20, nb), rng.uniform(-7, 7, nb)], 1)
        ok &= check("check_footprint_batch", np.array_equal(dp.check_footprint_batch(cm, fp, pb),
                                                            capi.check_footprint_batch(m, fp, pb, nthreads=nt).astype(bool)), rec)
    rec.update(ok=bool(ok), accepted=int(acc.sum()), converged=int(((st & 3) == 1).sum()))
    done += 1
    print(json.dumps(rec), flush=True)
print(json.dumps({"configs": done, "failures": len(fails), "first_failures": [(t, str(d)[:300]) for t, d in fails[:5]],
                  "seconds": round(time.time() - t_start, 1)}))
sys.exit(1 if fails else 0)
