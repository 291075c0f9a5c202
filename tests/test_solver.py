"""The multi-start goal-tolerance solve: dpose_problem_solve_batch against

  * the reference's own NLP callbacks (dpose_goal_tolerance.cpp:90-292, compiled into oracle/_ref) at EVERY iterate the
    device visited (trace): f / grad f within the north-star tolerance, g bit for bit;
  * the step rule restated in oracle/dpose_oracle.c (oracle_solver_*), replayed on the device's own (f, grad f): the
    next trial point of every evaluation bit for bit;
  * the reference's check_footprint (dpose_costmap.cpp:460-582) on every final pose, and the ranking recomputed in numpy;
  * the sampler of starting offsets against the plugin's own uniform_pose_distribution (:333-357).

Ipopt is absent from this image: the step rule is the product's (solve.cu header); what is the reference's is what the
rule evaluates, the feasible set, the validation and the sampler.
"""
import numpy as np
import pytest

import fixtures as fx
from oracle import capi, refapi

CFG = dict(lin_tolerance=20.0, rot_tolerance=1.0, weight_goal_lin=1, weight_goal_rot=1, weight_start_lin=1,
           weight_start_rot=1)


def sparse_map(seed=7, side=600, density=0.004):
    """a map on which some footprint placements are free and some collide"""
    return fx.bernoulli_map(side, side, density, seed)


# ---- CPU: sampler, step rule, CPU arms ------------------------------------------------------------------------------
@pytest.mark.parametrize("seed,lin,rot", [(3, 20.0, 1.0), (0, 1.0, 3.14), (12345, 16.67, 0.25)])
def test_sampler_matches_the_plugins_distribution(seed, lin, rot):
    want = refapi.sample_offsets(seed, lin, rot, 5000)  # uniform_pose_distribution of the compiled plugin source
    got = capi.sample_offsets(seed, lin, rot, 5000, first_zero=False)
    assert np.array_equal(got, want)
    assert np.all(np.hypot(got[:, 0], got[:, 1]) <= lin * (1 + 1e-15)) and np.all((got[:, 2] >= 0) & (got[:, 2] < rot))
    z = capi.sample_offsets(seed, lin, rot, 10, first_zero=True)  # attempt 0 is the goal itself (:431)
    assert np.array_equal(z[0], [0, 0, 0]) and np.array_equal(z[1:], want[:9])


def test_product_sampler_is_host_arithmetic_and_matches():
    import dpose_b200 as dp
    want = refapi.sample_offsets(3, 20.0, 1.0, 4096)
    assert np.array_equal(dp.sample_offsets(3, 20.0, 1.0, 4096, first_zero=False), want)
    assert np.array_equal(dp.sample_offsets(3, 20.0, 1.0, 4, first_zero=True)[1:], want[:3])
    with pytest.raises(dp.DposeError):
        dp.sample_offsets(3, -1.0, 1.0, 4)


def _problems(name="rect_pad2", m=None, goal=(300.0, 300.0, 0.3), cfg=CFG):
    fp, pad = fx.FOOTPRINTS[name]
    m = sparse_map() if m is None else m
    cd = capi.CostData(fp, pad)
    start = [goal[0] - 40.0, goal[1] + 25.0, goal[2] + 2.5]
    o = capi.Problem(cd, m)
    o.init(start, list(goal), **cfg)
    r = refapi.Problem(fp, pad, m)
    r.init(start, list(goal), **cfg)
    return fp, pad, m, cd, start, o, r


def test_step_rule_on_reference_callbacks():
    """the restated step rule driven by the compiled plugin's TNLP callbacks: feasible iterates, monotone accepted
    costs, every traced f / grad f equal to a fresh evaluation, deterministic, and close to the oracle-driven run"""
    fp, pad, m, cd, start, o, r = _problems(m=fx.bernoulli_map(600, 600, 0.05, 3))
    cfg = capi.solver_cfg(cd.box, CFG["lin_tolerance"], CFG["rot_tolerance"])
    x0 = capi.sample_offsets(5, CFG["lin_tolerance"], CFG["rot_tolerance"], 600)
    x, f, st, ev, tr = r.solve_batch(cfg, x0, nthreads=2, trace=True)
    x2, f2, st2, ev2 = r.solve_batch(cfg, x0, nthreads=1)
    assert np.array_equal(x, x2) and np.array_equal(f, f2) and np.array_equal(st, st2) and np.array_equal(ev, ev2)
    f0 = r.eval_batch(x0)[0]
    assert np.all(f <= f0) and f.mean() < 0.9 * f0.mean()
    assert set(np.unique(st & 3).tolist()) <= {1, 2, 3} and (st & 3 == 1).mean() > 0.5
    assert np.all((ev >= 1) & (ev <= cfg.max_iter + 1))
    for i in range(0, 600, 7):
        k = ev[i]
        rec = tr[i, :k]
        fr, gr, cg, _ = r.eval_batch(rec[:, :3])
        assert np.array_equal(rec[:, 3], fr) and np.array_equal(rec[:, 4:7], gr) and np.array_equal(rec[:, 7:9], cg)
        assert np.all(rec[:, 7] <= CFG["lin_tolerance"] ** 2 * (1 + 1e-12)) and np.all(rec[:, 8] <= CFG["rot_tolerance"] ** 2)
        assert rec[-1, 10] == (st[i] & 3) and np.all(rec[:-1, 10] == 0)
        trials, statuses, s = capi.solver_replay(cfg, x0[i], zip(rec[:, 3], rec[:, 4:7]))
        assert np.array_equal(trials, rec[:, :3]) and np.array_equal(np.array(s.x[:]), x[i]) and s.f == f[i]
    # the oracle's own evaluation differs from the compiled reference only by the summation order of the cells
    # (unordered_map rows); through the reference's finite-difference Jacobian (h = 1e-6, x 5e5) that becomes ~1e-7 of
    # gradient noise, which flips a line-search decision now and then: most starts end at the same point, all at a
    # comparable cost
    xo, fo, sto, evo = o.solve_batch(cfg, x0, nthreads=2)
    assert (np.abs(xo - x).max(axis=1) < 1e-6).mean() > 0.75
    assert abs(fo.mean() - f.mean()) < 0.02 * f.mean()


def test_step_rule_edge_cases():
    fp, pad, m, cd, start, o, r = _problems()
    # zero tolerances: the feasible set is the goal itself
    cfg0 = capi.solver_cfg(cd.box, 0.0, 0.0)
    x, f, st, ev = o.solve_batch(cfg0, [[3.0, -2.0, 0.5], [0, 0, 0]])
    assert np.array_equal(x, np.zeros((2, 3))) and np.all(ev == 1) and np.all(np.isin(st & 3, [1, 3]))
    # no budget beyond the first evaluation
    cfg1 = capi.solver_cfg(cd.box, 20.0, 1.0, max_iter=0)
    x0 = capi.sample_offsets(1, 20.0, 1.0, 50)
    x, f, st, ev = o.solve_batch(cfg1, x0)
    assert np.array_equal(x, x0) and np.all(ev == 1) and np.all(np.isin(st & 3, [1, 2]))
    # a start outside the tolerance is projected into it first
    x, f, st, ev, tr = o.solve_batch(capi.solver_cfg(cd.box, 20.0, 1.0), [[100.0, 0.0, 5.0]], trace=True)
    assert np.allclose(tr[0, 0, :3], [20.0, 0.0, 1.0])


# ---- GPU ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dp():
    import dpose_b200 as dp
    if dp.device_count() == 0:
        pytest.skip("no CUDA device")
    return dp


def _gpu_problem(dp, fp, pad, m, start, goal, cfg):
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    pr = dp.problem(pg, dp.Costmap(m))
    pr.init(start, list(goal), cfg)
    return pr


CASES = [
    ("rect_pad2", dict(), (300.0, 300.0, 0.3), 0.05, 4096),       # dense obstacles: long descents
    ("rect_pad2", dict(), (300.25, 299.5, -2.0), 0.004, 2000),    # sparse: some placements are free
    ("star40_pad2", dict(lin_tolerance=10.0, rot_tolerance=0.6), (250.5, 250.5, 1.0), 0.01, 700),  # 35 x 55 table, G = 32
    ("lshape_pad2", dict(lin_tolerance=5.0, weight_goal_lin=0, weight_goal_rot=0), (10.0, 12.0, 3.0), 0.02, 300),  # clamped box
    ("rect_pad2", dict(lin_tolerance=16.67, rot_tolerance=0.7), (200.0, 150.0, 0.4), 0.03, 40000),  # G = 2 lanes per start
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,cfg_over,goal,density,n", CASES)
def test_solve_batch_iterates_steps_and_validation(dp, name, cfg_over, goal, density, n):
    cfg = dict(CFG)
    cfg.update(cfg_over)
    fp, pad, m, cd, start, o, r = _problems(name, fx.bernoulli_map(600, 600, density, 11), goal, cfg)
    pr = _gpu_problem(dp, fp, pad, m, start, goal, cfg)
    x0 = capi.sample_offsets(3, cfg["lin_tolerance"], cfg["rot_tolerance"], n)
    out = pr.solve_batch(x0, trace=True)
    x, f, st, ev, tr = out["x"], out["f"], out["status"], out["evals"], out["trace"]
    scfg = capi.solver_cfg(cd.box, cfg["lin_tolerance"], cfg["rot_tolerance"])
    assert np.all((ev >= 1) & (ev <= scfg.max_iter + 1)) and set(np.unique(st & 3).tolist()) <= {1, 2, 3}
    # (a) every iterate the device evaluated, against the compiled plugin's callbacks at the same displacement
    idx = np.arange(n) if n <= 1000 else np.random.default_rng(0).choice(n, 1000, replace=False)
    recs = np.concatenate([tr[i, :ev[i]] for i in idx])
    fr, gr, cg, _ = r.eval_batch(recs[:, :3])
    assert fx.tol_ok(recs[:, 3], fr).all() and fx.tol_ok(recs[:, 4:7], gr).all()
    assert np.array_equal(recs[:, 7:9], cg)  # g0 = x^2 + y^2, g1 = theta^2: the same two products and one sum
    assert np.all(recs[:, 7] <= cfg["lin_tolerance"] ** 2 * (1 + 1e-12)) and np.all(recs[:, 8] <= cfg["rot_tolerance"] ** 2)
    # (b) the step rule, replayed on the device's own (f, grad f): bit for bit
    for i in idx:
        rec = tr[i, :ev[i]]
        trials, statuses, s = capi.solver_replay(scfg, x0[i], zip(rec[:, 3], rec[:, 4:7]))
        assert np.array_equal(trials, rec[:, :3]), i
        assert np.array_equal(statuses, rec[:, 10].astype(np.int64)), i
        assert np.array_equal(np.array(s.x[:]), x[i]) and s.f == f[i] and s.status == (st[i] & 3) and s.evals == ev[i]
    # (c) validation of every final pose with the reference's check_footprint, flags and ranking
    poses = x + np.asarray(goal)
    c, s_ = np.cos(poses[:, 2]), np.sin(poses[:, 2])
    fpf = np.asarray(fp, np.float64)
    free = np.zeros(n, bool)
    for i in idx:
        moved = np.stack([fx.round_half_away((c[i] * fpf[:, 0] + -s_[i] * fpf[:, 1]) + poses[i, 0]),
                          fx.round_half_away((s_[i] * fpf[:, 0] + c[i] * fpf[:, 1]) + poses[i, 1])], 1)
        free[i] = refapi.check_footprint(m, moved)
    assert np.array_equal(free[idx], (st[idx] & dp.SOLVE_FOOTPRINT_FREE) != 0)
    if n > 1000:  # the rest against the oracle's batch (itself pinned to the reference's check_footprint)
        free = capi.check_footprint_batch(m, fp, poses, nthreads=capi.max_threads()).astype(bool)
        assert np.array_equal(free, (st & dp.SOLVE_FOOTPRINT_FREE) != 0)
    acc = ((st & 3) == 1) & free
    assert np.array_equal(acc, (st & dp.SOLVE_ACCEPTED) != 0)
    sm = out["summary"]
    assert sm["n_converged"] == int(((st & 3) == 1).sum()) and sm["n_accepted"] == int(acc.sum())
    if acc.any():
        first = int(np.nonzero(acc)[0][0])
        best = int(np.nonzero(acc)[0][np.argmin(f[acc])])
        assert sm["first"] == first and sm["best"] == best and sm["best_f"] == f[best]
        assert np.array_equal(sm["best_pose"], poses[best]) and np.array_equal(sm["first_pose"], poses[first])
    else:
        assert sm["first"] == -1 and sm["best"] == -1
    # (d) against the CPU arm (same rule on the plugin's callbacks): the finite-difference Jacobian's noise flips a
    # decision now and then, most starts agree and the costs are statistically the same
    xr, frr, str_, evr = r.solve_batch(scfg, x0[idx], nthreads=capi.max_threads())
    assert (np.abs(xr - x[idx]).max(axis=1) < 1e-5).mean() > 0.6
    assert abs(frr.mean() - f[idx].mean()) <= 0.03 * abs(frr.mean()) + 1e-9
    assert np.all(f[idx] <= recs[np.cumsum(np.r_[0, ev[idx][:-1]]), 3] + 1e-12)  # never worse than the start


@pytest.mark.gpu
def test_solve_batch_acceptance_happens_and_is_the_references(dp):
    """a sparse map where the goal itself collides but free placements exist inside the tolerance: some starts must be
    accepted, and the accepted set is exactly converged-and-free by the reference's check_footprint"""
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    m = np.zeros((400, 400), np.uint8)
    m[195:206, 195:206] = 254  # a block on the goal
    goal, start = (200.0, 200.0, 0.0), [150.0, 150.0, 0.0]
    pr = _gpu_problem(dp, fp, pad, m, start, goal, CFG)
    x0 = dp.sample_offsets(9, CFG["lin_tolerance"], CFG["rot_tolerance"], 512)
    out = pr.solve_batch(x0)
    st = out["status"]
    assert not capi.check_footprint_batch(m, fp, np.array([goal]), nthreads=1)[0]  # attempt 0 starts in collision
    assert out["summary"]["n_accepted"] > 0 and out["summary"]["first"] >= 0
    poses = out["x"] + np.asarray(goal)
    free = capi.check_footprint_batch(m, fp, poses, nthreads=2).astype(bool)
    assert np.array_equal(free, (st & dp.SOLVE_FOOTPRINT_FREE) != 0)
    assert np.array_equal(((st & 3) == 1) & free, (st & dp.SOLVE_ACCEPTED) != 0)
    b = out["summary"]["best"]
    assert st[b] & dp.SOLVE_ACCEPTED and out["f"][b] == out["f"][(st & dp.SOLVE_ACCEPTED) != 0].min()


@pytest.mark.gpu
def test_solve_batch_edge_cases(dp):
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    m = sparse_map()
    goal, start = (300.0, 300.0, 0.3), [260.0, 325.0, 2.8]
    pr = _gpu_problem(dp, fp, pad, m, start, goal, CFG)
    # no starts
    out = pr.solve_batch(np.zeros((0, 3)))
    assert out["x"].shape == (0, 3) and out["summary"]["first"] == -1 and out["summary"]["n_accepted"] == 0
    # one start; no budget beyond the first evaluation
    out = pr.solve_batch([[1.0, 2.0, 0.1]], max_iter=0)
    assert np.array_equal(out["x"], [[1.0, 2.0, 0.1]]) and out["evals"][0] == 1 and (out["status"][0] & 3) in (1, 2)
    f_ref = pr.eval_batch([[1.0, 2.0, 0.1]], want=("f",))["f"][0]
    assert abs(out["f"][0] - f_ref) <= 1e-9 * (1 + abs(f_ref))
    # a start outside the tolerance is projected first
    out = pr.solve_batch([[100.0, 0.0, 5.0]], trace=True)
    assert np.allclose(out["trace"][0, 0, :3], [20.0, 0.0, 1.0])
    # zero tolerances: the feasible set is the goal itself
    pr0 = _gpu_problem(dp, fp, pad, m, start, goal, dict(CFG, lin_tolerance=0.0, rot_tolerance=0.0))
    out = pr0.solve_batch([[3.0, -2.0, 0.5], [0, 0, 0]])
    assert np.array_equal(out["x"], np.zeros((2, 3))) and np.all(out["evals"] == 1)
    # an empty costmap: only the regularisers are left, every start converges to a free pose... of a map with no cells:
    # is_inside fails for every vertex (dpose_costmap.cpp:186-199), so nothing is accepted
    pre = _gpu_problem(dp, fp, pad, np.zeros((0, 0), np.uint8), start, goal, CFG)
    out = pre.solve_batch(dp.sample_offsets(1, 20.0, 1.0, 64))
    assert np.all((out["status"] & 3) == 1) and out["summary"]["n_accepted"] == 0
    # bad options
    with pytest.raises(dp.DposeError):
        pr.solve_batch([[0, 0, 0]], tol=-1.0)
    # repeated calls on one problem give the same bits (no state leaks between solves)
    x0 = dp.sample_offsets(2, 20.0, 1.0, 300)
    a, b = pr.solve_batch(x0), pr.solve_batch(x0)
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["f"], b["f"]) and np.array_equal(a["status"], b["status"])
