#!/usr/bin/env python
"""Side measurements of the non-headline kernels of the path (run under gpurun, one GPU):

  precompute  : k_precompute device time per footprint vs the CPU pipelines (oracle C, cv2)
  lethal      : lethal_cells_within on the full 10000x10000 map and on a typical search box
  single      : BASELINE config 1 — one pose, explicit cell list (the reference's call shape)
  loop4096    : BASELINE config 4 call pattern — 4096 poses per call, 20 calls

Prints one JSON line per measurement; each is parity-checked against the oracle where cheap.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixtures as fx  # noqa: E402
import dpose_b200 as dp  # noqa: E402
from oracle import capi  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def best_of(fn, reps):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def bench_precompute():
    try:
        import cv2
        cv2.setNumThreads(1)
        cv2.ipp.setUseIPP(False)
        from oracle import cv_pipeline
    except Exception:
        cv2 = None
    for name in ("rect_pad2", "star40_pad2", "ship_pad10", "arrow_pad3"):
        fp, pad = fx.FOOTPRINTS[name]
        dp.cost_data(fp, pad)  # warm-up (module load)
        us = min(dp.cost_data(fp, pad).precompute_us() for _ in range(5))
        wall = best_of(lambda: dp.cost_data(fp, pad), 5)
        cd = dp.cost_data(fp, pad)
        oc = capi.CostData(fp, pad)
        exact = bool(np.array_equal(cd.get_data().view(np.uint32), oc.cost.view(np.uint32)) and
                     np.array_equal(cd.stages()[0], oc.edt_sq))
        cpu_oracle = best_of(lambda: capi.CostData(fp, pad), 5)
        cpu_cv2 = best_of(lambda: cv_pipeline.cv_cost_data(fp, pad), 5) if cv2 is not None else None
        emit(what="precompute", footprint=name, table=[cd.rows, cd.cols], gpu_kernel_us=us,
             gpu_call_wall_us=wall * 1e6, cpu_oracle_us=cpu_oracle * 1e6,
             cpu_cv2_us=None if cpu_cv2 is None else cpu_cv2 * 1e6, bit_exact_vs_oracle=exact)


def bench_lethal():
    import torch
    side = 10000
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    d_map = torch.where(torch.rand((side, side), generator=g, device="cuda") < 0.10,
                        torch.tensor(254, dtype=torch.uint8, device="cuda"),
                        torch.tensor(0, dtype=torch.uint8, device="cuda")).contiguous()
    cm = dp.Costmap(device_ptr=d_map.data_ptr(), size_x=side, size_y=side, pitch=side)
    for label, ring in (("full 10000x10000", fx.to_rectangle(side - 1, side - 1)),
                        ("search box 101x101", fx.to_rectangle((4950, 4950), (5050, 5050))),
                        ("rotated 2000x1200 box", fx.rotated_ring(fx.to_rectangle((-1000, -600), (1000, 600)),
                                                                  (5000.0, 5000.0, 0.5)))):
        cells = dp.lethal_cells_within(cm, ring, on_device=True)
        n = len(cells)
        del cells
        torch.cuda.synchronize()

        def run():
            c = dp.lethal_cells_within(cm, ring, on_device=True)
            del c
        wall = best_of(run, 5)
        dev_us = cm.lethal_us()
        xs, ys = ring[:, 0], ring[:, 1]
        box_bytes = int((xs.max() - xs.min() + 1) * (ys.max() - ys.min() + 1))
        algo = box_bytes + 8 * n  # the box's map bytes once + 8 B per lethal cell written
        for volatile in (False, True):  # clean plane (cached 1-bit view) / stale plane (the uint8 map is re-read)
            cm.set_volatile(volatile)
            run()
            wall = best_of(run, 5)
            dev_us = cm.lethal_us()
            emit(what="lethal_cells_within", box=label, cells=n, plane="rebuilt from the uint8 map in the call" if volatile
                 else "clean (cached)", call_wall_us=wall * 1e6, device_us=dev_us, algorithmic_bytes=algo,
                 achieved_gbs_device=algo / (dev_us * 1e-6) / 1e9, frac_of_hbm_peak=algo / (dev_us * 1e-6) / 1e9 / 6527.5)
        cm.set_volatile(False)
    # the reference's own perf workload (dpose_core/perf/dpose_core.cpp:73-90): 200 x 200, every 2nd cell lethal
    m = np.zeros((200, 200), np.uint8)
    m.reshape(-1)[::2] = fx.LETHAL
    cm2 = dp.Costmap(m)
    ring = fx.to_rectangle(198, 198)
    ok = bool(np.array_equal(dp.lethal_cells_within(cm2, ring), capi.lethal_cells_within(m, ring)))
    t_gpu = best_of(lambda: dp.lethal_cells_within(cm2, ring, on_device=True), 20)
    t_gpu_host = best_of(lambda: dp.lethal_cells_within(cm2, ring), 20)
    t_cpu = best_of(lambda: capi.lethal_cells_within(m, ring), 20)
    emit(what="perf_dpose_costmap literal (200x200, every 2nd cell lethal, box 198x198)", cells=int((m == fx.LETHAL).sum()),
         gpu_call_us_result_on_device=t_gpu * 1e6, gpu_call_us_result_on_host=t_gpu_host * 1e6, cpu_oracle_call_us=t_cpu * 1e6,
         parity=ok)
    # parity on a host map small enough for the oracle
    m = fx.bernoulli_map(2000, 2000, 0.10, 1)
    ring = fx.rotated_ring(fx.to_rectangle((-300, -200), (300, 200)), (1000.0, 1000.0, 0.7))
    ok = bool(np.array_equal(dp.lethal_cells_within(dp.Costmap(m), ring), capi.lethal_cells_within(m, ring)))
    emit(what="lethal_cells_within", parity_vs_oracle_2000x2000=ok)


def bench_single():
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    oc = capi.CostData(fp, pad)
    m = fx.bernoulli_map(200, 200, 0.10, 0)
    cm = dp.Costmap(m)
    pose = [100.0, 100.0, 0.3]
    ring = fx.rotated_ring(pg.get_data().get_box(), pose)
    cells = dp.lethal_cells_within(cm, ring)
    dcells = dp.cell_vector_device.upload(cells)
    c, J = pg.get_cost(pose, cells)
    rc, rJ = oc.get_cost(pose, cells)
    ok = bool(fx.tol_ok(c, rc) and fx.tol_ok(J, rJ).all())
    reps = 2000
    t_host = best_of(lambda: [pg.get_cost(pose, cells) for _ in range(reps)], 3) / reps
    t_dev = best_of(lambda: [pg.get_cost(pose, dcells) for _ in range(reps)], 3) / reps
    t_cpu = best_of(lambda: [oc.get_cost(pose, cells) for _ in range(reps)], 3) / reps
    emit(what="single_pose_get_cost (BASELINE cfg1)", cells=int(len(cells)), gpu_call_us_host_cells=t_host * 1e6,
         gpu_call_us_device_cells=t_dev * 1e6, cpu_oracle_call_us=t_cpu * 1e6, parity=ok,
         note="one pose per call is launch-latency bound on a GPU; the batched entry point exists for this reason")
    # the literal perf bench of the reference (dpose_core/perf/dpose_core.cpp:17-57)
    fp, pad = fx.FOOTPRINTS["arrow_pad3"]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    oc = capi.CostData(fp, pad)
    cells = np.array([[x, y] for x in range(10) for y in range(10)], np.int32)
    t_host = best_of(lambda: [pg.get_cost([0, 0, 0], cells) for _ in range(reps)], 3) / reps
    t_cpu = best_of(lambda: [oc.get_cost([0, 0, 0], cells) for _ in range(reps)], 3) / reps
    emit(what="perf_dpose_core literal (arrow pad 3, 100 cells)", gpu_call_us=t_host * 1e6, cpu_oracle_call_us=t_cpu * 1e6)


def bench_loop4096():
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    m = fx.bernoulli_map(2000, 2000, 0.10, 1)
    cm = dp.Costmap(m)
    rng = np.random.default_rng(3)
    r, a = rng.uniform(0, 20, 4096), rng.uniform(-np.pi, np.pi, 4096)
    poses = np.stack([1000 + r * np.cos(a), 1000 + r * np.sin(a), rng.uniform(-1, 1, 4096)], axis=1)
    out = np.empty((4096, 4))
    pg.get_cost_batch(cm, poses, out=out)
    t = best_of(lambda: [pg.get_cost_batch(cm, poses, out=out) for _ in range(20)], 5) / 20
    oc = capi.CostData(fp, pad)
    ref = oc.get_cost_batch(m, poses, nthreads=capi.max_threads())
    t_cpu = best_of(lambda: oc.get_cost_batch(m, poses, nthreads=capi.max_threads()), 3)
    emit(what="4096-pose call (BASELINE cfg4 call pattern, host buffers)", gpu_call_us=t * 1e6,
         gpu_poses_per_s=4096 / t, cpu_oracle_call_us=t_cpu * 1e6, cpu_threads=capi.max_threads(),
         parity=bool(fx.tol_ok(out, ref).all()))


def bench_cfg4():
    """BASELINE config 4: the goal-tolerance loop with 4096 multi-start displacements, one batched NLP evaluation
    per iteration.  Ipopt is absent, so the loop is emulated by 20 projected-gradient iterations (the step rule
    is ours; the evaluated functions f, grad f, g, jac g are the reference's)."""
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    m = fx.bernoulli_map(2000, 2000, 0.10, 1)
    cfg = dict(lin_tolerance=20.0, rot_tolerance=1.0, weight_goal_lin=1, weight_goal_rot=1, weight_start_lin=1,
               weight_start_rot=1)
    start, goal = [900.0, 950.0, 0.5], [1000.0, 1000.0, 0.0]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    pr = dp.problem(pg, dp.Costmap(m))
    oc = capi.Problem(capi.CostData(fp, pad), m)
    t_init = best_of(lambda: pr.init(start, goal, cfg), 5)
    t_init_cpu = best_of(lambda: oc.init(start, goal, **cfg), 5)
    rng = np.random.default_rng(3)
    n, iters = 4096, 20
    r, a = rng.uniform(0, cfg["lin_tolerance"], n), rng.uniform(-np.pi, np.pi, n)
    x0 = np.stack([np.cos(a) * r, np.sin(a) * r, rng.uniform(0, cfg["rot_tolerance"], n)], axis=1)

    def project(x):
        rad = np.hypot(x[:, 0], x[:, 1])
        s = np.minimum(1.0, cfg["lin_tolerance"] / np.maximum(rad, 1e-12))
        x[:, 0] *= s
        x[:, 1] *= s
        x[:, 2] = np.clip(x[:, 2], -cfg["rot_tolerance"], cfg["rot_tolerance"])
        return x

    def loop(evaluate):
        x = x0.copy()
        f_first = f_last = None
        t_eval = 0.0
        for _ in range(iters):
            t0 = time.perf_counter()
            f, grad = evaluate(x)
            t_eval += time.perf_counter() - t0
            f_first = f.mean() if f_first is None else f_first
            f_last = f.mean()
            x = project(x - 0.05 * grad / np.maximum(np.abs(grad).max(axis=1, keepdims=True), 1e-9))
        return t_eval, f_first, f_last, x

    def gpu_eval(x):
        o = pr.eval_batch(x, want=("f", "grad_f"))
        return o["f"], o["grad_f"]

    def cpu_eval(x):
        f, grad, _, _ = oc.eval_batch(x, nthreads=capi.max_threads())
        return f, grad
    loop(gpu_eval)
    tg, f0, f1, xg = loop(gpu_eval)
    tc, _, f1c, xc = loop(cpu_eval)
    emit(what="cfg4: 4096 multi-start x 20 iterations, batched NLP evaluation per iteration",
         gpu_eval_us_per_iteration=tg / iters * 1e6, gpu_iterations_per_s=iters / tg,
         cpu_oracle_eval_us_per_iteration=tc / iters * 1e6, cpu_threads=capi.max_threads(),
         problem_init_us_gpu=t_init * 1e6, problem_init_us_cpu=t_init_cpu * 1e6,
         mean_cost_first_iteration=float(f0), mean_cost_last_iteration=float(f1),
         final_iterate_max_abs_diff_gpu_vs_cpu=float(np.abs(xg - xc).max()),
         note="Ipopt absent: projected-gradient loop emulates the call pattern; evaluated functions are the reference's")


if __name__ == "__main__":
    which = sys.argv[1:] or ["precompute", "lethal", "single", "loop4096", "cfg4"]
    for w in which:
        {"precompute": bench_precompute, "lethal": bench_lethal, "single": bench_single,
         "loop4096": bench_loop4096, "cfg4": bench_cfg4}[w]()
