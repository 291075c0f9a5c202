#!/usr/bin/env python
"""Side measurements of the small-batch paths (run under gpurun, one GPU), one JSON line each:

  solve     : dpose_problem_solve_batch, BASELINE config 4 (4096 starts, max_iter 20) — wall time per call
  batch     : device time of one batched get_cost launch for n = 1e3 .. 1e5 poses (cooperative kernel up to 18944
              poses, thread-per-pose kernel above), clean map (no plane rebuild) and volatile map
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
from oracle import capi, refapi  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def best_of(fn, reps):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def bench_solve():
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    m = fx.bernoulli_map(2000, 2000, 0.10, 1)
    cfg = dict(lin_tolerance=20.0, rot_tolerance=1.0, weight_goal_lin=1, weight_goal_rot=1, weight_start_lin=1,
               weight_start_rot=1)
    start, goal = [900.0, 950.0, 0.5], [1000.0, 1000.0, 0.0]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    pr = dp.problem(pg, dp.Costmap(m))
    pr.init(start, goal, cfg)
    for n in (1, 256, 4096, 16384, 65536):
        x0 = dp.sample_offsets(3, 20.0, 1.0, n)
        out = pr.solve_batch(x0)
        l0 = dp.launch_count()
        t = best_of(lambda: pr.solve_batch(x0), 10)
        launches = (dp.launch_count() - l0) // 10
        emit(what="solve_batch", starts=n, call_us=t * 1e6, evals_total=int(out["evals"].sum()),
             pose_evals_per_s=float(out["evals"].sum() / t), converged=int(((out["status"] & 3) == 1).sum()),
             launches_per_call=int(launches), mean_f=float(out["f"].mean()))
    # CPU arms on the same problem (4096 starts)
    x0 = dp.sample_offsets(3, 20.0, 1.0, 4096)
    cd = capi.CostData(fp, pad)
    scfg = capi.solver_cfg(cd.box, 20.0, 1.0)
    nt = capi.max_threads()
    if refapi.available():
        r = refapi.Problem(fp, pad, m)
        r.init(start, goal, **cfg)
        r.solve_batch(scfg, x0[:64], nthreads=nt)
        t = best_of(lambda: r.solve_batch(scfg, x0, nthreads=nt), 3)
        xr, fr, sr, er = r.solve_batch(scfg, x0, nthreads=nt)
        emit(what="solve_batch CPU arm: compiled plugin callbacks + oracle step rule", starts=4096, threads=nt,
             call_us=t * 1e6, pose_evals_per_s=float(er.sum() / t), mean_f=float(fr.mean()))


def bench_batch():
    import torch
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    m = fx.bernoulli_map(2000, 2000, 0.10, 1)
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    for volatile in (False, True):
        cm = dp.Costmap(m)
        cm.set_volatile(volatile)
        for n in (1000, 4096, 9000, 18000, 37000, 40000, 100000):
            poses = torch.from_numpy(fx.random_poses(n, 2000, 2000, 2)).cuda()
            out = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
            st = torch.cuda.current_stream()
            for _ in range(5):
                pg.get_cost_batch_device(cm, poses.data_ptr(), n, out.data_ptr(), stream=st.cuda_stream)
            torch.cuda.synchronize()
            reps = 50
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(reps):
                pg.get_cost_batch_device(cm, poses.data_ptr(), n, out.data_ptr(), stream=st.cuda_stream)
            b.record(st)
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / reps
            emit(what="batch launch (device time, back to back)", poses=n, volatile_map=volatile, us_per_call=us,
                 poses_per_s=n / us * 1e6)


def bench_shard():
    """one rank's share of BASELINE cfg5 under strong scaling on 8 GPUs: 1.25e6 poses on the 10000 x 10000 volatile
    map — device time per step and the host time it takes to queue a step"""
    import torch
    fp, pad = fx.FOOTPRINTS["rect_pad2"]
    side = 10000
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    d_map = torch.where(torch.rand((side, side), generator=g, device="cuda") < 0.10,
                        torch.tensor(254, dtype=torch.uint8, device="cuda"),
                        torch.tensor(0, dtype=torch.uint8, device="cuda")).contiguous()
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    cm = dp.Costmap(device_ptr=d_map.data_ptr(), size_x=side, size_y=side, pitch=side)
    for volatile in (True, False):
        cm.set_volatile(volatile)
        for n in (1_250_000, 2_500_000, 5_000_000, 10_000_000):
            u = torch.rand((n, 3), generator=g, device="cuda", dtype=torch.float64)
            poses = torch.stack([32 + u[:, 0] * (side - 64), 32 + u[:, 1] * (side - 64), -3.14159 + u[:, 2] * 6.28318], 1).contiguous()
            out = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
            st = torch.cuda.current_stream()
            for _ in range(5):
                pg.get_cost_batch_device(cm, poses.data_ptr(), n, out.data_ptr(), stream=st.cuda_stream)
            torch.cuda.synchronize()
            reps = 20
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record(st)
            for _ in range(reps):
                pg.get_cost_batch_device(cm, poses.data_ptr(), n, out.data_ptr(), stream=st.cuda_stream)
            b.record(st)
            issue = time.perf_counter() - t0
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / reps
            emit(what="one rank's shard of cfg5 (device time per step, back to back)", poses=n, volatile_map=volatile,
                 us_per_step=us, host_issue_us_per_step=issue / reps * 1e6, poses_per_s=n / us * 1e6)
            del poses, out, u


if __name__ == "__main__":
    which = sys.argv[1:] or ["solve", "batch"]
    for w in which:
        {"solve": bench_solve, "batch": bench_batch, "shard": bench_shard}[w]()
