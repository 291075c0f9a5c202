#!/usr/bin/env python
"""soak of the costmap's state machine: random sequences of dpose_map_update windows, rewrites of the
library's device copy behind its back (volatile maps only — that is the contract), volatile on / off, and — after every
mutation — a random reader: the host-buffer batch call, the device call, the row-band device call, a lethal extraction over a
random box, a problem init + eval.  Every reader is compared with the oracle on a host mirror of the map: the lazily rebuilt
lethal plane (dirty rows, volatile bands) must never serve a stale row.
usage (under gpurun): python tools/fuzz_map_state.py [sequences] [seed]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import fixtures as fx  # noqa: E402
import dpose_b200 as dp  # noqa: E402
from dpose_b200.dist import row_band  # noqa: E402
from oracle import capi  # noqa: E402

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
nt = capi.max_threads()
fp, pad = fx.RECT, 2
pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
oc = capi.CostData(fp, pad)
reach = pg.reach()
fails, ops_done, t_start = [], 0, time.time()


def device_view(cm, sy):
    ptr, pitch = cm.device_ptr()

    class Raw:
        __cuda_array_interface__ = {"shape": (sy, pitch), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}
    return torch.as_tensor(Raw(), device="cuda")


for s in range(n_seq):
    sx, sy = int(rng.integers(100, 900)), int(rng.integers(100, 700))
    hm = fx.bernoulli_map(sx, sy, float(rng.choice([0.0, 0.05, 0.2])), int(rng.integers(1 << 30)))
    cm = dp.Costmap(hm)
    view = device_view(cm, sy)
    volatile = False
    log = []
    for step in range(int(rng.integers(6, 16))):
        # ---- a mutation ------------------------------------------------------------------------------------------------
        op = rng.choice(["update", "update", "rewrite", "volatile", "none"])
        if op == "update":
            w, h = int(rng.integers(1, min(sx, 300) + 1)), int(rng.integers(1, min(sy, 200) + 1))
            x0, y0 = int(rng.integers(0, sx - w + 1)), int(rng.integers(0, sy - h + 1))
            win = fx.bernoulli_map(w, h, float(rng.choice([0.0, 0.1, 0.5])), int(rng.integers(1 << 30)))
            hm[y0:y0 + h, x0:x0 + w] = win
            cm.update(win, x0, y0)
            log.append(f"update {w}x{h}@{x0},{y0}")
        elif op == "rewrite" and volatile:
            y0 = int(rng.integers(0, sy))
            h = int(rng.integers(1, sy - y0 + 1))
            win = fx.bernoulli_map(sx, h, float(rng.choice([0.0, 0.1, 0.4])), int(rng.integers(1 << 30)))
            hm[y0:y0 + h] = win
            view[y0:y0 + h, :sx].copy_(torch.from_numpy(win))
            torch.cuda.synchronize()
            log.append(f"rewrite rows {y0}..{y0 + h}")
        elif op == "volatile":
            volatile = not volatile
            cm.set_volatile(volatile)
            log.append(f"volatile={volatile}")
        # ---- a reader --------------------------------------------------------------------------------------------------
        rd = rng.choice(["host batch", "device batch", "band", "lethal", "problem"])
        n = int(rng.choice([50, 3000, 9000]))
        poses = fx.random_poses(n, sx, sy, int(rng.integers(1 << 30)), margin=float(reach + 1))
        good = True
        if rd == "host batch":
            good = fx.tol_ok(pg.get_cost_batch(cm, poses), oc.get_cost_batch(hm, poses, shift=True, nthreads=nt)).all()
        elif rd in ("device batch", "band"):
            rows = None
            if rd == "band":
                y_lo = float(rng.uniform(reach + 1, sy - reach - 2))
                y_hi = float(rng.uniform(y_lo, sy - reach - 2))
                poses[:, 1] = rng.uniform(y_lo, y_hi, n) if y_hi > y_lo else y_lo
                rows = row_band(y_lo, y_hi, reach, sy)
            d_p = torch.from_numpy(poses).cuda()
            d_o = torch.full((n, 4), -1.0, dtype=torch.float64, device="cuda")
            pg.get_cost_batch_device(cm, d_p.data_ptr(), n, d_o.data_ptr(), rows=rows)
            torch.cuda.synchronize()
            good = fx.tol_ok(d_o.cpu().numpy(), oc.get_cost_batch(hm, poses, shift=True, nthreads=nt)).all()
        elif rd == "lethal":
            ring = fx.rotated_ring(fx.to_rectangle((-int(rng.integers(1, 200)), -int(rng.integers(1, 150))), (int(rng.integers(1, 200)), int(rng.integers(1, 150)))),
                                   (rng.uniform(0, sx), rng.uniform(0, sy), rng.uniform(-3, 3)))
            got, want = dp.lethal_cells_within(cm, ring), capi.lethal_cells_within(hm, ring)
            good = got.shape == want.shape and np.array_equal(got, want)
        else:
            cfg = dict(lin_tolerance=12.0, rot_tolerance=0.8, weight_goal_lin=1, weight_goal_rot=1, weight_start_lin=1, weight_start_rot=1)
            goal, start = [float(rng.uniform(0, sx - 1)), float(rng.uniform(0, sy - 1)), 0.3], [float(rng.uniform(0, sx - 1)), float(rng.uniform(0, sy - 1)), -1.0]
            pr = dp.problem(pg, cm)
            pr.init(start, goal, cfg)
            op_ = capi.Problem(oc, hm)
            op_.init(start, goal, **cfg)
            x0 = dp.sample_offsets(int(rng.integers(1 << 30)), 12.0, 0.8, 200)
            ev = pr.eval_batch(x0)
            f_ref, g_ref, _, _ = op_.eval_batch(x0, nthreads=nt)
            good = fx.tol_ok(ev["f"], f_ref).all() and fx.tol_ok(ev["grad_f"], g_ref, rel=1e-5, abs_=2e-6).all()
        log.append(rd + ("" if good else " FAILED"))
        ops_done += 1
        if not good:
            fails.append({"sequence": s, "map": [sx, sy], "log": log[-12:]})
            break
    del view
print(json.dumps({"sequences": n_seq, "reader_checks": ops_done, "failures": len(fails), "first_failures": fails[:3],
                  "seconds": round(time.time() - t_start, 1)}))
sys.exit(1 if fails else 0)
