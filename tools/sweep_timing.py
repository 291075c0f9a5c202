#!/usr/bin/env python
"""where the fixed cost of a thread-per-pose launch goes: device time per launch (clean plane, back to back) for batches of
exactly k full sweeps (k x 148 x 1024 poses) of random poses and of IDENTICAL poses (no lane / warp imbalance: what is left
is launch, staging, the first wave's ramp), on the cfg5 map.  One JSON line per case."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import dpose_b200 as dp
import fixtures as fx

side = 10000
fp, pad = fx.FOOTPRINTS["rect_pad2"]
m = (torch.rand((side, side), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) < 0.1).to(torch.uint8) * 254
pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
cm = dp.Costmap(device_ptr=m.data_ptr(), size_x=side, size_y=side, pitch=side, device=0)
st = torch.cuda.current_stream()
full = 148 * 1024
for kind in ("random", "identical"):
    for k in (1, 2, 4, 8, 16, 33):
        n = k * full
        g = torch.Generator(device="cuda").manual_seed(5)
        u = torch.rand((n, 3), generator=g, device="cuda", dtype=torch.float64)
        poses = torch.empty_like(u)
        poses[:, 0] = 32 + u[:, 0] * (side - 64); poses[:, 1] = 32 + u[:, 1] * (side - 64); poses[:, 2] = -np.pi + u[:, 2] * 2 * np.pi
        if kind == "identical":
            poses[:] = poses[0]
        out = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
        for _ in range(3):
            pg.get_cost_batch_device(cm, poses.data_ptr(), n, out.data_ptr(), stream=st.cuda_stream)
        torch.cuda.synchronize()
        reps = 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            pg.get_cost_batch_device(cm, poses.data_ptr(), n, out.data_ptr(), stream=st.cuda_stream)
        b.record(st)
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps
        print(json.dumps({"what": "thread-per-pose launch, clean plane, back to back", "poses": kind, "sweeps": k, "n": n,
                          "us_per_launch": us, "us_per_sweep": us / k}), flush=True)
