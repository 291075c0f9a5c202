"""does the ORDER of the poses change the host-buffer call's time?  same process, same buffers, list order / tile order alternately"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dpose_b200 as dp, fixtures as fx
import bench
side, n = 10000, 10_000_000
dev = torch.device("cuda:0")
d_map = bench.device_map(torch, dev, side, 4)
fp, pad = fx.FOOTPRINTS["rect_pad2"]
pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
cm = dp.Costmap(device_ptr=d_map.data_ptr(), size_x=side, size_y=side, pitch=side)
cm.set_volatile(True)
full = bench.device_poses(torch, dev, n, side, 5).cpu().numpy()
perm = dp.order_poses_by_tile(full, side, side)
h_a = torch.empty((n, 3), dtype=torch.float64, pin_memory=True); h_a.numpy()[:] = full
h_b = torch.empty((n, 3), dtype=torch.float64, pin_memory=True); h_b.numpy()[:] = full[perm]
h_o = torch.empty((n, 4), dtype=torch.float64, pin_memory=True)
def t(h, k=10):
    pg.get_cost_batch(cm, h.numpy(), out=h_o.numpy())
    t0 = time.perf_counter()
    for _ in range(k): pg.get_cost_batch(cm, h.numpy(), out=h_o.numpy())
    return (time.perf_counter() - t0) / k * 1e3
for rep in range(3):
    print(json.dumps({"list_ms": t(h_a), "tile_ms": t(h_b)}), flush=True)
# device-only time of one chunk-sized launch in both orders
d_a, d_b = torch.from_numpy(full[:131072]).cuda(), torch.from_numpy(full[perm][:131072]).cuda()
d_o = torch.empty((131072, 4), dtype=torch.float64, device="cuda")
cm.set_volatile(False)
for name, d in (("list", d_a), ("tile", d_b)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5): pg.get_cost_batch_device(cm, d.data_ptr(), 131072, d_o.data_ptr())
    e0.record()
    for _ in range(50): pg.get_cost_batch_device(cm, d.data_ptr(), 131072, d_o.data_ptr())
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"chunk_kernel_us_" + name: e0.elapsed_time(e1) * 1e3 / 50}), flush=True)
