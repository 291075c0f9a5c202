#!/usr/bin/env python
"""one full-map lethal extraction on a clean and on a volatile 10000 x 10000 map (for ncu captures)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import fixtures as fx
import dpose_b200 as dp
side = 10000
g = torch.Generator(device="cuda"); g.manual_seed(4)
d_map = torch.where(torch.rand((side, side), generator=g, device="cuda") < 0.10, torch.tensor(254, dtype=torch.uint8, device="cuda"),
                    torch.tensor(0, dtype=torch.uint8, device="cuda")).contiguous()
cm = dp.Costmap(device_ptr=d_map.data_ptr(), size_x=side, size_y=side, pitch=side)
ring = fx.to_rectangle(side - 1, side - 1)
for volatile in (False, True):
    cm.set_volatile(volatile)
    for _ in range(3):
        c = dp.lethal_cells_within(cm, ring, on_device=True)
        n = len(c); del c
    print(volatile, n, cm.lethal_us())
