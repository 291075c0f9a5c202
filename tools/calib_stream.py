#!/usr/bin/env python
"""calibration: what streaming passes over a 100 MB uint8 buffer cost on this GPU (CUDA events, best of 20), next to
k_pack_tiles (100 MB read + 12.6 MB written) as timed through a volatile-map extraction / batch call"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
side = 10000
m = (torch.rand((side, side), device="cuda") < 0.1).to(torch.uint8) * 254
dst = torch.empty_like(m)
i32 = m.view(torch.int32)
def best(fn, reps=20):
    b = 1e9
    for _ in range(reps):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        b = min(b, a.elapsed_time(e) * 1e3)
    return b
for name, fn, nbytes in (("copy_ 100 MB (read + write)", lambda: dst.copy_(m), 2e8),
                         ("int32 sum over 100 MB (read only)", lambda: i32.sum(), 1e8),
                         ("uint8 == 254 count (read only)", lambda: (m == 254).sum(), 1e8)):
    us = best(fn)
    print(json.dumps({"what": name, "us": us, "gbs": nbytes / us / 1e3}))
