"""Diagnostic for tests/test_gpu_parity.py::test_batch_ragged_sizes_bit_equal: for every batch size, through the host-buffer
call and through the device call, how many result rows differ from the prefix of one large batch, where, and by how
much.  One JSON line per (path, n).  Run under gpurun."""
import json
import sys

import numpy as np
import torch

import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures as fx  # noqa: E402
import dpose_b200 as dp  # noqa: E402

m = fx.bernoulli_map(900, 700, 0.10, 5)
pg = dp.pose_gradient(fx.RECT)
cm = dp.Costmap(m)
big = 148 * 1024 * 2 + 4129
poses = fx.random_poses(big, 900, 700, 77, margin=20.0)
sizes = (4737, 4738, 4769, 8192, 8193, 9473, 20001, 148 * 128 + 31, 131072, 131073, 148 * 1024 - 1, 148 * 1024, 148 * 1024 + 1,
         148 * 1024 + 32 * 148 + 5, big - 1, big)

d_p = torch.from_numpy(poses).cuda()
d_o = torch.empty((big, 4), dtype=torch.float64, device="cuda")


def dev(n, off=0):
    d_o.zero_()
    torch.cuda.synchronize()
    pg.get_cost_batch_device(cm, d_p.data_ptr() + 24 * off, n, d_o.data_ptr(), jacobian=True, stream=0)
    torch.cuda.synchronize()
    return d_o[:n].cpu().numpy()


ref_dev = dev(big)
ref_host = pg.get_cost_batch(cm, poses)
ref_host2 = pg.get_cost_batch(cm, poses)


def report(tag, n, out, ref):
    bad = np.flatnonzero((out != ref).any(axis=1))
    rec = {"path": tag, "n": int(n), "rows_differing": int(bad.size)}
    if bad.size:
        d = np.abs(out[bad] - ref[bad])
        rec.update(first=bad[:12].tolist(), last=bad[-4:].tolist(), max_abs=float(d.max()),
                   rel=float((d[:, 0] / np.maximum(np.abs(ref[bad, 0]), 1e-300)).max()),
                   sample_out=out[bad[0]].tolist(), sample_ref=ref[bad[0]].tolist())
    print(json.dumps(rec), flush=True)


report("host_big_vs_dev_big", big, ref_host, ref_dev)
report("host_big_again", big, ref_host2, ref_host)
for n in sizes:
    report("dev", n, dev(n), ref_dev[:n])
    report("dev_again", n, dev(n), ref_dev[:n])
    report("host", n, pg.get_cost_batch(cm, poses[:n]), ref_dev[:n])
for off in (big - 5000, big - 9000, 1000):
    report("dev_tail_off%d" % off, 5000, dev(5000, off), ref_dev[off:off + 5000])
    report("host_tail_off%d" % off, 5000, pg.get_cost_batch(cm, poses[off:off + 5000]), ref_dev[off:off + 5000])
