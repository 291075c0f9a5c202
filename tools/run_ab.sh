python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, time; sys.path.insert(0,'tests')
import numpy as np, fixtures as fx, dpose_b200 as dp
from oracle import capi
m = fx.bernoulli_map(2000, 2000, 0.10, 1); cm = dp.Costmap(m)
for name in ("ship_pad10", "arrow_pad3"):
    fp, pad = fx.FOOTPRINTS[name]
    pg = dp.pose_gradient(fp, dp.pose_gradient.parameter(pad))
    poses = fx.random_poses(200000, 2000, 2000, 2, margin=120.0)
    out = pg.get_cost_batch(cm, poses)
    t0 = time.perf_counter(); out = pg.get_cost_batch(cm, poses); dt = time.perf_counter() - t0
    ref = capi.CostData(fp, pad).get_cost_batch(m, poses[:2000], shift=True, nthreads=capi.max_threads())
    print(name, "table", pg.get_data().rows, pg.get_data().cols, "poses/s incl. PCIe %.3e" % (len(poses) / dt), "parity", bool(fx.tol_ok(out[:2000], ref).all()))
PY
