python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python - <<'PY'
import sys, time, json
sys.path.insert(0, "tests")
import numpy as np, fixtures as fx
import dpose_b200 as dp
from oracle import capi
m = fx.bernoulli_map(2000, 2000, 0.002, 1)
cm = dp.Costmap(m)
for name, fp in (("rect", fx.RECT), ("star40", fx.star40())):
    for n in (4096, 1000000):
        poses = fx.random_poses(n, 2000, 2000, 2)
        dp.check_footprint_batch(cm, fp, poses)
        t0 = time.perf_counter(); got = dp.check_footprint_batch(cm, fp, poses); tg = time.perf_counter() - t0
        t0 = time.perf_counter(); want = capi.check_footprint_batch(m, fp, poses, nthreads=capi.max_threads()); tc = time.perf_counter() - t0
        print(json.dumps({"what": "check_footprint_batch", "footprint": name, "poses": n, "gpu_call_us": tg * 1e6, "cpu_oracle_us": tc * 1e6,
                          "cpu_threads": capi.max_threads(), "free_fraction": float(want.mean()), "identical": bool(np.array_equal(got, want))}))
PY
