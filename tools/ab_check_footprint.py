"""A/B of the two check_footprint kernels (thread per pose vs CTA per pose) over the batch size.
usage (under gpurun): python tools/ab_check_footprint.py   -> JSON lines"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def child(coop_max, threads):
    import numpy as np
    import fixtures as fx
    import dpose_b200 as dp
    m = fx.bernoulli_map(2000, 2000, 0.0005, 3)
    cm = dp.Costmap(m)
    out = {}
    for name, fp in (("rect", fx.RECT), ("star40", fx.star40())):
        for n in (1, 128, 1024, 2048, 4096, 8192, 16384, 65536, 262144):
            poses = fx.random_poses(n, 2000, 2000, 7, margin=100.0)
            dp.check_footprint_batch(cm, fp, poses)
            reps = 200 if n <= 2048 else (50 if n <= 16384 else 10)
            t0 = time.perf_counter()
            for _ in range(reps):
                dp.check_footprint_batch(cm, fp, poses)
            out[f"{name}/{n}"] = round(1e6 * (time.perf_counter() - t0) / reps, 1)
    print(json.dumps({"coop_max": coop_max, "coop_threads": threads, "us_per_call": out}))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for cm, th in ((0, 0), (1 << 30, 128), (1 << 30, 64), (1 << 30, 32)):
            env = dict(os.environ, DPOSE_CHECK_COOP_MAX=str(cm), DPOSE_CHECK_COOP_THREADS=str(th))
            subprocess.run([sys.executable, __file__, str(cm), str(th)], env=env, check=True)
