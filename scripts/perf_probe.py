"""Quick throughput probe of both engines on a synthetic headline-shaped problem (not the bench contract)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcf_iv_b200 import _lib, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--p", type=int, default=100)
ap.add_argument("--sweeps", type=int, default=50)
ap.add_argument("--warm", type=int, default=20)
ap.add_argument("--engines", default="4,2")
a = ap.parse_args()

t0 = time.time()
w = workload.synthetic_sampler_inputs(a.n, a.p, seed=0)
print(f"data gen {time.time()-t0:.1f}s", flush=True)
for eng in [int(e) for e in a.engines.split(",")]:
    s = _lib.Sampler(engine=eng, lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1)
    t0 = time.time()
    s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
    t_set = time.time() - t0
    s.run(a.warm, 0)
    s.run(a.sweeps, 0)
    ms = s.last_run_ms
    sc = s.scalars(); c = s.counters()
    sizes = [len(s.get_tree(t)[0]) for t in range(0, s.T, 10)]
    print(f"engine {eng}: set_data {t_set:.2f}s  {a.sweeps} sweeps in {ms:.1f} ms -> {1000*a.sweeps/ms:.1f} sweeps/s "
          f"({ms/a.sweeps/s.T*1000:.2f} us/tree)  sigma={sc['sigma']:.4f} acc={c} mean_nodes={np.mean(sizes):.2f}", flush=True)
    s.close()
