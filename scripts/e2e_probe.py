"""Where the end-to-end call spends its time: create, set_data (H2D + binning), run, read-back (wall clock, pinned host buffers)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bcf_iv_b200 import _lib, workload
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = workload.synthetic_sampler_inputs(n, 100, seed=0)
host, keep = {}, []
for k in ("y", "z", "x_con"):
    host[k], t = bench.pinned_like(np.asfortranarray(w[k]) if k == "x_con" else w[k]); keep.append(t)
host["x_mod"] = host["x_con"][:, :100]
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    s = _lib.Sampler(lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1)
    t1 = time.perf_counter()
    s.set_data(host["y"], host["z"], host["x_con"], host["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
    t2 = time.perf_counter()
    s.run(0, K)
    t3 = time.perf_counter()
    tau = s.tau_mean()
    t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms  set_data {1e3*(t2-t1):.1f} ms (h2d {s.h2d_bytes/1e6:.0f} MB)  run({K}) {1e3*(t3-t2):.1f} ms (device {s.last_run_ms:.1f})  "
          f"tau_mean {1e3*(t4-t3):.1f} ms  total {1e3*(t4-t0):.1f} ms -> {K/(t4-t0):.1f} sweeps/s", flush=True)
    s.close()
