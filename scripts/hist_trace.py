"""Timeline of k_hist_mma's CTA 0 (BCFGPU_HIST_TRACE=1) at the headline shape."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcf_iv_b200 import _lib, workload
n, p = 1_000_000, 100
w = workload.synthetic_sampler_inputs(n, p, seed=0)
s = _lib.Sampler(lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1)
s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
r = np.random.default_rng(0).standard_normal(n)
nleaf = int(sys.argv[1]) if len(sys.argv) > 1 else 1
slot = (np.arange(n) % nleaf).astype(np.int32)
for _ in range(3):
    s.op_hist_all(1, slot, nleaf, r)
os.environ["BCFGPU_HIST_TRACE"] = "1"
print("ms", s.op_hist_all(1, slot, nleaf, r)[2], file=sys.stderr)
