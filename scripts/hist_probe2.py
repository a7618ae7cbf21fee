"""Timing experiments on k_hist_mma (BCFGPU_HIST_PROBE: 1 = no MMAs, 2 = no bins fetched, 3 = neither): which stage bounds the kernel."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcf_iv_b200 import _lib, workload
n, p = 1_000_000, 100
w = workload.synthetic_sampler_inputs(n, p, seed=0)
s = _lib.Sampler(lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1)
s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
r = np.random.default_rng(0).standard_normal(n)
for enc in ("i8",):
    for probe in (0, 1, 2, 3):
        os.environ["BCFGPU_HIST_PROBE"] = str(probe)
        for nleaf in (1, 4, 8, 12):
            slot = (np.arange(n) % nleaf).astype(np.int32)
            best = min(s.op_hist_all(1, slot, nleaf, r)[2] for _ in range(4))
            print(json.dumps({"enc": enc, "probe": probe, "nleaf": nleaf, "us": round(best * 1e3, 2)}), flush=True)
