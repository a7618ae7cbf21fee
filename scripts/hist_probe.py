"""K2' (all-candidate histogram) at the headline shape: device time and algorithmic GB/s (n * (9 + p * b_x))."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcf_iv_b200 import _lib, workload
n, p = 1_000_000, 100
w = workload.synthetic_sampler_inputs(n, p, seed=0)
s = _lib.Sampler(lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1)
s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
r = np.random.default_rng(0).standard_normal(n)
for path in ("k_hist_mma (tcgen05 int8, TMA)", "k_hist_bits (CUDA cores, BCFGPU_NO_HIST_MMA=1)"):
    if path.startswith("k_hist_bits"):
        os.environ["BCFGPU_NO_HIST_MMA"] = "1"
    for nleaf in (1, 2, 4, 8, 12):
        if nleaf > 8 and path.startswith("k_hist_bits"):
            continue
        slot = (np.arange(n) % nleaf).astype(np.int32)
        best = 1e9
        for _ in range(5):
            cnt, sm, ms = s.op_hist_all(1, slot, nleaf, r)
            best = min(best, ms)
        byts = n * (9 + p * 1)
        print(json.dumps({"kernel": path, "forest": "moderate (100 uint8 columns)", "nleaf": nleaf, "ms": best, "algorithmic_GB": byts / 1e9,
                          "GBps": byts / 1e9 / (best * 1e-3), "frac_of_6545.9": byts / 1e9 / (best * 1e-3) / 6545.9}), flush=True)
