"""Repeatability stress: the same chain, many times, every engine; any run that differs from the first is a race."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcf_iv_b200 import _lib, workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
w = workload.synthetic_sampler_inputs(n, 100, seed=0)
ref = None
for engine in (5, 4, 2):
    bad = 0
    for rep in range(reps):
        s = _lib.Sampler(engine=engine, lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1)
        s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
        s.run(sweeps // 2, sweeps - sweeps // 2)
        got = (s.scalars()["sigma"], tuple(s.counters().values()), s.trace()[0].tobytes(), s.tau_mean().tobytes(), s.sigma().tobytes())
        mism = s.verify_leaf_cache()
        s.close()
        if ref is None: ref = got
        if got != ref or mism:
            bad += 1
            print(f"engine {engine} rep {rep}: DIFFERENT sigma {got[0]!r} vs {ref[0]!r} counters {got[1]} vs {ref[1]} cache mismatches {mism}", flush=True)
    print(f"engine {engine}: {bad} of {reps} runs differ", flush=True)

