"""Sweeps/s of the pipelined engine where the decision chain is the only thing that matters (small n), BCF and probit BART."""
import json, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import make_problem, gpu_sampler
from bcf_iv_b200 import _lib

out = {}
for n in (500, 50000, 300000):
    pr = make_problem(n=n, p=10, seed=3)
    g = gpu_sampler(pr, seed=7, engine=_lib.ENGINE_PIPELINED)
    g.run(200, 0)
    t0 = time.perf_counter(); g.run(2000, 0); dt = time.perf_counter() - t0
    out[f"bcf_n{n}"] = round(2000 / dt, 1)
    out[f"bcf_n{n}_device"] = round(2000 / (g.last_run_ms / 1e3), 1)
    g.close()
print(json.dumps(out))
