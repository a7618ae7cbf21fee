"""Wall time of whole bcf_iv() / bcf_itt() calls on BASELINE.json's shapes (one GPU), with the share spent in the sampler.

usage: python scripts/bcf_iv_walltime.py [c1] [c2] [c3]     (default: all three)
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcf_iv_b200 import bayesiv, bcf as bcfmod, datasets  # noqa: E402

which = [a for a in sys.argv[1:]] or ["c1", "c2", "c3"]
calls = []
_orig = bcfmod.bcf


def _timed_bcf(*a, **k):
    t0 = time.perf_counter()
    r = _orig(*a, **k)
    calls.append((time.perf_counter() - t0, sum(r["device_ms"]) * 1e-3, len(a[0])))
    return r


bayesiv._bcf.bcf = _timed_bcf


def run(name, fn, gen, **kw):
    t0 = time.perf_counter()
    d = gen()
    t_gen = time.perf_counter() - t0
    calls.clear()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = fn(d["y"], d["w"], d["z"], d["X"], **kw)
    wall = time.perf_counter() - t0
    sweeps = kw.get("n_burn", 500) + kw.get("n_sim", 500)
    print(json.dumps({"config": name, "wall_s": round(wall, 3), "data_gen_s": round(t_gen, 2), "sampler_calls": len(calls),
                      "n_s": calls[0][2], "sweeps_per_call": sweeps, "bcf_call_s": [round(c[0], 3) for c in calls],
                      "sampler_device_s": [round(c[1], 3) for c in calls],
                      "sweeps_per_s_device": [round(sweeps / c[1], 1) for c in calls],
                      "rows": (None if out is None else int(len(out)))}), flush=True)


if "c1" in which:   # README.md:79-97
    run("C1 bcf_iv README n=1000 p=10 n_burn=2000 n_sim=2000", bayesiv.bcf_iv,
        lambda: datasets.generate_dataset(n=1000, p=10, compliance=0.75, seed=1), n_burn=2000, n_sim=2000, max_depth=2)
if "c2" in which:
    run("C2 bcf_itt n=1e5 p=50", bayesiv.bcf_itt,
        lambda: datasets.generate_dataset_wide(n=100000, p=50, compliance=0.75, seed=1), n_burn=500, n_sim=500)
if "c3" in which:
    run("C3 bcf_iv n=1e6 p=100 (one chain)", bayesiv.bcf_iv,
        lambda: datasets.generate_dataset_wide(n=1000000, p=100, compliance=0.75, seed=1), n_burn=500, n_sim=500)
