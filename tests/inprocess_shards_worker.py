"""Worker of tests/test_gpu_e2e.py::test_in_process_shards...: runs in its own process (a sweep kernel that gives up traps,
which ends the CUDA context) and prints one JSON line."""
import json
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bcf_iv_b200 import _lib  # noqa: E402
from helpers import gpu_sampler, make_problem  # noqa: E402

nshards, engine = int(sys.argv[1]), int(sys.argv[2])
n = 30011
pr = make_problem(n=n, p=8, seed=12)
P = pr.P
kw = dict(lambda_=P.lambda_, sigma_init=P.sigma_init, con_sd=P.con_sd, mod_sd=P.mod_sd, seed=5, ntree_con=30, ntree_mod=10)
shards = [_lib.Sampler(engine=engine, device=0, **kw) for _ in range(nshards)]     # (the library gives each its share of the SMs)
_lib.set_shard_local(shards)
bounds = [r * n // nshards for r in range(nshards + 1)]
errs = []


def work(r):
    try:
        lo, hi = bounds[r], bounds[r + 1]
        shards[r].set_data(P.yscale[lo:hi], P.z[lo:hi], P.x_c[lo:hi], P.x_m[lo:hi], pr.cuts_c_flat, pr.off_c, pr.cuts_m_flat, pr.off_m)
        shards[r].run(6, 6)
    except BaseException as e:
        errs.append(str(e))


th = [threading.Thread(target=work, args=(r,)) for r in range(nshards)]
for t in th:
    t.start()
for t in th:
    t.join(timeout=300)
if errs or any(t.is_alive() for t in th):
    print(json.dumps({"error": errs or ["a shard did not finish"]}), flush=True)
    os._exit(0)
s = gpu_sampler(pr, seed=5, ntree_con=30, ntree_mod=10, engine=1)
s.run(6, 6)
ref, sc = s.tau_mean(), s.scalars()
tau = np.concatenate([sh.tau_mean() for sh in shards])
out = {"tau_identical": bool(np.array_equal(tau, ref)),
       "scalars_identical": all(all(sh.scalars()[k] == sc[k] for k in ("sigma", "mscale", "bscale1", "bscale0", "tau_con", "sweep")) for sh in shards),
       "moves_identical": all(np.array_equal(sh.trace()[0], s.trace()[0]) and sh.counters() == s.counters() for sh in shards),
       "engine_in_use": [list(sh.engine_in_use) for sh in shards]}
print(json.dumps(out), flush=True)
