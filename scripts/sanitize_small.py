"""Small chains through every engine, for compute-sanitizer (memcheck / racecheck): n = 700, 30 + 12 trees."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import gpu_sampler, make_problem  # noqa: E402

pr = make_problem(n=700, p=6, seed=5, continuous=True)
for engine in (5, 4, 2, 3, 1):
    s = gpu_sampler(pr, seed=2, engine=engine, ntree_con=30, ntree_mod=12)
    s.run(3, 2)
    print("engine", engine, s.engine_in_use, float(s.tau_mean().sum()), s.verify_leaf_cache(), flush=True)
    s.close()

# K2' on the tensor cores (k_hist_mma: TMA box + bulk copies + tcgen05) and the probit-BART sweep with the latent drawn in k_pipe<true>
import numpy as np  # noqa: E402
from bcf_iv_b200 import _lib, bart, prep  # noqa: E402

rng = np.random.default_rng(1)
n, p = 2100, 9
X = rng.integers(0, 2, (n, p)).astype(np.float64)
z = rng.integers(0, 2, n).astype(np.int32)
flat, off = prep.pack_cuts([np.array([0.5])] * p)
s = _lib.Sampler(lambda_=1.0, sigma_init=1.0, seed=1, ntree_con=2, ntree_mod=2)
s.set_data(rng.standard_normal(n), z, X, X, flat, off, flat, off)
for nleaf in (1, 5, 12):
    cnt, sm, _ = s.op_hist_all(1, rng.integers(-1, nleaf, n).astype(np.int32), nleaf, rng.standard_normal(n))
    print("k_hist_mma leaves", nleaf, int(cnt.sum()), float(sm.sum()), flush=True)
s.close()
y01 = (rng.random(n) < 0.4).astype(np.float64)
xc = np.asfortranarray(np.column_stack([X, z.astype(np.float64)]))
cuts, offb = prep.pack_cuts([bart.bart_cutpoints(xc[:, v]) for v in range(xc.shape[1])])
b = _lib.Sampler(**bart.bart_config(seed=3, treatment_col=p, n_trees=20))
b.set_data(y01, z, xc, None, cuts, offb)
b.run(6, 4)
print("probit BART", b.engine_in_use, float(b.tau_mean().sum()), flush=True)
b.close()
