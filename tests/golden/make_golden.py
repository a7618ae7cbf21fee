"""Generates tests/golden/*.npz from the CPU oracle (oracle/bcf_oracle.c).

The reference has no golden vectors for this path and its sampler (third-party `bcf`) cannot run here, so
these fixtures pin OUR restatement: they freeze the oracle's output on fixed seeded problems so that (a) the oracle
cannot drift unnoticed and (b) the GPU box, where nothing but this repo exists, has recorded answers to compare with.
Inputs are regenerated from seeds by tests/helpers.make_problem (NumPy generator streams are stable across versions).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import make_problem, oracle_sampler  # noqa: E402

CASES = {
    # name: (problem kwargs, sampler kwargs, sweeps of the trace checkpoint, (nburn, nsim, nthin))
    "c1_readme": (dict(n=500, p=10, seed=3), dict(seed=11), 12, (20, 10, 2)),
    "continuous_ragged": (dict(n=1237, p=6, seed=5, continuous=True), dict(seed=2, ntree_con=40, ntree_mod=12), 10, (8, 6, 1)),
    "low_compliance_correlated": (dict(n=900, p=12, seed=8, rho=0.5, compliance=0.3), dict(seed=5, ntree_con=60, ntree_mod=20), 8, (6, 6, 1)),
}


def run_case(pk, sk, sweeps, run):
    pr = make_problem(**pk)
    o = oracle_sampler(pr, max_leaves=128, **sk)
    o.sweeps(sweeps)
    tr, lr = o.trace()
    sc = o.scalars()
    ac, am = o.fits()
    inv = pr.P.inv_perm
    sizes = np.array([len(o.get_tree(t)[0]) for t in range(o.T)])
    cnt = o.counters()
    o2 = oracle_sampler(pr, max_leaves=128, **sk)
    out = o2.run(*run)
    return dict(trace=tr, logr=lr, scalars=np.array([sc[k] for k in ("sigma", "mscale", "bscale1", "bscale0", "delta_con", "tau_con", "tau_mod")]),
                allfit_con=ac[inv], allfit_mod=am[inv], tree_sizes=sizes,
                counters=np.array([cnt[k] for k in ("birth_prop", "birth_acc", "death_prop", "death_acc")]),
                tau_mean=out["tau_mean"][inv], mu_mean=out["mu_mean"][inv], sigma=out["sigma"], y_check=pr.P.yscale[:8])


if __name__ == "__main__":
    for name, (pk, sk, sweeps, run) in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(pk, sk, sweeps, run))
        print("wrote", name)
