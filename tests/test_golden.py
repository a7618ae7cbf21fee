"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle): the oracle must keep
reproducing them (CPU test), and the CUDA path must match them on the GPU box, where neither the reference nor anything
else but this repo exists.  Integer fields bit-exact; fp64 within 1e-9 (libm differences in log / cos between hosts)."""
import os
import sys

import numpy as np
import pytest

from helpers import gpu_sampler, make_problem, oracle_sampler

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES, run_case  # noqa: E402

TOL = 1e-9


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    g = _load(name)
    r = run_case(*CASES[name])
    assert np.array_equal(r["y_check"], g["y_check"]), "the seeded problem itself changed"
    for k in ("trace", "tree_sizes", "counters"):
        assert np.array_equal(r[k], g[k]), k
    m = np.isfinite(g["logr"])
    assert np.array_equal(np.isfinite(r["logr"]), m) and np.allclose(r["logr"][m], g["logr"][m], rtol=1e-10, atol=TOL)
    for k in ("scalars", "allfit_con", "allfit_mod", "tau_mean", "mu_mean", "sigma"):
        assert np.allclose(r[k], g[k], rtol=TOL, atol=TOL), k


@pytest.mark.gpu
@pytest.mark.parametrize("engine", [1, 2, 4, 5])
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_golden(name, engine):
    g = _load(name)
    pk, sk, sweeps, run = CASES[name]
    pr = make_problem(**pk)
    assert np.array_equal(pr.P.yscale[:8], g["y_check"])
    s = gpu_sampler(pr, engine=engine, **sk)
    s.run(sweeps, 0)
    tr, lr = s.trace()
    assert np.array_equal(tr, g["trace"]), "move trace differs from the golden chain"
    m = np.isfinite(g["logr"])
    assert np.allclose(lr[m], g["logr"][m], rtol=1e-10, atol=TOL)
    sc = s.scalars()
    assert np.allclose([sc[k] for k in ("sigma", "mscale", "bscale1", "bscale0", "delta_con", "tau_con", "tau_mod")],
                       g["scalars"], rtol=TOL, atol=TOL)
    ac, am = s.fits()
    assert np.allclose(ac, g["allfit_con"], rtol=TOL, atol=TOL) and np.allclose(am, g["allfit_mod"], rtol=TOL, atol=TOL)
    assert np.array_equal([len(s.get_tree(t)[0]) for t in range(s.T)], g["tree_sizes"])
    c = s.counters()
    assert np.array_equal([c[k] for k in ("birth_prop", "birth_acc", "death_prop", "death_acc")], g["counters"])
    s.close()
    s2 = gpu_sampler(pr, engine=engine, **sk)
    s2.run(*run)
    assert np.allclose(s2.tau_mean(), g["tau_mean"], rtol=TOL, atol=TOL)
    assert np.allclose(s2.mu_mean(), g["mu_mean"], rtol=TOL, atol=TOL)
    assert np.allclose(s2.sigma(), g["sigma"], rtol=TOL)
