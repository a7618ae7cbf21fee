"""Probit BART on the GPU (the sampler behind bartCause::bartc, R/bcf_iv.R:78-80, 198-202, R/bcf_itt.R:166-170)
against its CPU oracle restatement (oracle/bcf_oracle.c "Probit BART"): the latent draw, the chain move by move, the
S-learner's individual effects, and the bartc()-shaped front end."""
import numpy as np
import pytest

from bcf_iv_b200 import _lib, bart, prep
from helpers import oracle_probit_bart, oracle_probit_sampler
from oracle import orc

pytestmark = pytest.mark.gpu


def _problem(n=900, p=6, seed=0, continuous=False):
    rng = np.random.default_rng(seed)
    X = rng.random((n, p)) if continuous else (rng.random((n, p)) > 0.5).astype(float)
    x1 = X[:, 0] > 0.5
    z = (rng.random(n) < 0.5).astype(np.int32)
    pr = 0.15 + 0.6 * z * x1 + 0.1 * (X[:, 1] > 0.5)
    y = (rng.random(n) < pr).astype(np.float64)
    return X, z, y, x1


def _gpu_sampler(y, X, z, seed=0, chain=0, engine=0, **kw):
    xc = np.asfortranarray(np.column_stack([X, z.astype(np.float64)]))
    cuts, off = prep.pack_cuts([bart.bart_cutpoints(xc[:, v]) for v in range(xc.shape[1])])
    s = _lib.Sampler(engine=engine, **bart.bart_config(seed=seed, chain=chain, treatment_col=X.shape[1], **kw))
    s.set_data(y, z, xc, None, cuts, off)
    return s


def test_latent_draw_matches_the_oracle_in_both_tails():
    rng = np.random.default_rng(1)
    m = np.concatenate([rng.normal(0, 1.5, 4000), [-9.0, 9.0, -25.0, 25.0, 0.0, -37.0, 37.0]])
    y = rng.integers(0, 2, m.size).astype(np.int32)
    u = np.concatenate([rng.random(m.size - 4), [2.0 ** -54, 1 - 2.0 ** -53, 0.5, 1e-12]])    # (the stream's uniforms lie in [2^-54, 1 - 2^-54])
    got = _lib.op_probit_latent(m, y, u)
    ref = np.array([orc.probit_latent(a, b, c) for a, b, c in zip(m, y, u)])
    assert np.all((got > 0) == (y == 1))                          # the latent sits on the side the response says
    assert np.all(np.isfinite(got)) and np.allclose(got, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("engine", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("continuous", [False, True])
def test_probit_chain_matches_oracle_move_by_move(engine, continuous):
    X, z, y, _ = _problem(n=1100, p=5, seed=3, continuous=continuous)
    g = _gpu_sampler(y, X, z, seed=5, engine=engine)
    o, perm = oracle_probit_sampler(y, X, z, seed=5, max_leaves=128)
    inv = np.empty(perm.size, np.int64); inv[perm] = np.arange(perm.size)
    for sweep in range(12):
        g.run(1, 0)
        o.probit_sweeps(1)
        gt, glr = g.trace()
        ot, olr = o.trace()
        assert np.array_equal(gt, ot), f"sweep {sweep}: move trace differs"
        m = np.isfinite(olr)
        assert np.allclose(glr[m], olr[m], rtol=1e-10, atol=1e-9)
    assert g.counters() == o.counters()
    assert g.scalars()["sigma"] == 1.0 and o.scalars()["sigma"] == 1.0
    assert np.allclose(g.fits()[0], o.fits()[0][inv], rtol=1e-9, atol=1e-9)
    assert g.verify_leaf_cache() == 0
    g.close(); o.close()


def test_individual_effects_match_the_oracle_and_find_the_effect():
    X, z, y, x1 = _problem(n=1500, p=6, seed=7)
    kw = dict(n_samples=150, n_burn=150, n_chains=2, seed=11)
    g = bart.probit_bart(y, X, z, **kw)
    o = oracle_probit_bart(y, X, z, **kw)
    assert np.allclose(g["ite_mean"], o["ite_mean"], rtol=1e-7, atol=1e-7)
    assert np.allclose(g["prob_mean"], o["prob_mean"], rtol=1e-7, atol=1e-7)
    assert g["ite_mean"][x1].mean() > 0.4 and abs(g["ite_mean"][~x1].mean()) < 0.15      # truth: .6 / 0
    assert abs(g["prob_mean"].mean() - y.mean()) < 0.03
    assert np.all(g["ite_var"] >= 0)


def test_bartc_front_end_and_thinning():
    X, z, y, x1 = _problem(n=800, p=5, seed=9)
    r = bart.bartc(y, z, X, n_samples=100, n_burn=100, n_chains=2, seed=4)
    assert set(r) == {"ite_mean", "ite_var", "p_score", "prob_mean"}
    assert abs(r["p_score"].mean() - 0.5) < 0.05                  # randomised treatment
    assert r["ite_mean"][x1].mean() > r["ite_mean"][~x1].mean() + 0.25
    ro = bart.bartc(y, z, X, n_samples=100, n_burn=100, n_chains=2, seed=4, fit=oracle_probit_bart)
    assert np.allclose(r["ite_mean"], ro["ite_mean"], rtol=1e-6, atol=1e-6) and np.allclose(r["p_score"], ro["p_score"], rtol=1e-6, atol=1e-6)
    # thinning follows bcf's rule (sweep it saved when it >= nburn and it % nthin == 0)
    g = _gpu_sampler(y, X, z, seed=2)
    g.run(7, 5, 3)
    o, perm = oracle_probit_sampler(y, X, z, seed=2, max_leaves=128)
    out = o.probit_run(7, 5, 3)
    inv = np.empty(perm.size, np.int64); inv[perm] = np.arange(perm.size)
    assert g.num_saved == out["saved"] == 5
    assert np.allclose(g.tau_mean(), out["ite_mean"][inv], rtol=1e-9, atol=1e-9)
    g.close(); o.close()


def test_probit_needs_a_binary_response():
    X, z, y, _ = _problem(n=300, p=4, seed=1)
    with pytest.raises(_lib.BcfGpuError, match="0/1"):
        _gpu_sampler(y + 0.5, X, z)
    with pytest.raises(_lib.BcfGpuError):
        _lib.Sampler(model=_lib.MODEL_PROBIT_BART)               # two forests / scales are not a probit BART


def test_sweeps_with_wide_trees_are_rerun_by_the_batched_kernel():
    """Sweep-by-sweep running queues its launches without asking after each: the pipelined kernel refuses a sweep that holds
    a tree with more than 8 keys and everything queued behind it returns; the host reruns from there with the batched
    kernel.  Few trees + a strong continuous signal grow such trees: the chain must still be the oracle's, move by move."""
    rng = np.random.default_rng(2)
    n = 4000
    X = rng.random((n, 3))
    z = (rng.random(n) < 0.5).astype(np.int32)
    pr = 0.5 + 0.45 * np.sin(9 * X[:, 0]) * np.cos(7 * X[:, 1])
    y = (rng.random(n) < pr).astype(np.float64)
    g = _gpu_sampler(y, X, z, seed=3, n_trees=4)
    o, perm = oracle_probit_sampler(y, X, z, seed=3, n_trees=4, max_leaves=128)
    g.run(150, 40)
    out = o.probit_run(150, 40, 1)
    inv = np.empty(perm.size, np.int64); inv[perm] = np.arange(perm.size)
    assert np.array_equal(g.trace()[0], o.trace()[0]) and g.counters() == o.counters()
    assert np.allclose(g.tau_mean(), out["ite_mean"][inv], rtol=1e-8, atol=1e-8)
    assert max((len(g.get_tree(t)[0]) + 1) // 2 for t in range(4)) > 8        # a tree wider than the pipelined kernel takes
    st = g.engine_stats
    assert st["fallback_sweeps"] > 0 and st["pipelined_sweeps"] > 0, st
    g.close(); o.close()


def test_latent_drawn_inside_the_sweep_kernel_is_the_stand_alone_draw(monkeypatch):
    """k_pipe<true> draws the next sweep's latent in its finishing pass (burn-in then runs many sweeps per launch);
    BCFGPU_NO_FUSED_LATENT=1 keeps round 2's form, k_probit_latent before every sweep.  Both walk the oracle's chain, through
    wide-tree fallbacks as well, and give bit-identical posterior means."""
    res = []
    for n, ntree, nb, ns in ((3000, 75, 40, 25), (4000, 4, 120, 30)):
        rng = np.random.default_rng(n)
        X = rng.random((n, 3))
        z = (rng.random(n) < 0.5).astype(np.int32)
        pr = 0.5 + 0.45 * np.sin(9 * X[:, 0]) * np.cos(7 * X[:, 1])
        y = (rng.random(n) < pr).astype(np.float64)
        o, perm = oracle_probit_sampler(y, X, z, seed=11, n_trees=ntree, max_leaves=128)
        out = o.probit_run(nb, ns, 1)
        inv = np.empty(perm.size, np.int64); inv[perm] = np.arange(perm.size)
        for fused in (True, False):
            if fused:
                monkeypatch.delenv("BCFGPU_NO_FUSED_LATENT", raising=False)
            else:
                monkeypatch.setenv("BCFGPU_NO_FUSED_LATENT", "1")
            g = _gpu_sampler(y, X, z, seed=11, n_trees=ntree)
            g.run(nb, ns)
            assert np.array_equal(g.trace()[0], o.trace()[0]) and g.counters() == o.counters(), (n, fused)
            assert np.allclose(g.tau_mean(), out["ite_mean"][inv], rtol=1e-8, atol=1e-8)
            res.append((g.tau_mean().copy(), g.fits()[0].copy(), g.kernel_launches))
            g.close()
        o.close()
        (ta, fa, la), (tb, fb, lb) = res[-2], res[-1]
        assert np.array_equal(ta, tb) and np.array_equal(fa, fb)
        assert la < lb                                            # fewer launches: no latent kernels, burn-in in few launches
