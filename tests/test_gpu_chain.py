"""GPU chain vs the CPU oracle, move by move (same Philox stream): the end-to-end parity test of the hot path.

Reference call sites: R/bcf_iv.R:89-91, R/bcf_itt.R:71-75 (bcf(...)$tau -> colMeans).  PARITY UNPINNED: the
oracle restates the un-vendored `bcf` package; agreement is with that restatement.
"""
import os

import numpy as np

import helpers
import pytest

from helpers import gpu_sampler, make_problem, oracle_sampler, tree_signature

pytestmark = pytest.mark.gpu

TOL = 1e-9   # fp64 state after many sweeps; per-step sums agree to ~1e-13 (fixed point 2^-44 vs sequential fp64)


def _compare_state(g, o, pr, check_trees=True):
    gs, os_ = g.scalars(), o.scalars()
    for k in ("sigma", "mscale", "bscale1", "bscale0", "delta_con", "tau_con", "tau_mod"):
        assert gs[k] == pytest.approx(os_[k], rel=TOL, abs=TOL), k
    assert gs["sweep"] == os_["sweep"]
    assert g.counters() == o.counters()
    gt, glr = g.trace()
    ot, olr = o.trace()
    assert np.array_equal(gt, ot), "move trace differs"
    m = np.isfinite(olr)
    assert np.array_equal(np.isfinite(glr), m)
    # log MH ratios: north_star asks 1e-10 relative
    assert np.allclose(glr[m], olr[m], rtol=1e-10, atol=1e-9)
    ac_g, am_g = g.fits()
    ac_o, am_o = o.fits()
    inv = pr.P.inv_perm
    assert np.allclose(ac_g, ac_o[inv], rtol=TOL, atol=TOL)
    assert np.allclose(am_g, am_o[inv], rtol=TOL, atol=TOL)
    if check_trees:
        for t in range(g.T):
            hg, vg, cg, mg = g.get_tree(t)
            ho, vo, co, mo = o.get_tree(t)
            assert tree_signature(hg, vg, cg) == tree_signature(ho, vo, co), f"tree {t} structure differs"
            og = np.argsort(hg); oo = np.argsort(ho)
            leaf = vg[og] < 0
            assert np.allclose(mg[og][leaf], mo[oo][leaf], rtol=TOL, atol=TOL)


@pytest.mark.parametrize("engine", [1, 2, 3, 4, 5])
def test_chain_matches_oracle_readme_shape(engine):
    """Config C1 shape (n_s = 500, p = 10 binary + pihat): 30 sweeps, compared after 1, 5 and 30."""
    pr = make_problem(n=500, p=10, seed=3)
    g = gpu_sampler(pr, seed=11, engine=engine)
    o = oracle_sampler(pr, seed=11, max_leaves=128)
    done = 0
    for upto in (1, 5, 30):
        g.run(upto - done, 0)
        o.sweeps(upto - done)
        done = upto
        _compare_state(g, o, pr)
    assert g.verify_leaf_cache() == 0


@pytest.mark.parametrize("engine", [1, 2, 3, 4, 5])
def test_chain_matches_oracle_continuous_covariates(engine):
    """Continuous X -> 10 000-cutpoint uint16 grids on most columns; ragged n (not a multiple of the tile)."""
    pr = make_problem(n=1237, p=6, seed=5, continuous=True)
    g = gpu_sampler(pr, seed=2, engine=engine, ntree_con=40, ntree_mod=12)
    o = oracle_sampler(pr, seed=2, max_leaves=128, ntree_con=40, ntree_mod=12)
    g.run(25, 0)
    o.sweeps(25)
    _compare_state(g, o, pr)
    assert g.verify_leaf_cache() == 0


def test_posterior_outputs_match_oracle():
    """run(nburn, nsim, nthin): saved draws, running means, sigma and scale traces."""
    pr = make_problem(n=400, p=10, seed=9)
    kw = dict(seed=5, ntree_con=30, ntree_mod=10)
    g = gpu_sampler(pr, keep_draws=7, **kw)
    o = oracle_sampler(pr, max_leaves=128, **kw)
    g.run(10, 8, 2)
    out = o.run(10, 8, 2, draws=True)
    assert g.num_saved == out["saved"] == 8
    inv = pr.P.inv_perm
    assert np.allclose(g.tau_mean(), out["tau_mean"][inv], rtol=TOL, atol=TOL)
    assert np.allclose(g.mu_mean(), out["mu_mean"][inv], rtol=TOL, atol=TOL)
    assert np.allclose(g.yhat_mean(), out["yhat_mean"][inv], rtol=TOL, atol=TOL)
    assert np.allclose(g.sigma(), out["sigma"], rtol=TOL)
    ms, bs = g.scales()
    assert np.allclose(ms, out["msd"], rtol=TOL) and np.allclose(bs, out["bsd"], rtol=TOL)
    assert np.allclose(g.draws("tau"), out["tau"][:, inv], rtol=TOL, atol=TOL)
    assert np.allclose(g.draws("mu"), out["mu"][:, inv], rtol=TOL, atol=TOL)
    assert np.allclose(g.draws("yhat"), out["yhat"][:, inv], rtol=TOL, atol=TOL)
    tv = g.tau_var()
    assert np.allclose(tv, out["tau"][:, inv].var(axis=0, ddof=1), rtol=1e-6, atol=1e-9)


def test_engines_are_bit_identical():
    """The residual is an integer (fixed-point) state updated exactly and every reduction is an integer sum, so the
    engines -- graph, persistent with the state in HBM, persistent with the state in shared memory, the batched
    engine (8 trees per pass, exact cross-count correction) and its pipelined version -- walk the same chain bit for bit whatever
    their thread / CTA partitioning."""
    pr = make_problem(n=3000, p=12, seed=21)
    kw = dict(seed=4, ntree_con=50, ntree_mod=20)
    a = gpu_sampler(pr, engine=1, **kw)
    b = gpu_sampler(pr, engine=3, **kw)
    c = gpu_sampler(pr, engine=2, **kw)
    e = gpu_sampler(pr, engine=4, **kw)
    f = gpu_sampler(pr, engine=5, **kw)
    for s in (a, b, c, e, f):
        s.run(5, 10)
    for other in (b, c, e, f):
        assert np.array_equal(a.tau_mean(), other.tau_mean())
        assert np.array_equal(a.mu_mean(), other.mu_mean())
        assert np.array_equal(a.sigma(), other.sigma())
        assert a.scalars() == other.scalars()
        assert a.counters() == other.counters()


def test_engine_handover():
    """The chain state parked in HBM is engine-independent: sweeps can alternate between engines."""
    pr = make_problem(n=2000, p=10, seed=8)
    kw = dict(seed=9, ntree_con=30, ntree_mod=10)
    ref = gpu_sampler(pr, engine=1, **kw)
    ref.run(12, 0)
    mix = gpu_sampler(pr, engine=2, **kw)
    mix.run(4, 0)
    o = oracle_sampler(pr, max_leaves=128, **kw)
    o.sweeps(12)
    # continue the resident chain and compare with the oracle after 12 sweeps in total
    mix.run(8, 0)
    _compare_state(mix, o, pr)
    _compare_state(ref, o, pr)


@pytest.mark.parametrize("engine", [2, 4, 5])
def test_chain_with_big_and_small_trees_mixed(engine):
    """Trees with more than 8 keys take the batched engine's solo path (shared-atomics table), the others are batched
    around them; install deep trees in both forests, then walk the chain against the oracle."""
    from test_gpu_kernels import _random_tree
    pr = make_problem(n=2500, p=8, seed=31, continuous=True)
    kw = dict(seed=6, ntree_con=21, ntree_mod=11)
    g = gpu_sampler(pr, engine=engine, **kw)
    o = oracle_sampler(pr, max_leaves=128, **kw)
    rng = np.random.default_rng(12)
    for tid, nl in ((0, 12), (3, 30), (4, 9), (9, 5), (20, 17), (21, 10), (25, 3), (31, 25)):
        cuts = pr.P.cuts_c if tid < 21 else pr.P.cuts_m
        heap, var, cut, mu = _random_tree(rng, [len(c) for c in cuts], nl)
        mu = mu * 0.05
        g.set_tree(tid, heap, var, cut, mu)
        o.set_tree(tid, heap, var, cut, mu)
    # the oracle keeps its fits incrementally: rebuild them from the installed trees
    o.refit()
    for k in (1, 3, 8):
        g.run(k, 0)
        o.sweeps(k)
        _compare_state(g, o, pr)
    assert g.verify_leaf_cache() == 0


def test_batched_engine_with_the_state_in_hbm(monkeypatch):
    """BCFGPU_NO_RESIDENT forces the batched kernel's HBM variant (what large n runs): same chain, bit for bit."""
    pr = make_problem(n=5000, p=9, seed=14, continuous=True)
    kw = dict(seed=8, ntree_con=40, ntree_mod=16)
    a = gpu_sampler(pr, engine=4, **kw)
    a.run(6, 6)
    assert a.engine_in_use[0] == 3
    monkeypatch.setenv("BCFGPU_NO_RESIDENT", "1")
    b = gpu_sampler(pr, engine=4, **kw)
    b.run(6, 6)
    assert b.engine_in_use[0] == 5
    assert a.scalars() == b.scalars() and a.counters() == b.counters()
    assert np.array_equal(a.tau_mean(), b.tau_mean()) and np.array_equal(a.trace()[0], b.trace()[0])
    assert b.verify_leaf_cache() == 0
    o = oracle_sampler(pr, max_leaves=128, **kw)
    o.sweeps(12)
    _compare_state(b, o, pr)


@pytest.mark.parametrize("engine", [4, 5])
@pytest.mark.parametrize("ntree", [(7, 0), (5, 3), (1, 1), (13, 6)])
def test_odd_forest_sizes(engine, ntree):
    """Forests that do not fill the engines' batches (and no tau forest at all): batch boundaries, forest switch."""
    pr = make_problem(n=900, p=8, seed=23)
    kw = dict(seed=3, ntree_con=ntree[0], ntree_mod=ntree[1])
    g = gpu_sampler(pr, engine=engine, **kw)
    o = oracle_sampler(pr, max_leaves=128, **kw)
    for k in (1, 7):
        g.run(k, 0)
        o.sweeps(k)
        _compare_state(g, o, pr)
    assert g.verify_leaf_cache() == 0


def test_shared_covariate_matrix_is_copied_once_and_changes_nothing():
    """bcf(y, z, x, x, pihat): x_moderate handed over as the leading columns of x_control's buffer is binned and copied
    once (bcfgpu_h2d_bytes); the chain is bit-identical to the one fed two separate matrices."""
    pr = make_problem(n=1200, p=9, seed=5)
    P = pr.P
    assert np.shares_memory(P.x_m, P.x_c)
    a = gpu_sampler(pr, seed=4)
    shared_bytes = a.h2d_bytes
    a.run(3, 4)
    from bcf_iv_b200 import _lib
    b = _lib.Sampler(lambda_=P.lambda_, sigma_init=P.sigma_init, con_sd=P.con_sd, mod_sd=P.mod_sd, seed=4)
    b.set_data(P.yscale, P.z, P.x_c, np.array(P.x_m, order="F"), pr.cuts_c_flat, pr.off_c, pr.cuts_m_flat, pr.off_m)
    # (the second copy of these two-valued columns crosses the bus as bytes: set_data's staging pass compares on the way)
    assert b.h2d_bytes - shared_bytes == P.x_m.size * (8 if os.environ.get("BCFGPU_NO_HOST_PACK") else 1)
    b.run(3, 4)
    assert a.trace()[0].tobytes() == b.trace()[0].tobytes()
    assert a.tau_mean().tobytes() == b.tau_mean().tobytes()
    assert a.sigma().tobytes() == b.sigma().tobytes()
    for t in range(a.T):
        assert tree_signature(*a.get_tree(t)[:3]) == tree_signature(*b.get_tree(t)[:3])


@pytest.mark.parametrize("max_ctas,n", [(3, 22000), (2, 12000)])
def test_pipelined_engine_with_several_tiles_per_warp(monkeypatch, max_ctas, n):
    """More than 27 tiles per CTA (here forced with BCFGPU_MAX_CTAS; at full grid: n > 1.02e6): the tile warps take several
    tiles each, rebuild the previous batch's columns from the stored keys, and the batched kernel behind k_pipe keeps its
    state in HBM.  Same chain as the batched engine, bit for bit."""
    monkeypatch.setenv("BCFGPU_MAX_CTAS", str(max_ctas))
    pr = make_problem(n=n, p=8, seed=31)
    out = {}
    for engine in (5, 4):
        g = gpu_sampler(pr, engine=engine, seed=3, ntree_con=20, ntree_mod=8)
        g.run(4, 4)
        assert g.verify_leaf_cache() == 0
        out[engine] = (g.trace()[0].tobytes(), g.tau_mean().tobytes(), g.sigma().tobytes(), g.engine_in_use)
    assert out[5][3][0] == 4                      # k_pipe really ran
    assert out[5][:3] == out[4][:3]


def test_a_share_of_the_gpu_walks_the_same_chain():
    """config.max_ctas (several fits side by side on one GPU): the chain does not depend on how many CTAs replay it."""
    pr = make_problem(n=9000, p=8, seed=41)
    out = []
    for max_ctas in (0, 74, 5):
        g = gpu_sampler(pr, seed=6, ntree_con=24, ntree_mod=8, max_ctas=max_ctas)
        g.run(4, 4)
        assert g.verify_leaf_cache() == 0
        out.append((g.trace()[0].tobytes(), g.tau_mean().tobytes(), g.sigma().tobytes()))
    assert out[0] == out[1] == out[2]


def test_saved_forests_predict_in_and_out_of_sample():
    """keep_draws bit 3: every saved sweep's forests stay on the device; bcfgpu_predict walks them over new rows binned with
    the training cutpoints.  In sample the predictions ARE the fit's tau / mu draws; out of sample they equal
    the oracle's pointer-tree walk over the raw doubles of the same chain, draw by draw."""
    pr = make_problem(n=900, p=8, seed=14, continuous=True)
    kw = dict(seed=8, ntree_con=40, ntree_mod=15)
    g = gpu_sampler(pr, keep_draws=7 | 8, **kw)
    g.run(12, 9, 2)
    nd, nn = g.saved_forests
    assert nd == 9 and nn >= 9 * 55
    P = pr.P
    ins = g.predict(P.x_c, P.x_m, draws=True)
    assert np.allclose(ins["tau"], g.draws("tau"), rtol=1e-12, atol=1e-12)
    assert np.allclose(ins["mu"], g.draws("mu"), rtol=1e-12, atol=1e-12)
    assert np.allclose(ins["tau_mean"], g.tau_mean(), rtol=1e-12, atol=1e-12)
    assert np.allclose(ins["tau_var"], g.tau_var(), rtol=1e-8, atol=1e-12)
    # new rows: some inside the training range, some outside it (bins clamp at the end cutpoints, as x < cut comparisons do)
    rng = np.random.default_rng(4)
    m = 333
    xn = np.column_stack([rng.integers(0, 2, (m, 2)).astype(float), rng.uniform(-0.3, 1.3, (m, P.x_m.shape[1] - 2))])
    xcn = np.column_stack([xn, 0.2 * rng.standard_normal(m)])
    out = g.predict(xcn, xn, draws=True)
    o = oracle_sampler(pr, max_leaves=128, **kw)
    ref = o.run_predict(12, 9, 2, xcn, xn, draws=True)
    assert ref["saved"] == 9
    assert np.allclose(out["tau"], ref["tau"], rtol=TOL, atol=TOL)
    assert np.allclose(out["tau_mean"], ref["tau_mean"], rtol=TOL, atol=TOL)
    assert np.allclose(out["mu_mean"], ref["mu_mean"], rtol=TOL, atol=TOL)
    # a run without the bit has nothing to predict from
    g2 = gpu_sampler(pr, **kw)
    g2.run(3, 2)
    import pytest as _pt
    from bcf_iv_b200 import _lib
    with _pt.raises(_lib.BcfGpuError, match="no saved forests"):
        g2.predict(xcn, xn)
    g.close(); g2.close(); o.close()


def test_bcf_front_end_predicts_at_new_rows():
    from bcf_iv_b200 import bcf, datasets
    d = datasets.generate_dataset_wide(n=700, p=6, seed=6)
    pihat = 0.05 * np.random.default_rng(1).standard_normal(700)
    fit = bcf.bcf(d["y"], d["z"], d["X"], d["X"], pihat, nburn=120, nsim=40, random_seed=5, n_chains=2,
                  x_predict_control=d["X"][:50], pihat_predict=pihat[:50], predict_draws=True)
    assert fit["tau_predict"].shape == (80, 50)
    # predicting at training rows reproduces their fitted draws (chains stacked the same way)
    assert np.allclose(fit["tau_predict"], fit["tau"][:, :50], rtol=1e-10, atol=1e-10)
    assert np.allclose(fit["tau_predict_mean"], fit["tau_mean"][:50], rtol=1e-10, atol=1e-10)
    assert np.allclose(fit["mu_predict_mean"], fit["mu_mean"][:50], rtol=1e-10, atol=1e-10)


def test_saved_forests_with_wide_trees():
    """keep_draws bit 3 runs the sampling part sweep by sweep; sweeps the pipelined kernel refuses (a tree with more than 8
    keys) are rerun by the batched kernel: the saved forests still reproduce the fit's draws."""
    pr = make_problem(n=5000, p=4, seed=33, continuous=True)
    # a strong smooth signal on a continuous covariate and only three mu trees: they must grow wide
    from bcf_iv_b200 import prep
    d = pr.data
    ywide = 3.0 * np.sin(9.0 * d["X"][:, 2]) + d["y"]
    pr.P = helpers.prepare(ywide, d["z"], d["X"], d["X"], 0.05 * np.random.default_rng(1).standard_normal(pr.n))
    pr.cuts_c_flat, pr.off_c = prep.pack_cuts(pr.P.cuts_c)
    pr.cuts_m_flat, pr.off_m = prep.pack_cuts(pr.P.cuts_m)
    kw = dict(seed=12, ntree_con=3, ntree_mod=2)
    g = gpu_sampler(pr, keep_draws=7 | 8, **kw)
    g.run(120, 30)
    assert max((len(g.get_tree(t)[0]) + 1) // 2 for t in range(5)) > 8
    assert g.engine_stats["fallback_sweeps"] > 0
    ins = g.predict(pr.P.x_c, pr.P.x_m, draws=True)
    assert np.allclose(ins["tau"], g.draws("tau"), rtol=1e-12, atol=1e-12)
    assert np.allclose(ins["mu"], g.draws("mu"), rtol=1e-12, atol=1e-12)
    o = oracle_sampler(pr, max_leaves=128, **kw)
    out = o.run(120, 30, 1)
    assert np.allclose(g.tau_mean(), out["tau_mean"][pr.P.inv_perm], rtol=1e-7, atol=1e-7)
    g.close(); o.close()
