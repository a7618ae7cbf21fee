"""Host-side pieces that stay on the CPU: bcf's R-wrapper preamble (prep), the workload generator (datasets) and the
closed forms the bcf_iv / bcf_itt twins use in place of rpart / AER / stats (bayesiv).  No GPU needed."""
import numpy as np

import helpers
import pytest

from bcf_iv_b200 import bayesiv, datasets, prep


# ---- .cp_quantile / prepare (bcf's R wrapper) -------------------------------------------------------
def test_cutpoints_binary_few_valued_and_continuous():
    assert np.array_equal(prep.cp_quantile(np.array([0., 1, 1, 0, 1])), [0.5])
    assert np.allclose(prep.cp_quantile(np.array([1., 2, 4, 4, 1, 2, 7])), [1.5, 3.0, 5.5])
    with pytest.warns(UserWarning):
        assert np.array_equal(prep.cp_quantile(np.full(9, 3.0)), [3.0])
    x = np.random.default_rng(0).standard_normal(5000)
    c = prep.cp_quantile(x)
    assert len(c) == 10000 and np.all(np.diff(c) >= 0) and c[0] == x.min() and c[-1] <= x.max()
    # approxfun(sort(x), quantile(x, k/n)) is close to the identity on [min, max], evaluated on an equispaced grid
    assert np.allclose(c, np.linspace(x.min(), x.max(), 10000), atol=0.05 * (x.max() - x.min()))


def test_prepare_follows_bcf_conventions():
    d = datasets.generate_dataset_wide(n=300, p=4, seed=1)
    pihat = np.linspace(-1, 1, 300)
    P = helpers.prepare(d["y"], d["z"], d["X"], d["X"], pihat)
    assert P.x_c.shape == (300, 5) and P.x_m.shape == (300, 4)           # include_pi = "control"
    assert np.isclose(P.yscale.mean(), 0, atol=1e-12) and np.isclose(P.yscale.std(ddof=1), 1)
    z = P.z[P.perm]
    assert np.all(z[:P.ntrt] == 1) and np.all(z[P.ntrt:] == 0)            # treated rows first ...
    assert np.all(np.diff(P.perm[:P.ntrt]) > 0) and np.all(np.diff(P.perm[P.ntrt:]) > 0)   # ... ties in original order
    assert P.con_sd == 2.0 and P.mod_sd == pytest.approx(1 / 0.674)
    from scipy.stats import chi2
    assert P.lambda_ == pytest.approx(P.sighat ** 2 * chi2.ppf(0.1, 3) / 3)
    A = np.column_stack([np.ones(300), d["z"], P.x_c])
    res = P.yscale - A @ np.linalg.lstsq(A, P.yscale, rcond=None)[0]
    assert P.sighat == pytest.approx(np.sqrt(res @ res / (300 - np.linalg.matrix_rank(A))))
    with pytest.raises(ValueError):
        helpers.prepare(d["y"], d["z"] + 2, d["X"], d["X"], pihat)
    with pytest.raises(ValueError):
        helpers.prepare(np.r_[d["y"][:-1], np.nan], d["z"], d["X"], d["X"], pihat)


# ---- generate_dataset (R/generate_dataset.R:22-68) ---------------------------------------------------------------
def test_generate_dataset_keeps_reference_quirks():
    d = datasets.generate_dataset(n=2000, p=15, seed=3)
    assert d["X"].shape == (2000, 10)                    # ten columns whatever p is (R/generate_dataset.R:31-41)
    with pytest.raises(ValueError):
        datasets.generate_dataset(n=100, p=5)            # the reference indexes X[, 10]
    assert set(np.unique(d["X"])) == {0.0, 1.0} and set(np.unique(d["z"])) == {0.0, 1.0}
    assert np.all(d["w"] <= d["z"])                      # one-sided non-compliance
    assert abs(d["w"][d["z"] == 1].mean() - 0.75) < 0.05
    x1, x2, z, y = d["X"][:, 0], d["X"][:, 1], d["z"], d["y"]
    for cell, itt in (((0, 0), 1.5), ((1, 1), -1.5), ((0, 1), 0.0)):
        m = (x1 == cell[0]) & (x2 == cell[1])
        assert y[m & (z == 1)].mean() - y[m & (z == 0)].mean() == pytest.approx(itt, abs=0.35)


def test_wide_generator_correlation_and_variants():
    d = datasets.generate_dataset_wide(n=20000, p=6, rho=0.5, seed=0)
    c = np.corrcoef(d["X"][:, 2], d["X"][:, 3])[0, 1]
    assert 0.25 < c < 0.42                               # tetrachoric 0.5 -> phi coefficient 1/3
    dc = datasets.generate_dataset_wide(n=500, p=5, continuous=True, binary_outcome=True, seed=1)
    assert len(np.unique(dc["X"][:, 3])) > 400 and set(np.unique(dc["y"])) <= {0.0, 1.0}


# ---- closed forms of the twins ------------------------------------------------------------------------------------
def test_oracle_glm_solves_the_score_equations_and_drops_aliased_columns():
    rng = np.random.default_rng(2)
    X = rng.standard_normal((800, 3))
    eta_true = 0.3 + X @ np.array([1.0, -0.5, 0.0])
    z = (rng.random(800) < 1 / (1 + np.exp(-eta_true))).astype(float)
    from oracle import glm_oracle
    eta, inf = glm_oracle.glm_logit(z, X, full=True)
    assert inf["converged"] and inf["rank"] == 4
    A = np.column_stack([np.ones(800), X])
    beta = np.linalg.lstsq(A, eta, rcond=None)[0]
    mu = 1 / (1 + np.exp(-eta))
    assert np.allclose(A.T @ (z - mu), 0, atol=1e-6)     # score equations of the MLE
    assert np.allclose(beta, [0.3, 1.0, -0.5, 0.0], atol=0.35)
    # an aliased column is dropped, as glm does
    eta2, inf2 = glm_oracle.glm_logit(z, np.column_stack([X, X[:, 0]]), full=True)
    assert np.allclose(eta, eta2, atol=1e-9) and inf2["rank"] == 4 and np.isnan(inf2["beta"][-1])


def test_oracle_lm_sigma_matches_lstsq_and_counts_the_rank():
    from oracle import glm_oracle
    rng = np.random.default_rng(3)
    X = (rng.random((600, 6)) < 0.5).astype(float)
    z = (rng.random(600) < 0.5).astype(int)
    y = 0.5 * z + X[:, 0] - X[:, 1] + rng.standard_normal(600)
    A = np.column_stack([np.ones(600), z, X])
    res = y - A @ np.linalg.lstsq(A, y, rcond=None)[0]
    assert np.isclose(glm_oracle.lm_sigma(y, z, X), np.sqrt(res @ res / (600 - 8)), rtol=1e-12)
    # pihat = a linear predictor of (1, X) is aliased in lm(y ~ z + [X | pihat]): the rank, and so sigma, do not change
    pihat = 0.2 + X @ np.arange(6.0)
    s2, inf = glm_oracle.lm_sigma(y, z, np.column_stack([X, pihat]), full=True)
    assert inf["rank"] == 8 and np.isnan(inf["beta"][-1]) and np.isclose(s2, np.sqrt(res @ res / (600 - 8)), rtol=1e-10)


def test_cart_matches_sklearn_and_rpart_rules():
    from sklearn.tree import DecisionTreeRegressor
    rng = np.random.default_rng(4)
    X = rng.integers(0, 2, (1500, 6)).astype(float)
    y = 1.5 * (X[:, 0] == 0) * (X[:, 1] == 0) - 1.5 * (X[:, 0] == 1) * (X[:, 1] == 1) + 0.3 * rng.standard_normal(1500)
    nodes, where = bayesiv.cart(X, y, maxdepth=2, cp=0.01, minsplit=10)
    sk = DecisionTreeRegressor(max_depth=2, min_samples_split=10, min_samples_leaf=3).fit(X, y)
    assert sorted(nodes) == [1, 2, 3, 4, 5, 6, 7]
    assert {nodes[1].var, nodes[2].var, nodes[3].var} == {0, 1} and nodes[1].cut == 0.5
    pred = np.array([nodes[w].mean for w in where])
    assert np.allclose(pred, sk.predict(X))
    assert bayesiv.rule_text(nodes[4].rule, [f"x{j+1}" for j in range(6)]).count("<") == 2
    assert np.array_equal(bayesiv.rule_mask(nodes[5].rule, X), where == 5)
    # cp prunes splits that explain less than cp of the root deviance; minsplit / minbucket stop small nodes
    nodes2, _ = bayesiv.cart(X, rng.standard_normal(1500), maxdepth=2, cp=0.01, minsplit=10)
    assert sorted(nodes2) == [1]
    nodes3, _ = bayesiv.cart(X[:12], y[:12], maxdepth=2, cp=0.0, minsplit=20)
    assert sorted(nodes3) == [1]


def test_tsls_matches_two_stage_least_squares():
    rng = np.random.default_rng(6)
    n = 4000
    z = rng.integers(0, 2, n).astype(float)
    w = z * (rng.random(n) < 0.7)
    y = 2.0 * w + rng.standard_normal(n)
    beta, p, pweak = bayesiv.tsls(y, w, z)
    # explicit two stages
    Z = np.column_stack([np.ones(n), z])
    what = Z @ np.linalg.lstsq(Z, w, rcond=None)[0]
    b2 = np.linalg.lstsq(np.column_stack([np.ones(n), what]), y, rcond=None)[0]
    assert beta == pytest.approx(b2[1], rel=1e-10)
    assert beta == pytest.approx(2.0, abs=0.15) and p < 1e-6 and pweak < 1e-6
    # Wald ratio identity: ITT / first stage
    itt = y[z == 1].mean() - y[z == 0].mean()
    fs = w[z == 1].mean() - w[z == 0].mean()
    assert beta == pytest.approx(itt / fs, rel=1e-10)
    # an irrelevant instrument is flagged by the weak-instrument test
    _, _, pw = bayesiv.tsls(y, rng.integers(0, 2, n).astype(float), z)
    assert pw > 0.01


def test_p_adjust_known_answers():
    p = np.array([0.01, 0.04, 0.03, 0.005])
    assert np.allclose(bayesiv.p_adjust(p, "bonferroni"), [0.04, 0.16, 0.12, 0.02])
    assert np.allclose(bayesiv.p_adjust(p, "holm"), [0.03, 0.06, 0.06, 0.02])
    assert np.allclose(bayesiv.p_adjust(p, "hochberg"), [0.03, 0.04, 0.04, 0.02])
    assert np.allclose(bayesiv.p_adjust(p, "BH"), [0.02, 0.04, 0.04, 0.02])
    assert np.allclose(bayesiv.p_adjust(p, "none"), p)
    assert np.all(bayesiv.p_adjust(p, "hommel") <= bayesiv.p_adjust(p, "holm") + 1e-12)
    assert np.all(bayesiv.p_adjust(p, "BY") >= bayesiv.p_adjust(p, "BH"))
    with pytest.raises(ValueError):
        bayesiv.p_adjust(p, "nope")


def test_bcf_front_end_validates_before_touching_the_device():
    from bcf_iv_b200 import bcf
    d = datasets.generate_dataset_wide(n=50, p=3, seed=0)
    with pytest.raises(TypeError):
        bcf.bcf(d["y"], d["z"], d["X"], d["X"], np.zeros(50))
    with pytest.raises(NotImplementedError):
        bcf.bcf(d["y"], d["z"], d["X"], d["X"], np.zeros(50), nburn=10, nsim=10, w=np.ones(50))
    with pytest.raises(ValueError):                       # binary = TRUE needs a 0/1 outcome (checked before any device work)
        bayesiv.bcf_iv(d["y"], d["w"], d["z"], d["X"], binary=True)


def test_prepare_shares_the_matrix_when_bcf_is_called_with_the_same_x_twice():
    """R/bcf_iv.R:89 calls bcf(y, z, x, x, pihat): x_moderate becomes the leading columns of x_control = [x | pihat]
    (one buffer), with the same values and cutpoints as two separate matrices would give."""
    from bcf_iv_b200 import prep
    rng = np.random.default_rng(2)
    n, p = 300, 6
    X = (rng.random((n, p)) < 0.4).astype(float)
    X[:, 3] = rng.standard_normal(n)
    y, z, pihat = rng.standard_normal(n), (rng.random(n) < 0.5).astype(int), 0.1 * rng.standard_normal(n)
    a = helpers.prepare(y, z, X, X, pihat)
    b = helpers.prepare(y, z, X, X.copy(), pihat)
    assert np.shares_memory(a.x_m, a.x_c) and a.x_c.flags.f_contiguous and a.x_m.flags.f_contiguous
    assert not np.shares_memory(b.x_m, b.x_c)
    assert np.array_equal(a.x_c, b.x_c) and np.array_equal(a.x_m, b.x_m)
    assert len(a.cuts_m) == len(b.cuts_m) == p
    for u, v in zip(a.cuts_m, b.cuts_m):
        assert np.array_equal(u, v)
    c = helpers.prepare(y, z, X, X, pihat, include_pi="both")      # pihat in both: two matrices again
    assert not np.shares_memory(c.x_m, c.x_c) and c.x_m.shape[1] == p + 1


def test_scan_columns_is_the_numpy_answer_for_every_kind_of_column():
    """bcfgpu_scan_columns (one threaded host pass of the library, no device): min, max, "two-valued" and "non-finite" per
    column, on a column-major matrix, on a view with a larger column stride, and on a row-major one"""
    from bcf_iv_b200 import _lib
    rng = np.random.default_rng(4)
    n = 10007
    cols = [rng.integers(0, 2, n).astype(float), np.full(n, 3.0), rng.standard_normal(n), rng.integers(0, 3, n).astype(float),
            np.where(rng.random(n) < 0.5, -1.5, 2.25), rng.standard_normal(n), rng.standard_normal(n)]
    cols[5][n - 1] = np.nan
    cols[6][17] = -np.inf
    X = np.asfortranarray(np.column_stack(cols))
    want_two = [1, 1, 0, 0, 1, 0, 0]
    want_bad = [0, 0, 0, 0, 0, 1, 1]
    for M in (X, np.asfortranarray(np.column_stack(cols + [np.zeros(n)] * 3))[:, :7], np.ascontiguousarray(X)):
        lo, hi, fl = _lib.scan_columns(M)
        assert list(fl & 1) == want_two and list((fl >> 1) & 1) == want_bad
        for v in range(5):
            assert lo[v] == X[:, v].min() and hi[v] == X[:, v].max()
    cuts = prep.make_cutpoints(X[:, :5])
    assert np.array_equal(cuts[0], [0.5]) and np.array_equal(cuts[4], [0.375])
    assert all(np.array_equal(a, prep.cp_quantile(X[:, v])) for v, a in enumerate(cuts) if v != 1)


def test_prepare_with_a_spare_column_builds_the_design_in_place():
    """bcf(x_spare_column=True): x_control is the leading columns of the caller's column-major buffer with one spare column;
    prepare() puts pihat there instead of copying n x p numbers -- same Prepared as the copying path"""
    from bcf_iv_b200 import bart
    rng = np.random.default_rng(9)
    n, p = 3001, 6
    Xall = np.column_stack([rng.integers(0, 2, (2 * n, p - 1)).astype(float), rng.standard_normal(2 * n)])
    rows = np.sort(rng.choice(2 * n, n, replace=False))
    y, z, pihat = rng.standard_normal(n), rng.integers(0, 2, n), 0.1 * rng.standard_normal(n)
    Xd = bart._as_design(Xall, 1, rows=rows)[:, :p]
    assert np.array_equal(Xd, Xall[rows]) and Xd.flags.f_contiguous
    P0 = helpers.prepare(y, z, Xall[rows], Xall[rows], pihat)
    P1 = helpers.prepare(y, z, Xd, Xd, pihat, spare_column=True)
    assert P1.x_c.ctypes.data == Xd.ctypes.data and P1.x_m.ctypes.data == Xd.ctypes.data      # in place
    assert np.array_equal(P0.x_c, P1.x_c) and np.array_equal(P0.x_m, P1.x_m)
    assert all(np.array_equal(a, b) for a, b in zip(P0.cuts_c, P1.cuts_c)) and len(P1.cuts_m) == p
    assert P0.sighat == P1.sighat and P0.lambda_ == P1.lambda_
    # not a leading-columns view of a buffer with room: the copying path, the caller's matrix untouched
    V = np.asfortranarray(Xall[rows])
    keep = V.copy()
    P2 = helpers.prepare(y, z, V, V, pihat, spare_column=True)
    assert np.array_equal(V, keep) and np.array_equal(P2.x_c, P0.x_c)
    with pytest.raises(ValueError):
        bad = Xd.copy(order="F"); bad[5, 2] = np.nan
        helpers.prepare(y, z, bad, bad, pihat)


def test_inference_rows_are_read_through_their_split_columns_only():
    """bcf_iv's per-node 2SLS looks at the inference sample only through the node rules: `_Rows` serves X[rows][:, v] one column
    at a time (no n x p row gather), and rule_mask over it equals rule_mask over the gathered matrix"""
    rng = np.random.default_rng(12)
    X = rng.random((500, 7))
    rows = rng.choice(500, 240, replace=False)
    view = bayesiv._Rows(X, rows)
    assert view.shape == (240, 7)
    rule = [(2, "<", 0.6), (5, ">=", 0.3), (2, ">=", 0.1)]
    assert np.array_equal(bayesiv.rule_mask(rule, view), bayesiv.rule_mask(rule, X[rows]))
    assert set(view._col) == {2, 5}                                  # only the split columns were ever gathered
    with pytest.raises(IndexError):
        view[3:5, 1]
