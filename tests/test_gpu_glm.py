"""GPU parity of the dense pre-steps through the C-ABI: bcfgpu_glm_logit / bcfgpu_lm_sigma against the CPU
restatement of R's glm.fit / lm (oracle/glm_oracle.py: Householder QR with dqrdc2's limited pivoting on the design itself).
Floating point: the two sides iterate the same IRLS map from the same start, so they agree far below glm's own
convergence tolerance; the tolerance written here is 1e-8 absolute on the log-odds (1e-9 relative on sigma)."""
from __future__ import annotations

import numpy as np
import pytest

from bcf_iv_b200 import _lib, bayesiv, datasets
from oracle import glm_oracle

pytestmark = pytest.mark.gpu
ETA_TOL = 1e-8


def _design(n, p, seed, continuous=False, rho=0.0):
    d = datasets.generate_dataset_wide(n=n, p=p, rho=rho, seed=seed, continuous=continuous)
    rng = np.random.default_rng(seed + 7)
    b = rng.normal(size=p) * (0.3 if not continuous else 0.15)
    z = (rng.random(n) < 1.0 / (1.0 + np.exp(-(d["X"] @ b - 0.2)))).astype(np.int32)
    return d["X"], z, d["y"]


@pytest.mark.parametrize("n,p,continuous,rho", [(500, 10, False, 0.0), (20000, 20, False, 0.5), (20000, 12, True, 0.0), (3000, 100, False, 0.0)])
def test_glm_logit_matches_the_oracle(n, p, continuous, rho):
    X, z, _ = _design(n, p, 11, continuous, rho)
    eta_o, io = glm_oracle.glm_logit(z, X, full=True)
    eta_g, ig = _lib.glm_logit(z, X, full=True)
    assert ig["converged"] and io["converged"] and ig["iterations"] == io["iterations"] and ig["rank"] == io["rank"] == p + 1
    assert ig["launches"] > 0
    assert np.max(np.abs(eta_g - eta_o)) <= ETA_TOL * max(1.0, np.max(np.abs(eta_o)))
    assert np.allclose(ig["beta"], io["beta"], rtol=1e-7, atol=1e-8)
    assert abs(ig["deviance"] - io["deviance"]) <= 1e-9 * io["deviance"]
    # the layout of the caller's matrix does not matter, bit for bit (row-major is transposed on the device)
    eta_f = _lib.glm_logit(z, np.asfortranarray(X))
    assert np.array_equal(eta_f, eta_g)
    # and the call is deterministic
    assert np.array_equal(_lib.glm_logit(z, X), eta_g)
    assert np.array_equal(bayesiv.logit_link(z, X), eta_g)


def test_glm_logit_aliased_columns_get_no_coefficient():
    X, z, _ = _design(5000, 8, 5)
    Xa = np.column_stack([X, X[:, 0], X[:, 1] + X[:, 2], np.ones(len(z))])       # a copy, a sum, a second intercept
    eta_o, io = glm_oracle.glm_logit(z, Xa, full=True)
    eta_g, ig = _lib.glm_logit(z, Xa, full=True)
    assert ig["rank"] == io["rank"] == 9
    assert np.array_equal(np.isnan(ig["beta"]), np.isnan(io["beta"])) and np.all(np.isnan(ig["beta"][-3:]))
    assert np.max(np.abs(eta_g - eta_o)) <= ETA_TOL
    assert np.max(np.abs(eta_g - _lib.glm_logit(z, X))) <= ETA_TOL                # same fit as without them


def test_glm_logit_wide_design_runs_the_gram_matrix_in_several_launches():
    X, z, _ = _design(20000, 200, 3)                     # config 5's p: 51 column blocks -> 1326 block pairs, 2 launches
    eta_o, io = glm_oracle.glm_logit(z, X, full=True)
    eta_g, ig = _lib.glm_logit(z, X, full=True)
    assert ig["iterations"] == io["iterations"] and ig["rank"] == 201
    assert np.max(np.abs(eta_g - eta_o)) <= ETA_TOL * max(1.0, np.max(np.abs(eta_o)))


def test_glm_logit_rejects_bad_input():
    X, z, _ = _design(200, 3, 1)
    with pytest.raises(_lib.BcfGpuError):
        _lib.glm_logit(z + 1, X)
    with pytest.raises(_lib.BcfGpuError):
        _lib.glm_logit(z, np.zeros((200, 400)))


@pytest.mark.parametrize("n,p,continuous", [(500, 10, False), (30000, 50, False), (20000, 12, True)])
def test_lm_sigma_matches_the_oracle(n, p, continuous):
    X, z, y = _design(n, p, 21, continuous)
    ys = (y - y.mean()) / y.std(ddof=1)
    pihat = glm_oracle.glm_logit(z, X)
    xc = np.column_stack([X, pihat])                     # bcf's x_control: pihat is a linear combination of (1, X)
    s_o, io = glm_oracle.lm_sigma(ys, z, xc, full=True)
    s_g, ig = _lib.lm_sigma(ys, z, xc, full=True)
    assert ig["rank"] == io["rank"] == p + 2 and np.isnan(ig["beta"][-1])
    assert abs(s_g - s_o) <= 1e-9 * s_o
    assert _lib.lm_sigma(ys, z, np.asfortranarray(xc)) == s_g
    s0_o, s0_g = glm_oracle.lm_sigma(ys, None, X), _lib.lm_sigma(ys, None, X)
    assert abs(s0_g - s0_o) <= 1e-9 * s0_o


def test_full_size_fit_solves_the_score_equations():
    """config 3's discovery sample (n_s = 5e5, p = 100): the oracle's QR would take minutes, the score equations of the
    maximum-likelihood fit A'(z - mu) = 0 and the normal equations of lm are size-independent checks."""
    n, p = 500_000, 100
    rng = np.random.default_rng(0)
    X = (rng.random((n, p)) < 0.5).astype(np.float64)
    b = rng.normal(size=p) * 0.1
    z = (rng.random(n) < 1.0 / (1.0 + np.exp(-(X @ b - 0.3)))).astype(np.int32)
    eta, info = _lib.glm_logit(z, X, full=True)
    assert info["converged"] and info["rank"] == p + 1
    mu = 1.0 / (1.0 + np.exp(-eta))
    score = np.concatenate([[np.sum(z - mu)], X.T @ (z - mu)])
    assert np.max(np.abs(score)) <= 1e-6 * n             # |score| / n: far below one observation's contribution
    assert np.allclose(info["beta"][1:], b, atol=0.05)
    y = X[:, 0] - X[:, 1] + 0.5 * z + rng.standard_normal(n)
    s, inf = _lib.lm_sigma(y, z, X, full=True)
    res = y - (inf["beta"][0] + inf["beta"][1] * z + X @ inf["beta"][2:])
    A1 = np.concatenate([[res.sum(), res @ z], X.T @ res])
    assert np.max(np.abs(A1)) <= 1e-7 * n and abs(s - np.sqrt(res @ res / (n - p - 2))) <= 1e-10
    assert abs(s - 1.0) < 0.01
