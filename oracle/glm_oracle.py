"""CPU restatement of the two dense pre-steps of bcf_iv() / bcf() -- TEST INFRASTRUCTURE, never the product path.

  glm_logit : stats::glm(z ~ x, family = binomial) + predict(type = "link")     reference call sites R/bcf_iv.R:72-75,
              R/bcf_itt.R:58-61
  lm_sigma  : summary(lm(yscale ~ z + x_c))$sigma                                bcf's R wrapper

Both live in base R (package `stats`, not vendored by the reference, no R in this image), so they are restated from the
published algorithm: glm.fit's IRLS for binomial(link = "logit") -- mustart = (y + 0.5) / 2, eta clamped inside linkinv /
mu.eta at +-log(1 / DBL_EPSILON), mu.eta floored at DBL_EPSILON, convergence |dev - devold| / (|dev| + 0.1) < 1e-8, at
most 25 iterations -- around LINPACK's dqrdc2 as modified for R: Householder QR with LIMITED column pivoting, where a
column whose norm has shrunk below tol times its original norm is moved to the end and declared aliased (NA coefficient;
tol = 1e-11 in glm.fit, 1e-7 in lm).  Unlike the device path (normal equations, Cholesky) this works on the n x m design
itself.  Parity unpinned: the reference has no tests or golden values for these steps.
"""
from __future__ import annotations

import numpy as np

DBL_EPS = np.finfo(np.float64).eps
THRESH = -np.log(DBL_EPS)


def dqrls(A, y, tol):
    """Least squares of y on the columns of A by Householder QR with dqrdc2's limited pivoting.
    Returns (coef with NaN for aliased columns, rank)."""
    A = np.array(A, dtype=np.float64, order="F")
    y = np.array(y, dtype=np.float64)
    n, m = A.shape
    piv = list(range(m))
    norm0 = np.sqrt((A * A).sum(axis=0))
    norm0[norm0 == 0.0] = 1.0
    lup, l = m, 0
    while l < min(n, lup):
        # cycle negligible columns to the end
        while l < lup and np.sqrt((A[l:, l] ** 2).sum()) < norm0[piv[l]] * tol:
            A[:, l:] = np.concatenate([A[:, l + 1:], A[:, l:l + 1]], axis=1)
            piv.append(piv.pop(l))
            lup -= 1
        if l >= lup:
            break
        v = A[l:, l].copy()
        nrm = np.sqrt((v * v).sum())
        if v[0] < 0:
            nrm = -nrm
        v /= nrm
        v[0] += 1.0
        # apply H = I - v v' / v[0] to the remaining columns and to y
        rest = A[l:, l + 1:]
        rest -= np.outer(v, (v @ rest) / v[0])
        y[l:] -= v * ((v @ y[l:]) / v[0])
        A[l, l] = -nrm
        A[l + 1:, l] = 0.0
        l += 1
    k = min(lup, n)
    coef = np.full(m, np.nan)
    if k > 0:
        R = np.triu(A[:k, :k])
        b = np.linalg.solve(R, y[:k]) if k > 1 else y[:1] / R[0, 0]
        for a in range(k):
            coef[piv[a]] = b[a]
    return coef, k


def _linkinv(eta):
    e = np.clip(eta, -THRESH, THRESH)
    ex = np.exp(-np.abs(e))
    return np.where(e >= 0, 1.0 / (1.0 + ex), ex / (1.0 + ex))


def _mu_eta(eta):
    e = np.clip(eta, -THRESH, THRESH)
    ex = np.exp(-np.abs(e))
    return np.maximum(ex / ((1.0 + ex) * (1.0 + ex)), DBL_EPS)


def _deviance(z, eta):
    mu = _linkinv(eta)
    return float(-2.0 * np.sum(np.log(np.where(z == 1, mu, 1.0 - mu))))


def glm_logit(z, X, max_iter=25, epsilon=1e-8, full=False):
    z = np.asarray(z, dtype=np.float64).ravel()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    A = np.column_stack([np.ones(len(z)), X])
    eta = np.log(((z + 0.5) / 2.0) / (1.0 - (z + 0.5) / 2.0))
    devold = _deviance(z, eta)
    coef, rank, conv, it, dev = None, 0, False, 0, devold
    for it in range(1, max_iter + 1):
        mu = _linkinv(eta)
        me = _mu_eta(eta)
        t = eta + (z - mu) / me
        w = np.sqrt(me)                       # sqrt(weights * mu.eta^2 / variance(mu)), variance = mu (1 - mu) = mu.eta
        coef, rank = dqrls(A * w[:, None], t * w, min(1e-7, epsilon / 1000.0))
        eta = A @ np.nan_to_num(coef, nan=0.0)
        dev = _deviance(z, eta)
        if abs(dev - devold) / (abs(dev) + 0.1) < epsilon:
            conv = True
            break
        devold = dev
    if full:
        return eta, {"beta": coef, "deviance": dev, "iterations": it, "converged": conv, "rank": rank}
    return eta


def lm_sigma(y, z, X, full=False):
    y = np.asarray(y, dtype=np.float64).ravel()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    cols = [np.ones(len(y))] + ([] if z is None else [np.asarray(z, dtype=np.float64).ravel()]) + [X]
    A = np.column_stack(cols)
    coef, rank = dqrls(A, y, 1e-7)
    resid = y - A @ np.nan_to_num(coef, nan=0.0)
    sigma = float(np.sqrt(resid @ resid / max(len(y) - rank, 1)))
    if full:
        return sigma, {"beta": coef, "rank": rank}
    return sigma
