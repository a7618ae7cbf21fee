"""Synthetic inputs of the sampler at the shapes BASELINE.json names: what bcf() hands to the native
call after the R wrapper's preamble, generated directly so that n = 1e6..1e7 does not need an R-style pipeline.

X is the reference's binary design (R/generate_dataset.R:25-30: 1{N(0, Sigma) > 0}) with ALL p columns kept, the
outcome follows R/generate_dataset.R:44-62, pihat is the logit of a randomised instrument (~ N(0, p/n)).
"""
from __future__ import annotations

import numpy as np

from . import prep


def synthetic_sampler_inputs(n, p, seed=0, rho=0.0, compliance=0.75, effect_size=2.0, continuous=False, analytic_cuts=False):
    """continuous: columns 3.. are X = Phi(raw) ~ U(0, 1) -> 10 000-cutpoint grids (uint16 bins), SURVEY 8d's stress variant;
    analytic_cuts: their cutpoints are the quantiles of U(0, 1) instead of the sample's (no sort of n values per column)."""
    rng = np.random.default_rng(seed)
    X = np.empty((n, p), dtype=np.float64, order="F")
    common = rng.standard_normal(n) * np.sqrt(rho) if rho > 0 else None
    for v in range(p):
        raw = rng.standard_normal(n) * np.sqrt(1.0 - rho)
        if common is not None:
            raw += common
        if continuous and v >= 2:
            from scipy.special import ndtr
            X[:, v] = ndtr(raw)
        else:
            X[:, v] = raw > 0
    x1, x2 = X[:, 0], X[:, 1]
    w1 = rng.random(n) < compliance
    y0 = rng.standard_normal(n)
    eff = np.where((x1 == 0) & (x2 == 0), effect_size, np.where((x1 == 1) & (x2 == 1), -effect_size, 0.0))
    z = (rng.random(n) < 0.5).astype(np.int32)
    y = y0 + z * w1 * eff
    pihat = rng.standard_normal(n) * np.sqrt(p / n)
    muy, sdy = float(y.mean()), float(y.std(ddof=1))
    yscale = (y - muy) / sdy
    x_con = np.empty((n, p + 1), dtype=np.float64, order="F")
    x_con[:, :p] = X
    x_con[:, p] = pihat
    cuts_x = []
    for v in range(p):
        if not (continuous and v >= 2):
            cuts_x.append(np.array([0.5]))
        elif analytic_cuts:
            cuts_x.append(np.linspace(0.0, 1.0, prep.CP_NUM + 2)[1:-1])
        else:
            cuts_x.append(prep.cp_quantile(X[:, v]))
    cuts_c = cuts_x + [prep.cp_quantile(pihat)]
    cc, oc = prep.pack_cuts(cuts_c)
    cm, om = prep.pack_cuts(cuts_x)
    # sighat: residual sd of y ~ z + x1*x2 cells is close to the noise sd; the lm over all columns is an O(n p^2)
    # host step that is not part of the sampler, so the probe uses the sd of yscale within cells
    cell = (2 * x1 + x2).astype(np.int64) * 2 + z
    sig = np.sqrt(np.mean([(yscale[cell == k]).var() for k in range(8) if np.any(cell == k)]))
    lam = sig ** 2 * prep.qchisq(0.1, 3.0) / 3.0
    return dict(y=yscale, z=z, x_con=x_con, x_mod=x_con[:, :p],   # x_mod: the leading columns of x_con, as in bcf(y, z, x, x, pihat)
                cuts_con=cc, off_con=oc, cuts_mod=cm, off_mod=om,
                lambda_=float(lam), sigma_init=float(yscale.std(ddof=1)), muy=muy, sdy=sdy, truth=eff * compliance)


# ---------------------------------------------------------------------------------------------------------------
# The same workload for observation shards (BASELINE config 5: n = 1e7, p = 200): every rank generates ONLY its rows.
# Rows come in blocks of BLOCK rows, block b drawn from the generator seeded (seed, b), so rows [lo, hi) are the same
# whatever the number of ranks; what the ranks must agree on bit for bit (the standardisation of y, sigma's prior scale
# lambda, the cutpoints of pihat) is computed from per-block sufficient statistics every rank can form (`allreduce`
# sums them over the ranks: torch.distributed in bench.py, identity on one rank).
# ---------------------------------------------------------------------------------------------------------------
BLOCK = 1 << 19


def _shard_block(seed, b, n, p, compliance, effect_size):
    """rows [b * BLOCK, min(n, (b + 1) * BLOCK)) of the binary design (rho = 0): X as uint8, raw y, z, pihat"""
    rows = min(n, (b + 1) * BLOCK) - b * BLOCK
    rng = np.random.default_rng([seed, b])
    X = rng.integers(0, 2, size=(p, rows), dtype=np.uint8)            # 1{N(0,1) > 0} for independent columns
    w1 = rng.random(rows) < compliance
    y0 = rng.standard_normal(rows)
    eff = np.where((X[0] == 0) & (X[1] == 0), effect_size, np.where((X[0] == 1) & (X[1] == 1), -effect_size, 0.0))
    z = (rng.random(rows) < 0.5).astype(np.int32)
    y = y0 + z * w1 * eff
    pihat = rng.standard_normal(rows) * np.sqrt(p / n)
    return X, y, z, pihat


def synthetic_shard_inputs(n, p, lo, hi, seed=0, compliance=0.75, effect_size=2.0, allreduce=None, pinned_alloc=None):
    """rows [lo, hi) of the n-row workload as the sampler takes them (column-major fp64 x_con = [X | pihat], x_mod = its
    leading p columns), plus the quantities every shard must share.  `allreduce(np.float64 array) -> summed array`.
    `pinned_alloc(shape_rows, cols) -> Fortran-ordered float64 array` lets the caller place x_con in pinned memory."""
    allreduce = allreduce or (lambda a: a)
    m = hi - lo
    x_con = pinned_alloc(m, p + 1) if pinned_alloc else np.empty((m, p + 1), dtype=np.float64, order="F")
    y = np.empty(m); z = np.empty(m, dtype=np.int32)
    # moments by (x1, x2, z) cell: count, sum y, sum y^2 -- for mean / sd of y and the within-cell variance (sighat)
    cell = np.zeros((8, 3))
    for b in range(lo // BLOCK, (hi - 1) // BLOCK + 1):
        Xb, yb, zb, pb = _shard_block(seed, b, n, p, compliance, effect_size)
        a0, a1 = max(lo, b * BLOCK) - b * BLOCK, min(hi, (b + 1) * BLOCK) - b * BLOCK
        d0 = b * BLOCK + a0 - lo
        for v in range(p):                                               # column by column: both sides contiguous
            x_con[d0:d0 + a1 - a0, v] = Xb[v, a0:a1]
        x_con[d0:d0 + a1 - a0, p] = pb[a0:a1]
        y[d0:d0 + a1 - a0] = yb[a0:a1]; z[d0:d0 + a1 - a0] = zb[a0:a1]
        c = (4 * Xb[0, a0:a1].astype(np.int64) + 2 * Xb[1, a0:a1] + zb[a0:a1])
        for k in range(8):
            v = yb[a0:a1][c == k]
            cell[k] += (v.size, v.sum(), (v * v).sum())
    cell = allreduce(cell.ravel().copy()).reshape(8, 3)
    cnt, s1, s2 = cell[:, 0], cell[:, 1], cell[:, 2]
    muy = s1.sum() / n
    sdy = np.sqrt((s2.sum() - n * muy * muy) / (n - 1))
    yscale = (y - muy) / sdy
    var_cell = (s2 - s1 * s1 / np.maximum(cnt, 1)) / np.maximum(cnt - 1, 1) / (sdy * sdy)
    sig = float(np.sqrt(np.mean(var_cell[cnt > 1])))
    lam = sig ** 2 * prep.qchisq(0.1, 3.0) / 3.0
    # pihat's cutpoint grid must be the same on every rank: pihat ~ N(0, p/n) -- its 10 000-point quantile grid is taken
    # from the distribution (deterministic, rank-free) instead of from the 1e7 draws
    from scipy.special import ndtri
    q = (np.arange(prep.CP_NUM) + 0.5) / prep.CP_NUM
    cut_pihat = np.sort(ndtri(q) * np.sqrt(p / n))
    cuts_x = [np.array([0.5])] * p
    cc, oc = prep.pack_cuts(cuts_x + [cut_pihat])
    cm, om = prep.pack_cuts(cuts_x)
    return dict(y=yscale, z=z, x_con=x_con, x_mod=x_con[:, :p], cuts_con=cc, off_con=oc, cuts_mod=cm, off_mod=om,
                lambda_=float(lam), sigma_init=1.0, muy=float(muy), sdy=float(sdy))
