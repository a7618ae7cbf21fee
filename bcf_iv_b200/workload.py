"""Synthetic inputs of the sampler at the shapes BASELINE.json names: what bcf() hands to the native
call after the R wrapper's preamble, generated directly so that n = 1e6..1e7 does not need an R-style pipeline.

X is the reference's binary design (R/generate_dataset.R:25-30: 1{N(0, Sigma) > 0}) with ALL p columns kept, the
outcome follows R/generate_dataset.R:44-62, pihat is the logit of a randomised instrument (~ N(0, p/n)).
"""
from __future__ import annotations

import numpy as np

from . import prep


def synthetic_sampler_inputs(n, p, seed=0, rho=0.0, compliance=0.75, effect_size=2.0, continuous=False):
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
        cuts_x.append(np.array([0.5]) if not (continuous and v >= 2) else prep.cp_quantile(X[:, v]))
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
