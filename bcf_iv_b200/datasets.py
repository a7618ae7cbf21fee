"""Workload generator: the reference's `generate_dataset` (R/generate_dataset.R:22-68), restated.

`generate_dataset` keeps the reference's behaviour (needs p >= 10, returns a 10-column X whatever p is,
R/generate_dataset.R:31-41).  `generate_dataset_wide` is the generalisation SURVEY 8d calls for (all p columns
kept; optional continuous covariates and binary outcome) used by the measurement configs with p = 50/100/200.
The random stream is NumPy's, not R's: draws differ from R for the same seed, the distribution does not.
"""
from __future__ import annotations

import numpy as np


def _equicorrelated_normals(rng, n, p, rho):
    # mvrnorm(n, 0, (1-rho) I + rho 11')  ==  sqrt(1-rho) e_ij + sqrt(rho) f_i
    if not (0.0 <= rho < 1.0):
        raise ValueError("rho must be in [0, 1)")
    e = rng.standard_normal((n, p))
    if rho == 0.0:
        return e
    f = rng.standard_normal((n, 1))
    return np.sqrt(1.0 - rho) * e + np.sqrt(rho) * f


def _potential_outcomes(rng, X, n, null, effect_size, compliance):
    x1, x2 = X[:, 0], X[:, 1]
    w1 = (rng.random(n) < compliance).astype(np.float64)      # rbinom(n, 1, compliance)
    y0 = rng.standard_normal(n)
    eff = np.where((x1 == 0) & (x2 == 0), effect_size,
                   np.where((x1 == 1) & (x2 == 1), -effect_size, null))
    y1 = y0 + w1 * eff
    z = (rng.random(n) < 0.5).astype(np.float64)              # rbinom(n, 1, 0.5)
    w = z * w1                                                # one-sided non-compliance (w0 = 0)
    y = z * y1 + (1.0 - z) * y0
    return y, z, w


def generate_dataset(n=1000, p=10, rho=0, null=0, effect_size=2, compliance=0.75, seed=None):
    """Same signature and return value as the reference (a dict with y, z, w, X); `seed` is an addition
    (R users call set.seed first)."""
    if p < 10:
        raise ValueError("generate_dataset needs p >= 10 (the reference indexes X[, 10])")
    rng = np.random.default_rng(seed)
    raw = _equicorrelated_normals(rng, n, p, float(rho))
    X = (raw > 0).astype(np.float64)[:, :10]                  # qbinom(pnorm(raw), 1, .5) == 1{raw > 0}; cols 1..10 kept
    y, z, w = _potential_outcomes(rng, X, n, null, effect_size, compliance)
    return {"y": y, "z": z, "w": w, "X": X}


def generate_dataset_wide(n=1000, p=10, rho=0.0, null=0.0, effect_size=2.0, compliance=0.75, seed=None,
                          continuous=False, binary_outcome=False):
    """All p columns kept.  continuous=True: X = Phi(raw) for columns 3.. (columns 1,2 stay binary so the
    effect cells are defined) -> 10 000-cutpoint uint16 grids.  binary_outcome=True: y = 1{y > 0} (config C4)."""
    if p < 2:
        raise ValueError("need p >= 2")
    rng = np.random.default_rng(seed)
    raw = _equicorrelated_normals(rng, n, p, float(rho))
    X = (raw > 0).astype(np.float64)
    if continuous and p > 2:
        from scipy.special import ndtr
        X[:, 2:] = ndtr(raw[:, 2:])
    y, z, w = _potential_outcomes(rng, X, n, null, effect_size, compliance)
    if binary_outcome:
        y = (y > 0).astype(np.float64)
    return {"y": y, "z": z, "w": w, "X": X}
