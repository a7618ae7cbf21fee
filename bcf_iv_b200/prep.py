"""Host-side preprocessing that bcf's R wrapper does before and after the native call.

Restates the R code of `bcf::bcf` [bcf-recall; the package is imported
by the reference at NAMESPACE:9 and called at R/bcf_iv.R:89, R/bcf_itt.R:71].  The O(n p) bookkeeping (standardisation,
cutpoint grids, permutation) is NumPy on the host; the O(n p^2) step -- `lm` for sighat -- runs on the device.
"""
from __future__ import annotations

import warnings

import numpy as np

CP_NUM = 10000      # .cp_quantile(num=10000)
CP_CAT_LEVELS = 8   # .cp_quantile(cat_levels=8)


def cp_quantile(x: np.ndarray, num: int = CP_NUM, cat_levels: int = CP_CAT_LEVELS) -> np.ndarray:
    """Cutpoint grid of one covariate column (bcf's `.cp_quantile`).

    1 distinct value -> that value (with a warning); fewer than `cat_levels` distinct values ->
    midpoints between them; otherwise `num` points of approxfun(sort(x), quantile(x, 0:(n-1)/n))
    evaluated on seq(min, max, length.out=num).
    """
    x = np.asarray(x, dtype=np.float64)
    lo, hi = x.min(), x.max()
    if lo != hi and np.all((x == lo) | (x == hi)):
        return np.array([lo + (hi - lo) / 2.0])       # two-valued column (the reference's covariates): no sort needed
    ux = np.unique(x)
    if ux.size == 1:
        warnings.warn("A supplied covariate contains a single distinct value.")
        return x[:1].copy()
    if ux.size < cat_levels:
        return ux[:-1] + np.diff(ux) / 2.0
    n = x.size
    xs = np.sort(x)
    # R quantile type 7 at probabilities k/n, k = 0..n-1, computed directly on the sorted column
    hq = (n - 1) * (np.arange(n) / n)
    lo = np.floor(hq).astype(np.int64)
    hi = np.minimum(lo + 1, n - 1)
    qs = xs[lo] + (hq - lo) * (xs[hi] - xs[lo])
    # approxfun(ties = mean): collapse tied abscissae, averaging their ordinates
    uxs, inv = np.unique(xs, return_inverse=True)
    mq = np.bincount(inv, weights=qs) / np.bincount(inv)
    grid = np.linspace(xs[0], xs[-1], num)
    return np.interp(grid, uxs, mq)


def scan_columns(X: np.ndarray):
    """(lo, hi, flags) per column in one threaded host pass of the native library (bcfgpu_scan_columns; no device needed):
    flags bit 0 = two-valued or constant column, bit 1 = a NaN / Inf is present."""
    from . import _lib
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 2 and X.shape[1] == 0:
        return np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.uint8)
    return _lib.scan_columns(X)


def make_cutpoints(X: np.ndarray, scan=None):
    """Per-column cutpoints of an n x p matrix -> list of p sorted float64 vectors.  `scan`: scan_columns(X) if the caller has
    it already; two-valued columns (the reference's covariates) then cost nothing more."""
    X = np.asarray(X, dtype=np.float64)
    if X.shape[1] == 0:
        return []
    lo, hi, fl = scan if scan is not None else scan_columns(X)
    return [np.array([lo[v] + (hi[v] - lo[v]) / 2.0]) if (fl[v] & 1) and lo[v] != hi[v] else cp_quantile(X[:, v])
            for v in range(X.shape[1])]


def pack_cuts(cut_list):
    """list of per-variable cutpoint vectors -> (flat float64 cuts, int32 offsets[p+1]) as the C-ABI takes them."""
    off = np.zeros(len(cut_list) + 1, dtype=np.int32)
    for v, c in enumerate(cut_list):
        off[v + 1] = off[v] + len(c)
    flat = np.concatenate([np.asarray(c, dtype=np.float64) for c in cut_list]) if cut_list else np.zeros(0)
    return np.ascontiguousarray(flat), off


def lm_sigma(yscale: np.ndarray, z: np.ndarray, x_c: np.ndarray, device: int = 0) -> float:
    """summary(lm(yscale ~ z + x_c))$sigma: bcf's default `sighat`.  Runs on the device (bcfgpu_lm_sigma: Gram matrix of
    (1, z, x_c, y) + residual pass in CUDA, rank-aware solve of the normal equations); no host path."""
    from . import _lib
    return _lib.lm_sigma(yscale, z, x_c, device=device)


def qchisq(p: float, df: float) -> float:
    from scipy.special import chdtri            # (scipy.stats takes two seconds to import)
    return float(chdtri(df, 1.0 - p))           # chdtri inverts the upper tail


def _all_finite(a: np.ndarray) -> bool:
    """no NaN / Inf in a: one pass (a sum propagates both); the element-wise test only when the sum itself is not finite"""
    return bool(np.isfinite(a.sum()) or np.all(np.isfinite(a)))


class Prepared:
    """Everything the native sampler needs, in bcf's internal conventions."""

    __slots__ = ("n", "ntrt", "perm", "inv_perm", "muy", "sdy", "yscale", "z", "x_c", "x_m", "cuts_c", "cuts_m",
                 "sighat", "lambda_", "con_sd", "mod_sd", "sigma_init")


def _spare_column_buffer(x: np.ndarray, extra: int):
    """x as the leading columns of its own column-major base buffer with `extra` spare trailing columns, or None"""
    b = x.base
    if (isinstance(b, np.ndarray) and b.ndim == 2 and b.dtype == np.float64 and b.flags.f_contiguous and b.flags.writeable
            and x.ndim == 2 and b.shape[0] == x.shape[0] and b.shape[1] >= x.shape[1] + extra and x.strides == b.strides
            and x.ctypes.data == b.ctypes.data):
        return b
    return None


def prepare(y, z, x_control, x_moderate, pihat, *, include_pi="control", sd_control=None, sd_moderate=None,
            nu=3.0, lambda_=None, sigq=0.9, sighat=None, use_tauscale=True, sigma_fn=None, device=0,
            spare_column=False) -> Prepared:
    """bcf's R-side preamble [bcf-recall, SURVEY A.1].  Returns arrays in the ORIGINAL row order plus `perm`
    (treated rows first, stable: R's order(z, decreasing=TRUE)); the native library applies the permutation.
    `sigma_fn(yscale, z, x_c)`: what computes the default sighat (the device's lm_sigma; the parity tests hand in the
    CPU oracle's so that problems can be prepared without a GPU).  `spare_column`: the caller states that x_control is the
    leading columns of a column-major buffer of its own with a spare trailing column that may be overwritten: pihat goes
    there and the n x p copy into [x | pihat] is saved (what bcf_iv does with its discovery sample)."""
    same_x = x_moderate is x_control            # bcf(y, z, x, x, pihat): what BayesIV calls (R/bcf_iv.R:89)
    y = np.asarray(y, dtype=np.float64).ravel()
    z = np.asarray(z).ravel()
    x_control = np.asarray(x_control, dtype=np.float64)
    x_moderate = np.asarray(x_moderate, dtype=np.float64)
    if x_control.ndim == 1:
        x_control = x_control[:, None]
    if x_moderate.ndim == 1:
        x_moderate = x_moderate[:, None]
    pihat = np.asarray(pihat, dtype=np.float64)
    if pihat.ndim == 1:
        pihat = pihat[:, None]
    n = y.size
    if not (z.size == n and x_control.shape[0] == n and x_moderate.shape[0] == n and pihat.shape[0] == n):
        raise ValueError("y, z, x_control, x_moderate and pihat must have the same number of rows")
    if not (_all_finite(y) and _all_finite(pihat)):
        raise ValueError("Missing or non-finite values in y, x_control, x_moderate or pihat")
    if not np.all((z == 0) | (z == 1)):
        raise ValueError("z must be a vector of 0's and 1's")
    if include_pi not in ("control", "moderate", "both", "none"):
        raise ValueError("include_pi must be one of 'control', 'moderate', 'both', 'none'")
    if np.unique(y).size < 5:
        warnings.warn("y appears to be discrete")

    if same_x and include_pi == "control":
        # one column-major matrix [x | pihat]; x_moderate is its leading columns (the library then bins and copies it once)
        pc = x_control.shape[1]
        buf = _spare_column_buffer(x_control, pihat.shape[1]) if spare_column else None
        if buf is not None:
            x_c = buf[:, :pc + pihat.shape[1]]        # in place: no n x p copy
        else:
            x_c = np.empty((n, pc + pihat.shape[1]), dtype=np.float64, order="F")
            if x_control.flags.c_contiguous and pc > 1:
                for r0 in range(0, n, 8192):          # row-major -> column-major in cache-sized blocks (1.5x a plain assignment)
                    x_c[r0:r0 + 8192, :pc] = x_control[r0:r0 + 8192]
            else:
                x_c[:, :pc] = x_control
        x_c[:, pc:] = pihat
        x_m = x_c[:, :x_control.shape[1]]
        m_is_prefix = True
    else:
        m_is_prefix = False
        x_c = np.column_stack([x_control, pihat]) if include_pi in ("control", "both") else x_control
        x_m = np.column_stack([x_moderate, pihat]) if include_pi in ("moderate", "both") else x_moderate
    # one threaded pass over the covariates: the NA check and what .cp_quantile needs to know about every column
    scan_c = scan_columns(x_c)
    scan_m = None if m_is_prefix else scan_columns(x_m)
    if np.any(scan_c[2] & 2) or (scan_m is not None and np.any(scan_m[2] & 2)):
        raise ValueError("Missing or non-finite values in y, x_control, x_moderate or pihat")

    P = Prepared()
    P.n = n
    P.muy = float(np.mean(y))
    P.sdy = float(np.std(y, ddof=1))
    P.yscale = (y - P.muy) / P.sdy
    P.z = z.astype(np.int32)
    P.x_c, P.x_m = x_c, x_m
    P.cuts_c = make_cutpoints(x_c, scan_c)
    P.cuts_m = P.cuts_c[:x_m.shape[1]] if m_is_prefix else make_cutpoints(x_m, scan_m)
    P.sighat = float(sighat) if sighat is not None else float(sigma_fn(P.yscale, z, x_c) if sigma_fn else lm_sigma(P.yscale, z, x_c, device))
    P.lambda_ = float(lambda_) if lambda_ is not None else P.sighat ** 2 * qchisq(1.0 - sigq, nu) / nu
    P.perm = np.argsort(-P.z, kind="stable")          # treated first, ties in original order
    P.inv_perm = np.empty(n, dtype=np.int64)
    P.inv_perm[P.perm] = np.arange(n)
    P.ntrt = int(np.sum(P.z == 1))
    sdc = 2.0 * P.sdy if sd_control is None else float(sd_control)
    sdm = P.sdy if sd_moderate is None else float(sd_moderate)
    P.con_sd = 2.0 if np.isclose(sdc, 2.0 * P.sdy) else sdc / P.sdy
    mod = 1.0 if np.isclose(sdm, P.sdy) else sdm / P.sdy
    P.mod_sd = mod / 0.674 if use_tauscale else mod
    P.sigma_init = float(np.std(P.yscale, ddof=1))
    return P
