"""`bartc()` -- the drop-in for the reference's calls into bartCause (probit BART S-learner), on the GPU.

Reference call sites (fbargaglistoffi/BCF-IV):
    R/bcf_iv.R:78-80    pic_bart <- bartc(w[-index], z[-index], x[-index,], n.samples = n_sim, n.burn = n_burn, n.chains = 2L)
                        tau_pic <- extract(pic_bart, type = "ite");  pic <- apply(tau_pic, 2, mean)      (complier proportion)
    R/bcf_iv.R:198-202, R/bcf_itt.R:166-170   the same with the 0/1 outcome as the response when binary = TRUE       (the ITT)
bartCause / dbarts are third-party and un-vendored like bcf: what is restated is the published algorithm -- Chipman, George &
McCulloch (2010) BART with Albert & Chib's probit augmentation, in bartc's arrangement [bartCause-recall]:
    1. treatment model   P(z = 1 | x) = Phi(f_z(x))            -> p.score = its posterior mean          (method.trt = "bart")
    2. response model    P(resp = 1 | x, ps, z) = Phi(f(x, ps, z))   (p.scoreAsCovariate = TRUE, S-learner)
    3. individual effect per draw  Phi(f(x, ps, 1)) - Phi(f(x, ps, 0));  chains pooled (n.chains = 2)
with dbarts' defaults: 75 trees, k = 2 (leaf sd 3 / (k sqrt(T))), base .95, power 2, binaryOffset 0.
Both samplers are libbcfgpu's probit-BART model (include/bcfgpu.h: BCFGPU_MODEL_PROBIT_BART); no CPU path.
Cutpoints: a column with fewer than 8 distinct values is cut at the midpoints between them (so a 0/1 column has one
cutpoint), any other at dbarts' 100 equally spaced points between its minimum and maximum.
"""
from __future__ import annotations

import numpy as np

from . import _lib, prep

N_TREES, K_SD, BASE, POWER, NUMCUT = 75, 2.0, 0.95, 2.0, 100


def bart_cutpoints(x, numcut=NUMCUT, cat_levels=prep.CP_CAT_LEVELS):
    x = np.asarray(x, dtype=np.float64)
    lo, hi = x.min(), x.max()
    if lo != hi and np.all((x == lo) | (x == hi)):
        return np.array([lo + (hi - lo) / 2.0])       # two-valued column: no sort needed
    ux = np.unique(x)
    if ux.size == 1:
        return ux.copy()
    if ux.size < cat_levels:
        return ux[:-1] + np.diff(ux) / 2.0
    lo, hi = ux[0], ux[-1]
    return lo + (hi - lo) * np.arange(1, numcut + 1) / (numcut + 1.0)


def bart_config(n_trees=N_TREES, k=K_SD, base=BASE, power=POWER, seed=0, chain=0, treatment_col=-1, offset=0.0, **extra):
    """bcfgpu_config fields of a probit BART fit"""
    cfg = dict(model=_lib.MODEL_PROBIT_BART, ntree_con=int(n_trees), ntree_mod=0, alpha_con=float(base), beta_con=float(power),
               con_sd=3.0 / float(k), use_muscale=0, use_tauscale=0, sigma_init=1.0, lambda_=1.0, seed=int(seed),
               chain=int(chain), treatment_col=int(treatment_col), probit_offset=float(offset))
    cfg.update(extra)
    return cfg


def _as_design(X, extra_cols=0, rows=None):
    """n x (p + extra_cols) column-major matrix whose leading columns are X (row-major input: transposed in cache-sized
    blocks, 1.5x a plain assignment; column-major input: a column-wise memcpy).  rows: an index vector -- the matrix of
    X[rows] without the row-major intermediate (gathered block by block, straight into the column-major result)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    p = X.shape[1]
    n = X.shape[0] if rows is None else len(rows)
    xc = np.empty((n, p + extra_cols), dtype=np.float64, order="F")
    if rows is not None:
        for r0 in range(0, n, 8192):
            xc[r0:r0 + 8192, :p] = X[rows[r0:r0 + 8192]]
    elif X.flags.c_contiguous and p > 1:
        for r0 in range(0, n, 8192):
            xc[r0:r0 + 8192, :p] = X[r0:r0 + 8192]
    else:
        xc[:, :p] = X
    return xc


def _fit_design(y01, xc, cuts_list, trt_col, n_samples, n_burn, n_thin, n_chains, seed, first_chain, devices, keep_sampler,
                bart_args):
    """the probit-BART chains of one model on the column-major design xc (treatment in column trt_col, or -1)"""
    n = y01.size
    z = xc[:, trt_col].astype(np.int32) if trt_col >= 0 else np.zeros(n, dtype=np.int32)
    if _lib.device_count() < 1:
        raise _lib.BcfGpuError(2, "no CUDA device available (libbcfgpu has no CPU fallback)")
    cuts, off = prep.pack_cuts(cuts_list)
    cfg = bart_config(seed=seed, chain=first_chain, treatment_col=trt_col, **bart_args)
    devs = list(devices) if devices is not None else list(range(min(_lib.device_count(), max(1, int(n_chains)))))
    samplers = _lib.fit_chains(cfg, int(n_chains), devs, y01, z, xc, None, cuts, off, None, None,
                               int(n_burn), int(n_samples), int(n_thin))
    res, err = [], None
    for i, s in enumerate(samplers):
        try:
            res.append((s.mu_mean(), s.tau_mean(), s.tau_var(), s.num_saved, s.last_run_ms))
        except BaseException as e:
            err = err or e
        finally:
            if not (keep_sampler and i == 0):
                s.close()
    if err is not None:
        if keep_sampler:
            samplers[0].close()
        raise err
    ns = res[0][3]
    ntot = ns * len(res)
    out = {"prob_mean": np.mean([r[0] for r in res], axis=0), "device_ms": [r[4] for r in res]}
    if keep_sampler:
        out["sampler"] = samplers[0]
    if trt_col >= 0:
        m = np.mean([r[1] for r in res], axis=0)
        out["ite_mean"] = m
        out["ite_var"] = (sum((ns - 1) * r[2] + ns * (r[1] - m) ** 2 for r in res) / (ntot - 1)) if ntot > 1 else np.zeros_like(m)
    return out


def probit_bart(y01, X, treatment=None, n_samples=500, n_burn=500, n_thin=1, n_chains=2, seed=0, first_chain=0, devices=None,
                keep_sampler=False, **bart_args):
    """Probit BART of the 0/1 vector y01 on the columns of X (and, S-learner, on `treatment` appended as the last column).
    Returns dict(prob_mean = posterior mean of P(y = 1 | x[, z observed]); with a treatment also ite_mean / ite_var = posterior
    mean / variance of Phi(f(x, 1)) - Phi(f(x, 0)) per row), chains pooled.  keep_sampler: the first chain's _lib.Sampler is
    returned open as out["sampler"] (bcf_iv grows its CART on the covariates it holds binned on the device; the caller closes it)."""
    y01 = np.asarray(y01, dtype=np.float64).ravel()
    if not np.all((y01 == 0) | (y01 == 1)):
        raise ValueError("probit BART needs a 0/1 response")
    if treatment is not None:
        z = np.asarray(treatment).ravel()
        if not np.all((z == 0) | (z == 1)):
            raise ValueError("treatment must be a vector of 0's and 1's")
        xc = _as_design(X, 1)
        xc[:, -1] = z
        trt_col = xc.shape[1] - 1
    else:
        xc = _as_design(X)
        trt_col = -1
    cuts_list = [bart_cutpoints(xc[:, v]) for v in range(xc.shape[1])]
    return _fit_design(y01, xc, cuts_list, trt_col, n_samples, n_burn, n_thin, n_chains, seed, first_chain, devices,
                       keep_sampler, bart_args)


def bartc(response, treatment, confounders, n_samples=500, n_burn=500, n_chains=2, seed=0, devices=None, fit=probit_bart,
          keep_sampler=False, **fit_args):
    """bartCause::bartc(response, treatment, confounders, n.samples, n.burn, n.chains) for a 0/1 response, then
    extract(type = "ite") averaged over the draws: dict(ite_mean, ite_var, p_score, prob_mean).  `fit`: the probit-BART
    fitting function (the parity tests pass the CPU oracle's); fit_args (e.g. max_ctas: a share of the GPU per chain, the
    chains then run side by side) go to it.  keep_sampler: the response model's first sampler comes back open as
    out["sampler"] (its leading columns are the confounders)."""
    if fit is not probit_bart:
        trt = fit(treatment, confounders, None, n_samples=n_samples, n_burn=n_burn, n_chains=n_chains, seed=seed,
                  first_chain=100, devices=devices, **fit_args)
        ps = trt["prob_mean"]
        Xr = np.column_stack([np.asarray(confounders, dtype=np.float64), ps])
        rsp = fit(response, Xr, treatment, n_samples=n_samples, n_burn=n_burn, n_chains=n_chains, seed=seed, first_chain=200,
                  devices=devices, **fit_args)
        return {"ite_mean": rsp["ite_mean"], "ite_var": rsp["ite_var"], "p_score": ps, "prob_mean": rsp["prob_mean"]}
    # One column-major design [confounders | p.score | treatment] serves both models (the treatment model reads its leading
    # columns), and the confounders' cutpoints are made once: at n = 5e5, p = 100 the copies and column scans of two
    # separate set-ups cost more than the chains themselves.
    resp = np.asarray(response, dtype=np.float64).ravel()
    trt01 = np.asarray(treatment, dtype=np.float64).ravel()
    for v, what in ((resp, "probit BART needs a 0/1 response"), (trt01, "treatment must be a vector of 0's and 1's")):
        if not np.all((v == 0) | (v == 1)):
            raise ValueError(what)
    xr = _as_design(confounders, 2)
    p = xr.shape[1] - 2
    cuts_x = [bart_cutpoints(xr[:, v]) for v in range(p)]
    common = dict(n_samples=n_samples, n_burn=n_burn, n_thin=1, n_chains=n_chains, seed=seed, devices=devices)
    trt = _fit_design(trt01, xr[:, :p], cuts_x, -1, first_chain=100, keep_sampler=False, bart_args=fit_args, **common)
    ps = trt["prob_mean"]
    xr[:, p] = ps
    xr[:, p + 1] = trt01
    cuts = cuts_x + [bart_cutpoints(xr[:, p]), bart_cutpoints(xr[:, p + 1])]
    rsp = _fit_design(resp, xr, cuts, p + 1, first_chain=200, keep_sampler=keep_sampler, bart_args=fit_args, **common)
    out = {"ite_mean": rsp["ite_mean"], "ite_var": rsp["ite_var"], "p_score": ps, "prob_mean": rsp["prob_mean"]}
    if keep_sampler:
        out["sampler"] = rsp.get("sampler")
    return out
