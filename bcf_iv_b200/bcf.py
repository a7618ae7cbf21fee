"""`bcf()` -- the drop-in for the one call BayesIV makes into the hot path.

Reference call sites (fbargaglistoffi/BCF-IV):
    R/bcf_iv.R:89    itt_bcf <- quiet(bcf(y[-index], z[-index], x[-index,], x[-index,], pihat, nburn=n_burn, nsim=n_sim))
    R/bcf_itt.R:71   bcf_fit <- quiet(bcf(... same ...))
and the only use of the result: `$tau` -> colMeans (R/bcf_iv.R:90-91, R/bcf_itt.R:74-75).

Same formal argument names, meanings, defaults and error behaviour as `bcf::bcf` [bcf-recall, SURVEY.md A.1 / 8b]; the
sampler itself is libbcfgpu.so (CUDA, sm_100a) behind the C-ABI of include/bcfgpu.h.  There is no CPU path: without a
CUDA device the call raises.  The R twin of this file is bcf_iv_b200/R/BayesIVgpu/R/bcf.R.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib, prep

DRAW_BYTES_CAP = 8 << 30   # keep nsim x n draws on the device only below this (3 matrices of doubles)


def _collect(s, keep, pred=None, close=True):
    """what bcf() returns from one finished chain (a _lib.Sampler); closes it unless the caller keeps the handle"""
    try:
        r = {"sigma": s.sigma(), "scales": s.scales(), "tau_mean": s.tau_mean(), "mu_mean": s.mu_mean(),
             "yhat_mean": s.yhat_mean(), "tau_var": s.tau_var(), "ms": s.last_run_ms, "moves": s.counters()}
        if keep:
            r["tau"], r["mu"], r["yhat"] = s.draws("tau"), s.draws("mu"), s.draws("yhat")
        if pred is not None:
            r["pred"] = s.predict(pred[0], pred[1], draws=pred[2])
        return r
    except BaseException:
        close = True
        raise
    finally:
        if close:
            s.close()


def bcf(y, z, x_control, x_moderate=None, pihat=None, nburn=None, nsim=None, nthin=1, update_interval=100,
        ntree_control=200, sd_control=None, base_control=0.95, power_control=2.0,
        ntree_moderate=50, sd_moderate=None, base_moderate=0.25, power_moderate=3.0,
        nu=3.0, lambda_=None, sigq=0.9, sighat=None, include_pi="control", use_muscale=True, use_tauscale=True,
        w=None, random_seed=None, n_chains=1, n_cores=None, n_threads=None, no_output=True,
        save_tree_directory=None, log_file=None, verbose=False,
        devices=None, keep_draws=None, engine=_lib.ENGINE_AUTO, max_ctas=0, first_chain=0,
        x_predict_control=None, x_predict_moderate=None, pihat_predict=None, predict_draws=False, keep_sampler=False,
        x_spare_column=False):
    """Fit the Bayesian Causal Forest  y = mu(x, pihat) + tau(x) z + eps  on the GPU.

    Returns a dict with bcf's fields: `sigma` (nsim*n_chains), `yhat`, `mu`, `tau` (nsim*n_chains x n, original row
    order, original y units), `mu_scale`, `tau_scale`, `perm`; plus `tau_mean` / `mu_mean` / `yhat_mean` / `tau_var`
    (what BayesIV actually consumes: colMeans) which are accumulated on the device and always present.

    Extras beyond bcf's signature: `devices` (CUDA ordinals, chains are dealt round-robin: one independent chain per
    GPU), `keep_draws` (None = keep the three nsim x n matrices only while they fit DRAW_BYTES_CAP),
    `engine`, `max_ctas` (0 = the whole GPU; k > 0 = at most k CTAs, so that several fits can run side by side on one GPU),
    `first_chain` (Philox stream of the first chain; chain c uses first_chain + c);
    `x_predict_control` / `x_predict_moderate` / `pihat_predict` (bcf 2.x's predict() in the same call): new covariate rows at
    which every saved sweep's forests are evaluated on the device -> `tau_predict_mean`, `tau_predict_var`,
    `mu_predict_mean` (and `tau_predict`, `mu_predict`: nsim*n_chains x n_new with predict_draws=True), original y units;
    `keep_sampler` (the first chain's _lib.Sampler is returned open as out["sampler"]: bcf_iv grows its CART on the binned
    covariates that are already on the device; the caller closes it); `x_spare_column` (the caller states that x_control --
    passed for x_moderate as well -- is the leading columns of a column-major buffer of its own with one spare trailing
    column, which bcf may overwrite with pihat: saves the n x p copy into [x | pihat]).  bcf 2.x arguments that have no meaning here (n_cores, n_threads, save_tree_directory, log_file,
    no_output, verbose, update_interval) are accepted and ignored; observation weights `w` are not implemented.
    """
    if nburn is None or nsim is None:
        raise TypeError("bcf() needs nburn and nsim")
    if w is not None:
        raise NotImplementedError("observation weights (bcf 2.x `w`) are not implemented; BayesIV never passes them")
    if x_moderate is None:
        x_moderate = x_control
    if pihat is None:
        raise TypeError("bcf() needs pihat")
    if int(nburn) < 0 or int(nsim) < 1 or int(nthin) < 1:
        raise ValueError("need nburn >= 0, nsim >= 1, nthin >= 1")
    if nburn < 100:
        warnings.warn("A low (<100) value for nburn was supplied")
    if _lib.device_count() < 1:
        raise _lib.BcfGpuError(2, "no CUDA device available (libbcfgpu has no CPU fallback)")
    P = prep.prepare(y, z, x_control, x_moderate, pihat, include_pi=include_pi, sd_control=sd_control,
                     sd_moderate=sd_moderate, nu=nu, lambda_=lambda_, sigq=sigq, sighat=sighat,
                     use_tauscale=use_tauscale, spare_column=bool(x_spare_column))
    cc, oc = prep.pack_cuts(P.cuts_c)
    cm, om = prep.pack_cuts(P.cuts_m)
    n = P.n
    if keep_draws is None:
        keep_draws = 3 * 8 * n * int(nsim) <= DRAW_BYTES_CAP
    seed = int(random_seed) if random_seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
    devs = list(devices) if devices is not None else list(range(min(_lib.device_count(), max(1, int(n_chains)))))
    cfg = dict(ntree_con=int(ntree_control), ntree_mod=int(ntree_moderate), alpha_con=float(base_control),
               beta_con=float(power_control), alpha_mod=float(base_moderate), beta_mod=float(power_moderate),
               con_sd=P.con_sd, mod_sd=P.mod_sd, nu=float(nu), lambda_=P.lambda_, sigma_init=P.sigma_init,
               use_muscale=int(bool(use_muscale)), use_tauscale=int(bool(use_tauscale)), seed=seed,
               keep_draws=7 if keep_draws else 0, engine=int(engine), max_ctas=int(max_ctas))
    # chains that share a GPU: side by side on a share of its SMs each while an SM's 27 tiles of 256 rows hold the share's
    # observations, else one after the other (each launch then fills the device).  All chains run inside ONE native call
    # (bcfgpu_fit_chains: one host thread per device, or per chain when they run side by side).
    per_dev = -(-int(n_chains) // len(devs))
    share = _lib.B200_SMS // per_dev if per_dev > 1 else 0
    if per_dev > 1 and int(max_ctas) == 0 and share >= 1 and n <= 27 * 256 * share - 512:
        cfg["max_ctas"] = share
    cfg["chain"] = int(first_chain)
    pred = None
    if x_predict_control is not None:
        xpc = np.asarray(x_predict_control, dtype=np.float64)
        xpm = np.asarray(x_predict_moderate if x_predict_moderate is not None else x_predict_control, dtype=np.float64)
        php = np.asarray(pihat_predict, dtype=np.float64).reshape(xpc.shape[0], -1) if pihat_predict is not None else None
        if include_pi != "none" and php is None:
            raise TypeError("predicting needs pihat_predict (include_pi != 'none')")
        xc_new = np.column_stack([xpc, php]) if include_pi in ("control", "both") else xpc
        xm_new = np.column_stack([xpm, php]) if include_pi in ("moderate", "both") else xpm
        pred = (xc_new, xm_new, bool(predict_draws))
        cfg["keep_draws"] |= 8                                # keep every saved sweep's forests on the device
    samplers = _lib.fit_chains(cfg, int(n_chains), devs, P.yscale, P.z, P.x_c, P.x_m, cc, oc, cm, om, int(nburn), int(nsim),
                               int(nthin))
    res, err = [], None
    for k, smp in enumerate(samplers):                         # every handle is closed even if one read-back fails
        try:
            res.append(_collect(smp, keep_draws, pred, close=not (keep_sampler and k == 0)))
        except BaseException as e:
            err = err or e
    if err is not None:
        if keep_sampler:
            samplers[0].close()
        raise err
    ns = len(res[0]["sigma"])
    ntot = ns * len(res)
    tau_mean = np.mean([r["tau_mean"] for r in res], axis=0)
    # variance of the stacked draws (what out["tau"] holds): within-chain sums of squares + spread of the chain means
    tau_var = (sum((ns - 1) * r["tau_var"] + ns * (r["tau_mean"] - tau_mean) ** 2 for r in res) / (ntot - 1)
               if ntot > 1 else np.zeros_like(tau_mean))
    sdy, muy = P.sdy, P.muy
    out = {
        "sigma": sdy * np.concatenate([r["sigma"] for r in res]),
        "mu_scale": sdy * np.concatenate([r["scales"][0] for r in res]),
        "tau_scale": sdy * np.concatenate([r["scales"][1] for r in res]),
        "perm": P.perm + 1,                                   # R's 1-based order(z, decreasing = TRUE)
        "tau_mean": sdy * tau_mean,
        "mu_mean": muy + sdy * np.mean([r["mu_mean"] for r in res], axis=0),
        "yhat_mean": muy + sdy * np.mean([r["yhat_mean"] for r in res], axis=0),
        "tau_var": sdy ** 2 * tau_var,
        "device_ms": [r["ms"] for r in res],
        "moves": [r["moves"] for r in res],
    }
    if keep_sampler:
        out["sampler"] = samplers[0]                          # the first chain's live handle (binned X still on the device): close() it
    if pred is not None:
        pm = np.mean([r["pred"]["tau_mean"] for r in res], axis=0)
        out["tau_predict_mean"] = sdy * pm
        out["mu_predict_mean"] = muy + sdy * np.mean([r["pred"]["mu_mean"] for r in res], axis=0)
        out["tau_predict_var"] = sdy ** 2 * (sum((ns - 1) * r["pred"]["tau_var"] + ns * (r["pred"]["tau_mean"] - pm) ** 2
                                                 for r in res) / (ntot - 1) if ntot > 1 else np.zeros_like(pm))
        if predict_draws:
            out["tau_predict"] = sdy * np.concatenate([r["pred"]["tau"] for r in res], axis=0)
            out["mu_predict"] = muy + sdy * np.concatenate([r["pred"]["mu"] for r in res], axis=0)
    if keep_draws:                                            # chains rbind-ed, as bcf 2.x does
        out["tau"] = sdy * np.concatenate([r["tau"] for r in res], axis=0)
        out["mu"] = muy + sdy * np.concatenate([r["mu"] for r in res], axis=0)
        out["yhat"] = muy + sdy * np.concatenate([r["yhat"] for r in res], axis=0)
    return out
