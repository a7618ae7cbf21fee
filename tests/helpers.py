"""Shared builders for the parity tests: one synthetic problem, fed identically to the oracle and the GPU path."""
from __future__ import annotations

import numpy as np

from bcf_iv_b200 import datasets, prep


def prepare(*a, **kw):
    """prep.prepare with the default sighat taken from the CPU oracle's lm (oracle/glm_oracle.py): the parity problems are
    prepared without a GPU, and the oracle and the GPU sampler receive the identical prior"""
    from oracle import glm_oracle
    kw.setdefault("sigma_fn", glm_oracle.lm_sigma)
    return prep.prepare(*a, **kw)


class Problem:
    pass


def make_problem(n=500, p=10, seed=1, continuous=False, rho=0.0, compliance=0.75, pihat_noise=0.05):
    """Discovery-sample-shaped problem: y, z (instrument), X twice, pihat appended to the control matrix.
    Everything the reference's bcf() call at R/bcf_iv.R:89 would receive."""
    d = datasets.generate_dataset_wide(n=n, p=p, rho=rho, compliance=compliance, seed=seed, continuous=continuous)
    rng = np.random.default_rng(seed + 1000)
    pihat = pihat_noise * rng.standard_normal(n)          # logit of a randomised instrument is ~0
    P = prepare(d["y"], d["z"], d["X"], d["X"], pihat)
    pr = Problem()
    pr.data, pr.P, pr.n, pr.p = d, P, n, p
    pr.cuts_c_flat, pr.off_c = prep.pack_cuts(P.cuts_c)
    pr.cuts_m_flat, pr.off_m = prep.pack_cuts(P.cuts_m)
    return pr


def oracle_sampler(pr, **kw):
    from oracle import orc
    P = pr.P
    args = dict(lambda_=P.lambda_, sigma_init=P.sigma_init, con_sd=P.con_sd, mod_sd=P.mod_sd)
    args.update(kw)
    perm = P.perm
    return orc.OracleSampler(P.yscale[perm], P.ntrt, P.x_c[perm], P.x_m[perm], P.cuts_c, P.cuts_m, **args)


def gpu_sampler(pr, **kw):
    from bcf_iv_b200 import _lib
    P = pr.P
    cfg = dict(lambda_=P.lambda_, sigma_init=P.sigma_init, con_sd=P.con_sd, mod_sd=P.mod_sd)
    cfg.update(kw)
    s = _lib.Sampler(**cfg)
    s.set_data(P.yscale, P.z, P.x_c, P.x_m, pr.cuts_c_flat, pr.off_c, pr.cuts_m_flat, pr.off_m)
    return s


def tree_signature(heap, var, cut):
    """structure of a serialised tree as a sorted tuple (order-free)"""
    return tuple(sorted((int(h), int(v), int(c)) for h, v, c in zip(heap, var, cut)))


def random_tree(rng, ncut, nleaves, max_depth=12, with_ranges=False):
    """random full binary tree as (heap, var, cut, mu) with valid (var, cut) given ancestors (bcf's rg rule);
    with_ranges: also the usable cutpoint range {var: (lo, hi)} of every leaf (variables not listed: all cutpoints)"""
    p = len(ncut)
    leaves = {1: {}}          # heap id -> {var: (lo, hi)}
    internal = {}
    while len(leaves) < nleaves:
        cand = [h for h in leaves if h.bit_length() - 1 < max_depth]
        if not cand:
            break
        h = cand[rng.integers(len(cand))]
        rngs = leaves[h]
        ok = [v for v in range(p) if rngs.get(v, (0, ncut[v] - 1))[1] >= rngs.get(v, (0, ncut[v] - 1))[0]]
        if not ok:
            break
        v = ok[rng.integers(len(ok))]
        lo, hi = rngs.get(v, (0, ncut[v] - 1))
        c = int(rng.integers(lo, hi + 1))
        del leaves[h]
        internal[h] = (v, c)
        l = dict(rngs); l[v] = (lo, c - 1)
        r = dict(rngs); r[v] = (c + 1, hi)
        leaves[2 * h] = l
        leaves[2 * h + 1] = r
    heap, var, cut, mu = [], [], [], []
    for h, (v, c) in internal.items():
        heap.append(h); var.append(v); cut.append(c); mu.append(0.0)
    for h in leaves:
        heap.append(h); var.append(-1); cut.append(-1); mu.append(float(rng.standard_normal()))
    out = (np.array(heap, np.uint32), np.array(var, np.int32), np.array(cut, np.int32), np.array(mu))
    return out + (leaves,) if with_ranges else out


def oracle_bcf(y, z, x_control, x_moderate=None, pihat=None, nburn=None, nsim=None, nthin=1, random_seed=None,
               first_chain=0, keep_draws=False, max_leaves=0, nthreads=1, **ignored):
    """bcf() with the CPU ORACLE as the sampler: the same R-side preamble (prep.prepare), the same Philox stream
    (seed, chain), the same outputs as bcf_iv_b200.bcf.bcf -- the oracle-driven twin of the end-to-end parity tests."""
    from oracle import orc
    if x_moderate is None:
        x_moderate = x_control
    P = prepare(y, z, x_control, x_moderate, pihat)
    perm = P.perm
    o = orc.OracleSampler(P.yscale[perm], P.ntrt, P.x_c[perm], P.x_m[perm], P.cuts_c, P.cuts_m, lambda_=P.lambda_,
                          sigma_init=P.sigma_init, con_sd=P.con_sd, mod_sd=P.mod_sd, seed=int(random_seed),
                          chain=int(first_chain), max_leaves=max_leaves, nthreads=nthreads)
    out = o.run(int(nburn), int(nsim), int(nthin))
    o.close()
    inv = P.inv_perm
    return {"tau_mean": P.sdy * out["tau_mean"][inv], "mu_mean": P.muy + P.sdy * out["mu_mean"][inv],
            "yhat_mean": P.muy + P.sdy * out["yhat_mean"][inv], "sigma": P.sdy * out["sigma"]}


def oracle_probit_sampler(y01, X, treatment=None, seed=0, chain=0, n_trees=75, k=2.0, **kw):
    """the CPU oracle set up as bcf_iv_b200.bart.probit_bart sets up the GPU sampler (same cutpoints, same Philox stream)"""
    from oracle import orc
    from bcf_iv_b200 import bart
    y01 = np.asarray(y01, dtype=np.float64).ravel()
    X = np.asarray(X, dtype=np.float64)
    n = y01.size
    if treatment is not None:
        z = np.asarray(treatment).ravel().astype(np.int32)
        xc = np.column_stack([X, z.astype(np.float64)])
        trt_col = X.shape[1]
    else:
        z = np.zeros(n, dtype=np.int32)
        xc = X
        trt_col = -1
    cuts = [bart.bart_cutpoints(xc[:, v]) for v in range(xc.shape[1])]
    perm = np.argsort(-z, kind="stable")
    o = orc.OracleSampler(np.zeros(n), int(z.sum()), xc[perm], np.zeros((n, 0)), cuts, [], ntree_con=n_trees, ntree_mod=0,
                          con_sd=3.0 / k, use_muscale=False, use_tauscale=False, sigma_init=1.0, lambda_=1.0, seed=seed,
                          chain=chain, **kw)
    o.probit_enable(y01[perm], perm.astype(np.int32), 0.0, trt_col)
    return o, perm


def oracle_probit_bart(y01, X, treatment=None, n_samples=500, n_burn=500, n_thin=1, n_chains=2, seed=0, first_chain=0,
                       devices=None, **bart_args):
    """bart.probit_bart with the CPU oracle as the sampler (chains pooled the same way)"""
    res = []
    for c in range(n_chains):
        o, perm = oracle_probit_sampler(y01, X, treatment, seed=seed, chain=first_chain + c, max_leaves=128, **bart_args)
        out = o.probit_run(n_burn, n_samples, n_thin)
        o.close()
        inv = np.empty(perm.size, np.int64); inv[perm] = np.arange(perm.size)
        res.append((out["prob_mean"][inv], out["ite_mean"][inv]))
    r = {"prob_mean": np.mean([a for a, _ in res], axis=0)}
    if treatment is not None:
        r["ite_mean"] = np.mean([b for _, b in res], axis=0)
        r["ite_var"] = np.zeros_like(r["ite_mean"])
    return r
