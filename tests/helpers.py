"""Shared builders for the parity tests: one synthetic problem, fed identically to the oracle and the GPU path."""
from __future__ import annotations

import numpy as np

from bcf_iv_b200 import datasets, prep


class Problem:
    pass


def make_problem(n=500, p=10, seed=1, continuous=False, rho=0.0, compliance=0.75, pihat_noise=0.05):
    """Discovery-sample-shaped problem: y, z (instrument), X twice, pihat appended to the control matrix.
    Everything the reference's bcf() call at R/bcf_iv.R:89 would receive."""
    d = datasets.generate_dataset_wide(n=n, p=p, rho=rho, compliance=compliance, seed=seed, continuous=continuous)
    rng = np.random.default_rng(seed + 1000)
    pihat = pihat_noise * rng.standard_normal(n)          # logit of a randomised instrument is ~0
    P = prep.prepare(d["y"], d["z"], d["X"], d["X"], pihat)
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
