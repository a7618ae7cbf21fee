"""`bcf_iv()` / `bcf_itt()` -- Python twins of the reference's two public functions, driving the GPU sampler.

Reference: R/bcf_iv.R:47-304 and R/bcf_itt.R:33-254 (fbargaglistoffi/BCF-IV).  Same arguments, defaults and returned
table; the steps are the reference's: honest split -> logistic propensity (on the LOGIT scale, R/bcf_iv.R:75) ->
complier proportion -> ITT by Bayesian Causal Forest (THE HOT PATH, R/bcf_iv.R:89, here bcf_iv_b200.bcf.bcf on the GPU)
-> tauhat = itt / pic -> depth-limited CART on tauhat -> per-node 2SLS and weak-instrument test on the inference
sample -> p-value adjustment on the leaves.  The R version of the same glue is bcf_iv_b200/R/BayesIVgpu/R/.

Everything except the samplers stays on the CPU (north_star): the pieces R takes from rpart / AER / stats are restated
here in closed form (`cart`, `tsls`, `p_adjust`); the logistic propensity model (`logit_link`) runs on the device.  The reference's second sampler, bartCause::bartc
(probit BART S-learner: the complier proportion of every call, R/bcf_iv.R:78-80, and the whole model when binary = TRUE,
R/bcf_iv.R:198-202, R/bcf_itt.R:166-170), is bcf_iv_b200.bart.bartc: the same library's probit-BART model on the GPU.
Deviations:
  * `pic_model="bcf"` (not the default) keeps round 1's stand-in for bartc: the complier proportion as the tau of a
    linear-probability BCF of w on z, and with binary = TRUE the ITT as the tau of a BCF on the 0/1 outcome;
  * the honest split uses NumPy's generator: the same seed gives a different split than R's sample().
"""
from __future__ import annotations

import math

import numpy as np

from . import bart as _bart
from . import bcf as _bcf

COLUMNS = ["node", "CCACE", "pvalue", "Weak_IV_test", "Pi_obs", "ITT", "Pi_compliers", "Adj_pvalue"]


# ---------------------------------------------------------------------------------------------------------------
# stats::glm(z ~ x, binomial) + predict(type = "link")   (R/bcf_iv.R:72-75)
# ---------------------------------------------------------------------------------------------------------------
def logit_link(z, X, device=0):
    """glm(z ~ X, family = binomial) followed by predict(): the linear predictor (log-odds), predict.glm's default scale.
    IRLS on the device (bcfgpu_glm_logit: weighted Gram matrix + streaming passes in CUDA); no host path."""
    return _bcf._lib.glm_logit(z, X, device=device)


# ---------------------------------------------------------------------------------------------------------------
# rpart::rpart(tauhat ~ ., method = "anova", maxdepth, cp, minsplit)   (R/bcf_iv.R:101-105, R/bcf_itt.R:82-84)
# ---------------------------------------------------------------------------------------------------------------
class CartNode:
    __slots__ = ("id", "n", "mean", "dev", "var", "cut", "left", "right", "rule")

    def __init__(self, id_, idx, y, rule):
        self.id, self.n, self.mean = id_, len(idx), float(y[idx].mean())
        self.dev = float(((y[idx] - self.mean) ** 2).sum())
        self.var = self.cut = self.left = self.right = None
        self.rule = rule                                    # list of (var, "<" or ">=", cut): path.rpart's conditions


def _best_split(X, y, idx, minbucket):
    best = (0.0, None, None)
    yy = y[idx]
    tot, n = yy.sum(), len(idx)
    base = tot * tot / n
    for v in range(X.shape[1]):
        xv = X[idx, v]
        lo, hi = xv.min(), xv.max()
        if lo == hi:
            continue
        is_lo = xv == lo
        if np.all(is_lo | (xv == hi)):          # two-valued column (the reference's covariates are binary): no sort needed
            nl1 = int(is_lo.sum())
            if nl1 < minbucket or n - nl1 < minbucket:
                continue
            c1 = float(yy[is_lo].sum())
            g1 = c1 * c1 / nl1 + (tot - c1) ** 2 / (n - nl1) - base
            if g1 > best[0] * (1 + 1e-12):
                best = (float(g1), v, 0.5 * (lo + hi))
            continue
        order = np.argsort(xv, kind="stable")
        xs, ys = xv[order], yy[order]
        cs = np.cumsum(ys)[:-1]
        nl = np.arange(1, n)
        ok = (xs[1:] != xs[:-1]) & (nl >= minbucket) & (n - nl >= minbucket)
        if not ok.any():
            continue
        gain = cs ** 2 / nl + (tot - cs) ** 2 / (n - nl) - base      # decrease of the sum of squares
        gain = np.where(ok, gain, -np.inf)
        j = int(np.argmax(gain))
        if gain[j] > best[0] * (1 + 1e-12):
            best = (float(gain[j]), v, 0.5 * (xs[j] + xs[j + 1]))
    return best


def cart(X, y, maxdepth=2, cp=0.01, minsplit=20, minbucket=None):
    """Regression tree with rpart's stopping rules: a node is split only if it has >= minsplit rows, each child gets
    >= minbucket = round(minsplit / 3), and the split improves the fit by at least cp * (root deviance).
    Nodes are numbered like rpart's frame: root 1, children 2i and 2i + 1; rows with x < cut go LEFT."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if minbucket is None:
        minbucket = int(round(minsplit / 3.0))
    root = CartNode(1, np.arange(len(y)), y, [])
    nodes, where = {1: root}, np.ones(len(y), dtype=np.int64)
    todo = [(root, np.arange(len(y)), 0)]
    while todo:
        nd, idx, depth = todo.pop()
        if depth >= maxdepth or nd.n < minsplit or nd.dev <= 0:
            continue
        gain, v, cut = _best_split(X, y, idx, minbucket)
        if v is None or gain < cp * root.dev:
            continue
        li, ri = idx[X[idx, v] < cut], idx[X[idx, v] >= cut]
        nd.var, nd.cut = v, cut
        nd.left = CartNode(2 * nd.id, li, y, nd.rule + [(v, "<", cut)])
        nd.right = CartNode(2 * nd.id + 1, ri, y, nd.rule + [(v, ">=", cut)])
        nodes[nd.left.id], nodes[nd.right.id] = nd.left, nd.right
        where[li], where[ri] = nd.left.id, nd.right.id
        todo += [(nd.left, li, depth + 1), (nd.right, ri, depth + 1)]
    return nodes, where


def cart_device(sampler, X, y, maxdepth=2, cp=0.01, minsplit=20, minbucket=None, forest=1):
    """`cart` with the split search on the GPU: the candidate statistics of EVERY (node, variable, cutpoint)
    of a tree level come from one pass of the all-candidate histogram kernel (K2', bcfgpu_op_hist_all) over the binned
    covariates the sampler already holds on the device; prefix sums over the bins give each split's left child, the
    arg-max and the rpart stopping rules are a few hundred numbers on the host.  Same result as `cart` when every column's
    candidate splits are the sampler's cutpoints, i.e. for columns with fewer than 8 distinct values (midpoints between
    them: rpart's own candidates) -- the reference's designs; the caller falls back to `cart` otherwise."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    if minbucket is None:
        minbucket = int(round(minsplit / 3.0))
    off = sampler.off_mod if forest == 1 else sampler.off_con
    p = X.shape[1]
    ncut = np.diff(off)[:p]
    cuts_all = sampler.cutpoints(forest)
    root = CartNode(1, np.arange(n), y, [])
    nodes, where = {1: root}, np.ones(n, dtype=np.int64)
    frontier = [root]
    for depth in range(maxdepth):
        live = [nd for nd in frontier if nd.n >= minsplit and nd.dev > 0]
        if not live:
            break
        lab = np.full(n, -1, dtype=np.int32)
        for k, nd in enumerate(live):
            lab[where == nd.id] = k
        cnt, sm, _ = sampler.op_hist_all(forest, lab, len(live), y)
        nxt = []
        for k, nd in enumerate(live):
            tot, base = nd.mean * nd.n, (nd.mean * nd.n) ** 2 / nd.n
            best = (0.0, None, None)
            for v in range(p):
                b0 = int(off[v]) + v
                c_ = np.cumsum(cnt[k, b0:b0 + ncut[v]])            # rows with bin <= c  <=>  x < cut[c]: the left child of cut c
                s_ = np.cumsum(sm[k, b0:b0 + ncut[v]])
                ok = (c_ >= minbucket) & (nd.n - c_ >= minbucket)
                if not ok.any():
                    continue
                with np.errstate(divide="ignore", invalid="ignore"):
                    gain = np.where(ok, s_ ** 2 / c_ + (tot - s_) ** 2 / (nd.n - c_) - base, -np.inf)
                j = int(np.argmax(gain))
                if gain[j] > best[0] * (1 + 1e-12):
                    best = (float(gain[j]), v, float(cuts_all[int(off[v]) + j]))
            gain, v, cut = best
            if v is None or gain < cp * root.dev:
                continue
            idx = np.nonzero(where == nd.id)[0]
            li, ri = idx[X[idx, v] < cut], idx[X[idx, v] >= cut]
            nd.var, nd.cut = v, cut
            nd.left = CartNode(2 * nd.id, li, y, nd.rule + [(v, "<", cut)])
            nd.right = CartNode(2 * nd.id + 1, ri, y, nd.rule + [(v, ">=", cut)])
            nodes[nd.left.id], nodes[nd.right.id] = nd.left, nd.right
            where[li], where[ri] = nd.left.id, nd.right.id
            nxt += [nd.left, nd.right]
        frontier = nxt
    return nodes, where


def rule_text(rule, names):
    """the node's path as the reference prints it: 'x1< 0.5 & x2>=0.5' (path.rpart's format)"""
    return " & ".join(f"{names[v]}{'< ' if op == '<' else '>='}{cut:.4g}" for v, op, cut in rule)


def prep_cat_levels():
    from . import prep
    return prep.CP_CAT_LEVELS


class _Rows:
    """X[rows] as far as the node rules need it: one column at a time (the inference sample is only ever looked at through
    the few split variables, so its n x p row gather is never made)"""

    def __init__(self, X, rows):
        self.X, self.rows, self.shape, self._col = X, rows, (len(rows), X.shape[1]), {}

    def __getitem__(self, key):
        sl, v = key
        if not (isinstance(sl, slice) and sl == slice(None)):
            raise IndexError("_Rows serves whole columns only")
        if v not in self._col:
            self._col[v] = self.X[self.rows, v]
        return self._col[v]


def rule_mask(rule, X):
    m = np.ones(X.shape[0], dtype=bool)
    for v, op, cut in rule:
        m &= (X[:, v] < cut) if op == "<" else (X[:, v] >= cut)
    return m


# ---------------------------------------------------------------------------------------------------------------
# AER::ivreg(y ~ w | z) + summary(diagnostics = TRUE)   (R/bcf_iv.R:127-132, 158-163)
# ---------------------------------------------------------------------------------------------------------------
def _t_sf2(t, df):
    from scipy.special import stdtr             # (scipy.stats takes two seconds to import)
    return float(2.0 * stdtr(df, -abs(t)))


def tsls(y, w, z):
    """Just-identified 2SLS of y on (1, w) with instrument z: coefficient of w, its t-test p-value (n - 2 df, as
    summary.ivreg), and the weak-instrument test (first-stage F of z, on (1, n - 2) df)."""
    y, w, z = (np.asarray(a, dtype=np.float64) for a in (y, w, z))
    n = len(y)
    if n <= 2:                                               # no residual degrees of freedom
        return float("nan"), float("nan"), float("nan")
    zc, wc, yc = z - z.mean(), w - w.mean(), y - y.mean()
    szw, szz = float(zc @ wc), float(zc @ zc)
    if szz == 0 or szw == 0:
        return float("nan"), float("nan"), float("nan")
    beta = float(zc @ yc) / szw
    alpha = y.mean() - beta * w.mean()
    res = y - alpha - beta * w
    s2 = float(res @ res) / (n - 2)
    se = math.sqrt(s2 * szz) / abs(szw)                       # sqrt(s2 / (sum (what - mean)^2)), what = fitted first stage
    p = _t_sf2(beta / se, n - 2) if se > 0 else float("nan")
    # first stage w ~ z
    g = szw / szz
    r1 = wc - g * zc
    s2f = float(r1 @ r1) / (n - 2)
    if s2f <= 0:
        pweak = 0.0
    else:
        from scipy.special import fdtrc
        pweak = float(fdtrc(1, n - 2, g * g * szz / s2f))
    return beta, p, pweak


# ---------------------------------------------------------------------------------------------------------------
# stats::p.adjust   (R/bcf_iv.R:183)
# ---------------------------------------------------------------------------------------------------------------
def p_adjust(p, method="holm"):
    p = np.asarray(p, dtype=np.float64)
    ok = np.isfinite(p)
    if not ok.all():                                         # R: NAs are left out, n = number of non-NA p-values
        out = np.full(len(p), np.nan)
        out[ok] = p_adjust(p[ok], method)
        return out
    n = len(p)
    if n == 0 or method == "none":
        return p.copy()
    o = np.argsort(p)
    ro = np.argsort(o)
    ps = p[o]
    i = np.arange(1, n + 1)
    if method == "bonferroni":
        return np.minimum(1.0, p * n)
    if method == "holm":
        return np.minimum(1.0, np.maximum.accumulate((n - i + 1) * ps))[ro]
    if method in ("hochberg", "hockberg"):                   # the reference's docstring spells it "hockberg"
        return np.minimum(1.0, np.minimum.accumulate(((n - i + 1) * ps)[::-1])[::-1])[ro]
    if method in ("BH", "fdr"):
        return np.minimum(1.0, np.minimum.accumulate((n / i * ps)[::-1])[::-1])[ro]
    if method == "BY":
        q = np.sum(1.0 / i)
        return np.minimum(1.0, np.minimum.accumulate((q * n / i * ps)[::-1])[::-1])[ro]
    if method == "hommel":
        q = pa = np.full(n, np.min(n * ps / i))
        for m in range(n - 1, 1, -1):
            i1 = np.arange(n - m + 1)
            i2 = np.arange(n - m + 1, n)
            q1 = np.min(m * ps[i2] / np.arange(2, m + 1))
            q = q.copy()
            q[i1] = np.minimum(m * ps[i1], q1)
            q[i2] = q[n - m]
            pa = np.maximum(pa, q)
        return np.maximum(pa, ps)[ro]
    raise ValueError(f"unknown p-value adjustment method {method!r}")


# ---------------------------------------------------------------------------------------------------------------
# the two public functions
# ---------------------------------------------------------------------------------------------------------------
def _split(n, inference_ratio, seed):
    rng = np.random.default_rng(seed)
    index = rng.choice(n, int(n * inference_ratio), replace=False)     # the INFERENCE rows (R/bcf_iv.R:57)
    disc = np.setdiff1d(np.arange(n), index)
    return disc, index


def _node_rows(nodes, where, names, inf_y, inf_w, inf_z, inf_X):
    """per-node 2SLS on the inference sample; rows indexed like the reference's bcfivMat (by rpart node number)"""
    ids = sorted(nodes)
    rows = {}
    for nid in ids:
        nd = nodes[nid]
        m = rule_mask(nd.rule, inf_X)
        ys, ws, zs = inf_y[m], inf_w[m], inf_z[m]
        if nid != 1 and len(np.unique(ws)) == 1 and len(np.unique(zs)) == 1:
            rows[nid] = None                                  # R/bcf_iv.R:157: node skipped, row stays NA
            continue
        eff, p, pweak = tsls(ys, ws, zs)
        comp = float(np.mean(zs == ws))
        rows[nid] = dict(node=(None if nid == 1 else rule_text(nd.rule, names)), CCACE=round(eff, 4), pvalue=round(p, 4),
                         Weak_IV_test=round(pweak, 4), Pi_obs=round(m.mean(), 4), ITT=round(eff * comp, 4),
                         Pi_compliers=round(comp, 4))
    return ids, rows


def bcf_iv(y, w, z, x, binary=False, n_burn=500, n_sim=500, inference_ratio=0.5, max_depth=2, cp=0.01, minsplit=10,
           adj_method="holm", seed=42, bcf_fn=None, bartc_fn=None, logit_fn=None, pic_model="bartc", details=None, **bcf_args):
    """Discovery and estimation of heterogeneity in the complier average causal effect (R/bcf_iv.R:47).
    Returns a pandas DataFrame with the reference's columns (COLUMNS), one row per CART node.
    `bcf_fn` / `bartc_fn` / `logit_fn`: the fitting functions (default: bcf_iv_b200.bcf.bcf, bcf_iv_b200.bart.bartc and
    logit_link, all on the GPU) -- the parity tests pass twins that drive the CPU oracle through the identical pipeline.
    `pic_model`: "bartc" (the reference: probit BART for the complier proportion and, with binary = TRUE, for the ITT) or
    "bcf" (both from the BCF sampler on the 0/1 vectors: a linear-probability stand-in).  `details`: a dict that receives the intermediate estimands
    (discovery rows, pic, itt, tauhat, the CART's split variables)."""
    import pandas as pd
    y, w, z = (np.asarray(a, dtype=np.float64).ravel() for a in (y, w, z))
    if binary and not np.all((y == 0) | (y == 1)):
        raise ValueError("binary = TRUE needs a 0/1 outcome")
    X = np.asarray(x, dtype=np.float64)
    names = list(getattr(x, "columns", [f"x{j + 1}" for j in range(X.shape[1])]))
    disc, index = _split(len(y), inference_ratio, seed)
    # the same matrix twice, as the reference passes it: copied to the GPU once; column-major from here on (what every native
    # entry streams): ONE pass over the discovery rows per call, gathered straight into a buffer with a spare trailing column
    # where bcf() puts pihat (its design [x | pihat] then needs no copy of its own)
    Xd = _bart._as_design(X, 1, rows=disc)[:, :X.shape[1]]
    pihat = (logit_fn or logit_link)(z[disc], Xd)                                          # R/bcf_iv.R:72-75
    fit_args = dict(nburn=n_burn, nsim=n_sim, random_seed=seed, keep_draws=False)
    fit_args.update(bcf_args)
    # the two fits are independent: run them side by side.  While a discovery sample fits in half of a B200's shared
    # memory (27 tiles of 256 rows on each of 74 SMs) each fit gets half of the SMs; at any size one fit's host-side
    # preamble hides under the other's sweeps.
    from concurrent.futures import ThreadPoolExecutor
    # At these sizes a sweep is bound by the latency of its decision chain, not by the SMs it has: the fits run side by side
    # on shares of the GPU while an SM's 27 tiles of 256 rows hold a share's observations -- the ITT forest on half of the
    # SMs, bartc's two chains on a quarter each (or on the other half, one after the other, when a quarter is too small).
    sms = _bcf._lib.B200_SMS
    fits_in = lambda ctas: len(disc) <= 27 * 256 * ctas - 512
    bart_ctas = 0
    if "max_ctas" not in fit_args and fits_in(sms // 2):
        fit_args["max_ctas"] = sms // 2
        bart_ctas = sms // 4 if fits_in(sms // 4) else sms // 2
    fit = bcf_fn or _bcf.bcf
    fit_bartc = bartc_fn or _bart.bartc
    live = []                                       # the ITT fit's sampler, kept open for the CART on the device
    if pic_model not in ("bartc", "bcf"):
        raise ValueError("pic_model must be 'bartc' or 'bcf'")
    with ThreadPoolExecutor(max_workers=2) as pool:
        # the reference's two samplers are independent (bartc / bcf): each fit has its own Philox streams, so itt and pic
        # do not share their Monte Carlo noise
        if pic_model == "bartc":
            bkw = dict(max_ctas=bart_ctas) if (bartc_fn is None and bart_ctas) else {}
            f_pic = pool.submit(lambda: fit_bartc(w[disc], z[disc], Xd, n_samples=n_sim, n_burn=n_burn, n_chains=2,
                                                  seed=seed, **bkw)["ite_mean"])             # R/bcf_iv.R:78-80
            if binary:                                                                          # R/bcf_iv.R:198-202
                def _itt_b():
                    r = fit_bartc(y[disc], z[disc], Xd, n_samples=n_sim, n_burn=n_burn, n_chains=2, seed=seed + 1,
                                  **dict(bkw, **({"keep_sampler": True} if bartc_fn is None else {})))
                    live.append(r.get("sampler"))
                    return r["ite_mean"]
                f_itt = pool.submit(_itt_b)
            else:
                def _itt():
                    own = dict(keep_sampler=True, x_spare_column=True) if bcf_fn is None else dict(keep_sampler=False)
                    r = fit(y[disc], z[disc], Xd, Xd, pihat, **dict(fit_args, **own))              # R/bcf_iv.R:89-91
                    live.append(r.get("sampler"))
                    return r["tau_mean"]
                f_itt = pool.submit(_itt)
        else:
            f_pic = pool.submit(lambda: fit(w[disc], z[disc], Xd, Xd, pihat, **dict(fit_args, first_chain=1))["tau_mean"])
            f_itt = pool.submit(lambda: fit(y[disc], z[disc], Xd, Xd, pihat, **fit_args)["tau_mean"])
        pic, itt = f_pic.result(), f_itt.result()
    tauhat = itt / pic                                                                      # R/bcf_iv.R:94
    smp = live[0] if live and live[0] is not None else None
    try:
        # the CART on tauhat: on the device when the sampler's cutpoints ARE rpart's candidate splits (columns with fewer
        # than 8 distinct values) and tauhat fits the kernel's fixed-point range; on the host otherwise
        # (binary = TRUE: the sampler is bartc's response model, one forest whose leading columns are the covariates)
        cforest = 0 if binary else 1
        on_device = (smp is not None
                     and np.all(np.diff(smp.off_con if binary else smp.off_mod)[:Xd.shape[1]] < prep_cat_levels() - 1)
                     and np.all(np.isfinite(tauhat)) and np.max(np.abs(tauhat)) < 16384.0)
        if on_device:
            nodes, where = cart_device(smp, Xd, tauhat, maxdepth=max_depth, cp=cp, minsplit=minsplit, forest=cforest)
        else:
            nodes, where = cart(Xd, tauhat, maxdepth=max_depth, cp=cp, minsplit=minsplit)
    finally:
        if smp is not None:
            smp.close()
    if details is not None:
        details.update(disc=disc, index=index, pic=pic, itt=itt, tauhat=tauhat, cart_on_device=bool(on_device),
                       split_vars=sorted({nd.var for nd in nodes.values() if nd.var is not None}))
    ids, rows = _node_rows(nodes, where, names, y[index], w[index], z[index], _Rows(X, index))
    leaves = {nid for nid in ids if nodes[nid].left is None}
    # rows are indexed by rpart NODE NUMBER as in the reference (bcfivMat[i, ], R/bcf_iv.R:174): the table starts with
    # one row per node and grows, NA rows in between, when a node number exceeds that count
    nrow = max(len(ids), max(nid for nid in ids if rows[nid] is not None or nid in leaves))
    index = list(range(1, nrow + 1))
    blank = {c: None for c in COLUMNS[:-1]}
    tab = pd.DataFrame([rows[n] if rows.get(n) is not None else blank for n in index], index=index, columns=COLUMNS[:-1])
    adj = pd.Series([None] * nrow, index=index, dtype=object)
    lv = [n for n in ids if n in leaves and rows[n] is not None]
    if lv:
        a = np.round(p_adjust([rows[n]["pvalue"] for n in lv], adj_method), 5)
        for n_, v in zip(lv, a):
            adj[n_] = None if not np.isfinite(v) else float(v)
    tab["Adj_pvalue"] = adj
    return tab


def bcf_itt(y, w, z, x, binary=False, max_depth=2, n_burn=500, n_sim=500, inference_ratio=0.5, seed=42, bcf_fn=None,
            bartc_fn=None, logit_fn=None, **bcf_args):
    """Heterogeneity in the intention-to-treat effect (R/bcf_itt.R:33; same positional order of the arguments): prints the
    node effects and returns None, as the reference does."""
    y, w, z = (np.asarray(a, dtype=np.float64).ravel() for a in (y, w, z))
    if binary and not np.all((y == 0) | (y == 1)):
        raise ValueError("binary = TRUE needs a 0/1 outcome")
    X = np.asarray(x, dtype=np.float64)
    names = list(getattr(x, "columns", [f"x{j + 1}" for j in range(X.shape[1])]))
    disc, index = _split(len(y), inference_ratio, seed)
    Xd = _bart._as_design(X, 1, rows=disc)[:, :X.shape[1]]          # (as in bcf_iv: one gather, a spare column for pihat)
    pihat = (logit_fn or logit_link)(z[disc], Xd)                                          # R/bcf_itt.R:58-61
    fit_args = dict(nburn=n_burn, nsim=n_sim, random_seed=seed, keep_draws=False)
    fit_args.update(bcf_args)
    if binary:                                                                              # R/bcf_itt.R:166-170
        tauhat = (bartc_fn or _bart.bartc)(y[disc], z[disc], Xd, n_samples=n_sim, n_burn=n_burn, n_chains=2, seed=seed)["ite_mean"]
    else:
        own = dict(x_spare_column=True) if bcf_fn is None else {}
        tauhat = (bcf_fn or _bcf.bcf)(y[disc], z[disc], Xd, Xd, pihat, **dict(fit_args, **own))["tau_mean"]   # R/bcf_itt.R:71-75
    nodes, where = cart(Xd, tauhat, maxdepth=max_depth)                                     # rpart defaults (R/bcf_itt.R:82)
    ids, rows = _node_rows(nodes, where, names, y[index], w[index], z[index], _Rows(X, index))
    for nid in ids:
        r = rows[nid]
        if r is None:
            continue
        if nid == 1:
            print(f"The effect on the overall sample is {r['CCACE']}")
        else:
            print(f"The effect on the subpopulation {r['node']} is {r['CCACE']}")
        print(f"P-value {r['pvalue']}")
        print(f"P-value Weak-Instrument Test {r['Weak_IV_test']}")
        print(f"Proportion of observations in the node:  {r['Pi_obs']:.2f}")
    return None
