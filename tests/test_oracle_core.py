"""The CPU oracle against plain NumPy definitions and against the model's own known truth.

PARITY UNPINNED: the reference has no golden vectors for this path and its sampler (third-party `bcf`)
cannot be run here, so the oracle is pinned from the other side: (1) its building blocks equal brute-force NumPy
restatements of SURVEY Appendix A; (2) identities the algorithm must satisfy (lil == lilhet, thread-count
independence of the move sequence); (3) the chain recovers the DGP's known effects (R/generate_dataset.R:52-55:
CACE = +2 / 0 / 0 / -2 by (x1, x2) cell, ITT = CACE * compliance).
"""
import numpy as np
import pytest

from helpers import make_problem, oracle_sampler
from oracle import orc


def _np_assign(heap, var, cut, X, cuts):
    """leaf heap id per row by walking the serialised tree: left iff x[v] < cuts[v][c]"""
    rule = {int(h): (int(v), int(c)) for h, v, c in zip(heap, var, cut)}
    out = np.zeros(X.shape[0], dtype=np.uint32)
    for i in range(X.shape[0]):
        h = 1
        while rule[h][0] >= 0:
            v, c = rule[h]
            h = 2 * h if X[i, v] < cuts[v][c] else 2 * h + 1
        out[i] = h
    return out


def _grow(o, tid, rules):
    """set tree `tid` to the full tree given {heap: (var, cut)} for internal nodes"""
    heap, var, cut, mu = [], [], [], []
    todo = [1]
    while todo:
        h = todo.pop()
        if h in rules:
            heap.append(h); var.append(rules[h][0]); cut.append(rules[h][1]); mu.append(0.0)
            todo += [2 * h, 2 * h + 1]
        else:
            heap.append(h); var.append(-1); cut.append(-1); mu.append(0.1 * h)
    o.set_tree(tid, np.array(heap, np.uint32), np.array(var, np.int32), np.array(cut, np.int32), np.array(mu))
    return np.array(heap), np.array(var), np.array(cut)


@pytest.mark.parametrize("continuous", [False, True])
def test_assign_allsuff_getsuff_match_numpy(continuous):
    pr = make_problem(n=700, p=6, seed=4, continuous=continuous)
    o = oracle_sampler(pr, ntree_con=2, ntree_mod=2)
    perm = pr.P.perm
    rng = np.random.default_rng(0)
    for tid, X, cuts in ((0, pr.P.x_c[perm], pr.P.cuts_c), (2, pr.P.x_m[perm], pr.P.cuts_m)):
        nc = [len(c) for c in cuts]
        rules = {1: (0, 0), 2: (1, 0), 3: (len(nc) - 1, nc[-1] // 2)}
        if nc[-1] > 4:
            rules[7] = (len(nc) - 1, nc[-1] // 2 + nc[-1] // 4)
        heap, var, cut = _grow(o, tid, rules)
        leaf = o.assign_leaves(tid)
        assert np.array_equal(leaf, _np_assign(heap, var, cut, X, cuts))
        phi = rng.uniform(0.5, 2.0, pr.n); R = rng.standard_normal(pr.n)
        h, n0, sp, spr = o.allsuff(tid, phi, R)
        assert list(h) == sorted(set(leaf.tolist()), key=lambda x: x << (30 - (int(x).bit_length() - 1)))  # DFS order
        for k, hh in enumerate(h):
            m = leaf == hh
            assert n0[k] == m.sum()
            assert sp[k] == pytest.approx(phi[m].sum(), rel=1e-12)
            assert spr[k] == pytest.approx((phi[m] * R[m]).sum(), rel=1e-10, abs=1e-10)
        hh = int(h[0]); v = 2 % len(nc); c = nc[v] // 2
        out = o.getsuff_birth(tid, hh, v, c, phi, R)
        m = leaf == hh
        left = m & (X[:, v] < cuts[v][c]); right = m & ~(X[:, v] < cuts[v][c])
        assert (out[0], out[3]) == (left.sum(), right.sum())
        assert out[1] == pytest.approx(phi[left].sum()) and out[4] == pytest.approx(phi[right].sum())
        assert out[2] == pytest.approx((phi * R)[left].sum(), abs=1e-10) and out[5] == pytest.approx((phi * R)[right].sum(), abs=1e-10)


def test_hist_all_is_the_set_of_all_candidate_splits():
    """prefix sums of the (leaf, var, bin) histogram over bins <= c are getsuff's left child for cut c"""
    pr = make_problem(n=600, p=5, seed=8, continuous=True)
    o = oracle_sampler(pr, ntree_con=1, ntree_mod=1)
    rng = np.random.default_rng(2)
    r = rng.standard_normal(pr.n)
    slot = np.zeros(pr.n, dtype=np.int32)
    cnt, s = o.hist_all(0, slot, 1, r)
    off = o.oc
    for v in (0, 3, 5):
        nb = off[v + 1] - off[v] + 1
        cs = np.cumsum(cnt[0, off[v] + v: off[v] + v + nb]); ss = np.cumsum(s[0, off[v] + v: off[v] + v + nb])
        assert cs[-1] == pr.n and ss[-1] == pytest.approx(r.sum(), abs=1e-9)
        for c in (0, (nb - 1) // 2):
            out = o.getsuff_birth(0, 1, v, c, np.ones(pr.n), r)
            assert out[0] == cs[c] and out[2] == pytest.approx(ss[c], abs=1e-9)


def test_lil_weighted_equals_homoskedastic_in_every_ratio():
    """lil(n, sy, sy2, sigma, tau) and the weighted form differ by terms that cancel in l + r - t"""
    rng = np.random.default_rng(0)
    for _ in range(50):
        sigma, tau = rng.uniform(0.3, 2), rng.uniform(0.05, 1)
        yl, yr = rng.standard_normal(rng.integers(5, 60)), rng.standard_normal(rng.integers(5, 60)) + 0.5
        yt = np.concatenate([yl, yr])
        phi = 1 / sigma ** 2
        w = sum(sg * orc.lil(phi * len(y), phi * y.sum(), tau) for sg, y in ((1, yl), (1, yr), (-1, yt)))
        h = sum(sg * orc.lil_homosked(len(y), y.sum(), (y * y).sum(), sigma, tau) for sg, y in ((1, yl), (1, yr), (-1, yt)))
        assert w == pytest.approx(h, abs=1e-9)


def test_move_sequence_does_not_depend_on_thread_count():
    pr = make_problem(n=900, p=10, seed=2)
    a = oracle_sampler(pr, seed=3, ntree_con=40, ntree_mod=10, nthreads=1)
    b = oracle_sampler(pr, seed=3, ntree_con=40, ntree_mod=10, nthreads=4)
    a.sweeps(12); b.sweeps(12)
    assert np.array_equal(a.trace()[0], b.trace()[0]) and a.counters() == b.counters()
    assert a.scalars()["sigma"] == pytest.approx(b.scalars()["sigma"], rel=1e-9)


def test_chain_recovers_the_known_effects_of_the_readme_workload():
    """Config C1 (README.md:79-97: n=1000 -> n_s=500, p=10, compliance .75): posterior-mean ITT per (x1,x2) cell
    within Monte-Carlo error of +1.5 / 0 / 0 / -1.5 (R/generate_dataset.R:52-55 times the compliance)."""
    pr = make_problem(n=500, p=10, seed=0)
    o = oracle_sampler(pr, seed=42)
    out = o.run(300, 300)
    assert out["saved"] == 300
    tau = pr.P.sdy * out["tau_mean"][pr.P.inv_perm]
    X = pr.data["X"]
    cell = lambda a, b: (X[:, 0] == a) & (X[:, 1] == b)
    assert tau[cell(0, 0)].mean() == pytest.approx(1.5, abs=0.45)
    assert tau[cell(1, 1)].mean() == pytest.approx(-1.5, abs=0.45)
    assert abs(tau[cell(0, 1)].mean()) < 0.45 and abs(tau[cell(1, 0)].mean()) < 0.45
    sig = pr.P.sdy * out["sigma"]
    assert 0.8 < sig.mean() < 1.35          # noise sd 1 plus the compliance mixture
    c = o.counters()
    assert 0.02 < c["birth_acc"] / c["birth_prop"] < 0.6 and c["death_acc"] > 0


def test_min_leaf_rule_and_tree_validity_along_the_chain():
    pr = make_problem(n=300, p=10, seed=6)
    o = oracle_sampler(pr, seed=1, ntree_con=20, ntree_mod=10)
    o.sweeps(60)
    for t in range(30):
        h, n0, _, _ = o.allsuff(t, np.ones(pr.n), np.zeros(pr.n))
        assert n0.sum() == pr.n
        if len(h) > 1:
            assert n0.min() >= 5, "a birth left a child with fewer than 5 observations"
        heap, var, cut, _ = o.get_tree(t)
        ids = set(int(x) for x in heap)
        for hh, v in zip(heap, var):
            assert (v >= 0) == (2 * int(hh) in ids) == (2 * int(hh) + 1 in ids)
