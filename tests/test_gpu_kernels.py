"""Kernel-level parity: each CUDA operator against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): bins, node assignments and leaf counts bit-exact; fp64 leaf sums within
1e-10 * sum|r| (fixed-point 2^-44 accumulation vs sequential fp64); log-likelihood terms within 1e-10 relative.
"""
import numpy as np
import pytest

from bcf_iv_b200 import _lib, prep
from helpers import gpu_sampler, make_problem, oracle_sampler

pytestmark = pytest.mark.gpu


from helpers import random_tree as _random_tree  # noqa: E402


def test_k0_bins_match_definition():
    """bin = #{cut <= x}: x < cut[c]  <=>  bin <= c, for binary, few-valued and 10 000-cutpoint columns."""
    rng = np.random.default_rng(0)
    for x in (rng.integers(0, 2, 5000).astype(float), rng.integers(0, 5, 5000).astype(float),
              rng.standard_normal(20000), np.repeat(rng.standard_normal(50), 100)):
        cuts = prep.cp_quantile(x)
        b = _lib.op_bin_column(x, cuts)
        assert np.array_equal(b, np.searchsorted(cuts, x, side="right"))
        for c in (0, len(cuts) // 2, len(cuts) - 1):
            assert np.array_equal(x < cuts[c], b <= c)


@pytest.mark.parametrize("continuous,nleaves", [(False, 6), (True, 9), (True, 40), (True, 128)])
def test_k1_k2_assignment_and_sufficient_statistics(continuous, nleaves):
    """set_tree -> K1 leaf assignment bit-exact with the oracle's tree walk over the raw doubles; K2 per-leaf counts
    exact and sums within 1e-10 * sum|r|, with and without a birth split; small trees take the REDUX path, trees
    with more than 8 keys the shared-atomics path."""
    pr = make_problem(n=4111, p=9, seed=13, continuous=continuous)
    g = gpu_sampler(pr, ntree_con=3, ntree_mod=2)
    o = oracle_sampler(pr, ntree_con=3, ntree_mod=2)
    rng = np.random.default_rng(5)
    perm, inv = pr.P.perm, pr.P.inv_perm
    for tid, cuts in ((1, pr.P.cuts_c), (3, pr.P.cuts_m)):
        ncut = [len(c) for c in cuts]
        heap, var, cut, mu = _random_tree(rng, ncut, nleaves)
        g.set_tree(tid, heap, var, cut, mu)
        o.set_tree(tid, heap, var, cut, mu)
        assert np.array_equal(g.leaf_ids(tid), o.assign_leaves(tid)[inv])
        assert g.verify_leaf_cache() == 0
        r = rng.standard_normal(pr.n) * 3.0
        z = pr.P.z
        st = g.op_suffstats(tid, r)
        # oracle: counts / sums per leaf for each group via unit weights restricted to the group
        rp = r[perm]
        zt = (z[perm] == 1).astype(float)
        h1, n1, _, s1 = o.allsuff(tid, zt, rp)            # phi = 1{treated}: sum phi, sum phi*r
        h0, n0, _, s0 = o.allsuff(tid, 1.0 - zt, rp)
        assert np.array_equal(st["heap"], h1)
        _, _, c1, _ = o.allsuff(tid, zt, rp)
        assert np.array_equal(st["cnt_trt"], np.round(o.allsuff(tid, zt, np.ones(pr.n))[3]).astype(np.int64))
        assert np.array_equal(st["cnt_ctl"], np.round(o.allsuff(tid, 1.0 - zt, np.ones(pr.n))[3]).astype(np.int64))
        tol = 1e-10 * np.abs(r).sum()
        assert np.all(np.abs(st["sum_trt"] - s1) <= tol) and np.all(np.abs(st["sum_ctl"] - s0) <= tol)
        # birth split of one leaf
        hl, vl, cl, _ = g.get_tree(tid)
        leaf_heaps = hl[vl < 0]
        lh = int(leaf_heaps[rng.integers(len(leaf_heaps))])
        v = int(rng.integers(len(ncut))); c = int(rng.integers(ncut[v]))
        sb = g.op_suffstats(tid, r, birth_leaf_heap=lh, var=v, cut=c)
        ot = o.getsuff_birth(tid, lh, v, c, zt, rp)        # {l.n0, l.n, l.sy, r.n0, r.n, r.sy} with phi = 1{treated}
        oc = o.getsuff_birth(tid, lh, v, c, 1.0 - zt, rp)
        b = sb["birth"]
        assert (b[0], b[1], b[4], b[5]) == (ot[1], oc[1], ot[4], oc[4])       # counts by group, exact
        assert abs(b[2] - ot[2]) <= tol and abs(b[3] - oc[2]) <= tol and abs(b[6] - ot[5]) <= tol and abs(b[7] - oc[5]) <= tol
        # the whole-leaf statistics are unchanged by a proposal
        assert np.array_equal(sb["cnt_trt"], st["cnt_trt"]) and np.allclose(sb["sum_trt"], st["sum_trt"], atol=tol, rtol=0)


def test_k2prime_all_candidate_histogram():
    """(count, residual sum) for every (leaf, variable, cutpoint bin) at once vs a plain loop; uint8 and uint16 columns."""
    pr = make_problem(n=3001, p=7, seed=17, continuous=True)
    g = gpu_sampler(pr, ntree_con=2, ntree_mod=2)
    o = oracle_sampler(pr, ntree_con=2, ntree_mod=2)
    rng = np.random.default_rng(3)
    r = rng.standard_normal(pr.n)
    for forest in (0, 1):
        for nleaf in (1, 3, 20):
            slot = rng.integers(-1, nleaf, pr.n).astype(np.int32)
            cg, sg, ms = g.op_hist_all(forest, slot, nleaf, r)
            co, so = o.hist_all(forest, slot[pr.P.perm], nleaf, r[pr.P.perm])
            assert np.array_equal(cg, co)
            assert np.allclose(sg, so, rtol=0, atol=1e-10 * np.abs(r).sum())
            assert ms > 0


def _exact_hist_two_valued(X, cuts, slot, nleaf, r):
    """per (leaf, column): exact counts and the exact sum of the 2^-44 fixed-point residuals (Python integers), bins 0 / 1"""
    fx = np.rint(r * 2.0 ** 44).astype(np.int64)
    cnt = np.zeros((nleaf, X.shape[1], 2), dtype=np.int64)
    sm = np.zeros((nleaf, X.shape[1], 2), dtype=np.float64)
    for l in range(nleaf):
        m = slot == l
        for v in range(X.shape[1]):
            one = m & (X[:, v] >= cuts[v])
            zero = m & ~(X[:, v] >= cuts[v])
            cnt[l, v, 0], cnt[l, v, 1] = zero.sum(), one.sum()
            sm[l, v, 0] = int(fx[zero].sum(dtype=object) if zero.any() else 0) / 2 ** 44      # (int / int: correctly rounded)
            sm[l, v, 1] = int(fx[one].sum(dtype=object) if one.any() else 0) / 2 ** 44
    return cnt, sm


@pytest.mark.parametrize("n,p", [(5000, 12), (70001, 33), (3000, 300), (600000, 20)])
def test_k2prime_on_the_tensor_cores_is_exact(n, p, monkeypatch):
    """K2' for two-valued columns as an int8 contraction on tcgen05 (k_hist_mma: TMA-fed bins, 7-bit limbs of the fixed-point
    residual, int32 accumulators in tensor memory): counts exact and sums BIT-EXACT against integer arithmetic on the host,
    for 1..12 leaves, skipped labels, more than 240 columns (two column
    passes); and equal to the CUDA-core kernel it replaces (k_hist_bits, fp64 lane sums rounded once) within that kernel's 1e-13."""
    rng = np.random.default_rng(n + p)
    X = rng.integers(0, 2, (n, p)).astype(np.float64)
    z = rng.integers(0, 2, n).astype(np.int32)
    y = rng.standard_normal(n)
    cuts = [np.array([0.5])] * p
    flat, off = prep.pack_cuts(cuts)
    s = _lib.Sampler(lambda_=1.0, sigma_init=1.0, seed=1, ntree_con=2, ntree_mod=2)
    s.set_data(y, z, X, X, flat, off, flat, off)
    r = rng.standard_normal(n) * 3.0
    r[::97] = 0.0
    r[5] = 32767.99                                                      # the edge of the fixed-point range, both signs
    r[6] = -32767.99
    for nleaf in ((1, 2, 5, 8, 12) if n < 100000 else (3, 12)):
        slot = rng.integers(-1, nleaf, n).astype(np.int32)
        monkeypatch.delenv("BCFGPU_NO_HIST_MMA", raising=False)
        want_c, want_s = _exact_hist_two_valued(X, np.full(p, 0.5), slot, nleaf, r)
        cg, sg, ms = s.op_hist_all(1, slot, nleaf, r)
        cg = cg.reshape(nleaf, p, 2); sg = sg.reshape(nleaf, p, 2)
        assert np.array_equal(cg, want_c)
        assert np.array_equal(sg, want_s)                               # bit-exact: integer sums, one rounding
        if nleaf <= 8:   # (k_hist_bits rounds a CTA's fp64 SUM to fixed point: it needs |sum| < 2^15, so a smaller r for the pair)
            rs = r / 1024.0
            cm, smm, _ = s.op_hist_all(1, slot, nleaf, rs)
            monkeypatch.setenv("BCFGPU_NO_HIST_MMA", "1")
            cb, sb, _ = s.op_hist_all(1, slot, nleaf, rs)
            assert np.array_equal(cb, cm) and np.array_equal(cm.reshape(nleaf, p, 2), cg)
            assert np.allclose(sb, smm, rtol=0, atol=1e-13 * np.abs(rs).sum())
    s.close()


def test_k3_integrated_likelihood():
    from oracle import orc
    rng = np.random.default_rng(1)
    a = rng.uniform(1, 1e6, 200); b = rng.standard_normal(200) * 1e3; t = rng.uniform(0.01, 2, 200)
    ref = np.array([orc.lil(x, y, z) for x, y, z in zip(a, b, t)])
    assert np.allclose(_lib.op_lil(a, b, t), ref, rtol=1e-10, atol=0)


def test_philox_stream_matches_oracle():
    from oracle import orc
    for args in ((42, 0, 0, 0, 0, 0), (42, 3, 17, 249, 3, 1234567), (2**63 + 5, 7, 4000, 0xFFFFFFFF, 15, 9)):
        g = _lib.op_rng_probe(*args)
        o = orc.rng_probe(*args)
        assert g[0] == o[0] and g[1] == o[1]                 # uniforms: integer arithmetic, bit-exact
        assert g[2] == pytest.approx(o[2], rel=1e-13, abs=1e-15)   # Box-Muller: device vs glibc log/cos


@pytest.mark.parametrize("pattern", ["random", "treated_first", "control_first", "all_control", "all_treated", "blocks"])
@pytest.mark.parametrize("pageable", [False, True])
def test_set_data_permutation_and_binning_layouts(pattern, pageable, monkeypatch):
    """set_data's device-side treated-first permutation (k_place) and the run-compacting binning kernel (k_bin), for the
    z patterns that stress them: sorted either way (a block's treated and control runs touch), a single group, long
    blocks; rows not a multiple of the 4096-row chunk / 2048-row block / 256-row tile; pinned-style direct copies and the
    pageable path through the pinned staging buffers.  Checked through K1: the leaf of every row under a deep random
    tree (it reads the binned columns at the permuted positions) against the oracle's walk over the raw doubles."""
    if pageable:
        monkeypatch.setenv("BCFGPU_FORCE_PAGEABLE", "1")
    n = 2 * 4096 + 2048 + 300 + 7
    pr = make_problem(n=n, p=9, seed=17, continuous=True)
    rng = np.random.default_rng(3)
    z = {"random": pr.P.z,
         "treated_first": (np.arange(n) < n // 3).astype(np.int32),
         "control_first": (np.arange(n) >= n // 3).astype(np.int32),
         "all_control": np.zeros(n, np.int32), "all_treated": np.ones(n, np.int32),
         "blocks": ((np.arange(n) // 1500) % 2).astype(np.int32)}[pattern]
    P = pr.P
    perm = np.argsort(-z, kind="stable")
    inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
    from oracle import orc
    o = orc.OracleSampler(P.yscale[perm], int(z.sum()), P.x_c[perm], P.x_m[perm], P.cuts_c, P.cuts_m, lambda_=P.lambda_,
                          sigma_init=P.sigma_init, ntree_con=2, ntree_mod=1)
    g = _lib.Sampler(lambda_=P.lambda_, sigma_init=P.sigma_init, ntree_con=2, ntree_mod=1)
    g.set_data(P.yscale, z, P.x_c, P.x_m, pr.cuts_c_flat, pr.off_c, pr.cuts_m_flat, pr.off_m)
    for tid, cuts in ((0, P.cuts_c), (2, P.cuts_m)):
        heap, var, cut, mu = _random_tree(rng, [len(c) for c in cuts], 60)
        g.set_tree(tid, heap, var, cut, mu)
        o.set_tree(tid, heap, var, cut, mu)
        assert np.array_equal(g.leaf_ids(tid), o.assign_leaves(tid)[inv])
    assert g.verify_leaf_cache() == 0
    # y went through the same permutation: the residual the sweeps start from is built on it
    o.refit()
    for _ in range(3):
        g.run(1, 0)
        o.sweeps(1)
        assert np.array_equal(g.trace()[0], o.trace()[0])
    assert abs(g.scalars()["sigma"] - o.scalars()["sigma"]) <= 1e-9 * o.scalars()["sigma"]
    g.close(); o.close()


def test_caching_allocator_reuses_blocks_and_releases_them():
    """A second fit of the same shape takes its device blocks from the library's cache (no new cudaMalloc), and
    bcfgpu_release_cached_memory hands everything back."""
    pr = make_problem(n=3000, p=10, seed=2)
    _lib.release_cached_memory()
    s = gpu_sampler(pr); s.run(1, 0); s.close()
    st0 = _lib.cache_stats()
    assert st0["device_bytes"] > 0
    s = gpu_sampler(pr); s.run(1, 0)
    st1 = _lib.cache_stats()
    assert st1["device_bytes"] < st0["device_bytes"]          # blocks were taken out of the cache
    s.close()
    assert _lib.cache_stats()["device_bytes"] >= st0["device_bytes"]
    _lib.release_cached_memory()
    assert _lib.cache_stats()["device_bytes"] == 0


@pytest.mark.parametrize("pageable", [False, True])
def test_one_cutpoint_columns_cross_the_bus_as_bytes(monkeypatch, pageable):
    """set_data's staging pass writes (cut <= x) as a byte for every run of one-cutpoint columns (the reference's two-valued
    covariates) instead of copying the double: same bins -- the leaf assignment of a random tree equals the oracle's walk
    over the raw doubles either way -- and an eighth of the bytes over PCIe for those columns."""
    if pageable:
        monkeypatch.setenv("BCFGPU_FORCE_PAGEABLE", "1")
    pr = make_problem(n=6007, p=12, seed=21)                       # 12 binary columns + the 10 000-cutpoint pihat column
    o = oracle_sampler(pr, ntree_con=2, ntree_mod=1)
    rng = np.random.default_rng(8)
    heap, var, cut, mu = _random_tree(rng, [len(c) for c in pr.P.cuts_c], 24)
    o.set_tree(0, heap, var, cut, mu)
    want = o.assign_leaves(0)[pr.P.inv_perm]
    moved = {}
    for packed in (True, False):
        if packed:
            monkeypatch.delenv("BCFGPU_NO_HOST_PACK", raising=False)
        else:
            monkeypatch.setenv("BCFGPU_NO_HOST_PACK", "1")
        g = gpu_sampler(pr, ntree_con=2, ntree_mod=1)
        g.set_tree(0, heap, var, cut, mu)
        assert np.array_equal(g.leaf_ids(0), want)
        assert g.verify_leaf_cache() == 0
        moved[packed] = g.h2d_bytes
        g.close()
    assert moved[False] - moved[True] == 12 * 6007 * 7            # 12 columns x n rows x (8 - 1) bytes
