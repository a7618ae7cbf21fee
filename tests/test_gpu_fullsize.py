"""Full-size parity (BASELINE.json headline shape: n_s = 1e6, p = 100 + pihat, 200 + 50 trees) through size-independent
properties -- the CPU oracle needs ~1 s per sweep on 16 cores at this size, so it is not the checker here:
  * the pipelined, batched and tree-at-a-time engines walk bit-identical chains (integer statistics);
  * the incrementally maintained leaf-slot cache equals a fresh tree walk over the binned columns (K1) for all 250 trees;
  * the cached leaf counts of every tree add up to n by treatment group;
  * K2 leaf sums of a tree add up to the plain sum of the residual (a checksum of checksums), to 1e-10 * sum|r|."""
import numpy as np
import pytest

from bcf_iv_b200 import _lib, workload

pytestmark = pytest.mark.gpu

N, P = 1_000_000, 100


@pytest.fixture(scope="module")
def inputs():
    return workload.synthetic_sampler_inputs(N, P, seed=0)


def _sampler(w, engine, **kw):
    s = _lib.Sampler(engine=engine, lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1, **kw)
    s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
    return s


def test_engines_bit_identical_at_headline_size(inputs):
    ref = None
    for engine in (_lib.ENGINE_PIPELINED, _lib.ENGINE_BATCHED, _lib.ENGINE_PERSISTENT):
        s = _sampler(inputs, engine)
        s.run(3, 2)
        got = (s.scalars(), s.counters(), s.trace()[0].copy(), s.tau_mean(), s.sigma())
        if engine == _lib.ENGINE_PIPELINED:
            assert s.verify_leaf_cache() == 0
        s.close()
        if ref is None:
            ref = got
        else:
            assert got[0] == ref[0] and got[1] == ref[1]
            assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3]) and np.array_equal(got[4], ref[4])


def test_leaf_statistics_are_conserved_at_headline_size(inputs):
    s = _sampler(inputs, _lib.ENGINE_AUTO)
    s.run(6, 0)
    assert s.verify_leaf_cache() == 0
    ntrt = int(inputs["z"].sum())
    r = np.random.default_rng(3).standard_normal(N)
    z = inputs["z"].astype(bool)
    for tree in (0, 57, 199, 200, 249):
        st = s.op_suffstats(tree, r)
        assert st["cnt_trt"].sum() == ntrt and st["cnt_ctl"].sum() == N - ntrt
        tol = 1e-10 * np.abs(r).sum()
        assert abs(st["sum_trt"].sum() - r[z].sum()) <= tol and abs(st["sum_ctl"].sum() - r[~z].sum()) <= tol
        ids = s.leaf_ids(tree)
        for h, ct, cc in zip(st["heap"], st["cnt_trt"], st["cnt_ctl"]):
            m = ids == h
            assert (m & z).sum() == ct and (m & ~z).sum() == cc
    # all-candidate histogram: every variable's bins partition the rows of every leaf
    slot = (np.arange(N) % 3).astype(np.int32)
    cnt, sm, ms = s.op_hist_all(1, slot, 3, r)
    off = s.off_mod
    for v in (0, 50, 99):
        blk = slice(int(off[v]) + v, int(off[v + 1]) + v + 1)
        assert np.array_equal(cnt[:, blk].sum(axis=1), np.bincount(slot, minlength=3))
        assert np.allclose(sm[:, blk].sum(axis=1), np.bincount(slot, weights=r, minlength=3), rtol=0, atol=1e-10 * np.abs(r).sum())
    s.close()


def test_chains_are_repeatable_run_to_run(inputs):
    """Same seed, same engine, several runs: identical bits every time (integer statistics, no ordering left to chance).
    Catches races that a single comparison would only see once in a while (scripts/stress_repro.py runs it longer)."""
    for engine in (_lib.ENGINE_PIPELINED, _lib.ENGINE_BATCHED):
        ref = None
        for _ in range(4):
            s = _sampler(inputs, engine)
            s.run(2, 2)
            got = (s.scalars(), s.counters(), s.trace()[0].tobytes(), s.tau_mean().tobytes())
            s.close()
            if ref is None:
                ref = got
            assert got == ref


def test_all_candidate_histogram_at_headline_size(inputs):
    """K2' on the full matrix (1e6 x 100 two-valued columns + the 10 000-cutpoint pihat column): per leaf and column the
    bins add up to the leaf's count and residual sum, and a few columns agree with a direct NumPy evaluation."""
    w = inputs
    n = len(w["y"])
    s = _lib.Sampler(lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1)
    s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
    rng = np.random.default_rng(5)
    r = rng.standard_normal(n)
    for forest, off in ((1, w["off_mod"]), (0, w["off_con"])):
        for nleaf in (1, 3):
            slot = rng.integers(0, nleaf, n).astype(np.int32)
            cnt, sm, ms = s.op_hist_all(forest, slot, nleaf, r)
            p = len(off) - 1
            nbt = int(off[p]) + p
            cnt = cnt.reshape(nleaf, nbt); sm = sm.reshape(nleaf, nbt)
            tol = 1e-10 * np.abs(r).sum()
            for l in range(nleaf):
                m = slot == l
                for v in range(p):
                    lo, hi = int(off[v]) + v, int(off[v + 1]) + v + 1
                    assert cnt[l, lo:hi].sum() == m.sum()
                    assert abs(sm[l, lo:hi].sum() - r[m].sum()) <= tol
            X = w["x_mod"] if forest == 1 else w["x_con"]
            cuts = w["cuts_mod"] if forest == 1 else w["cuts_con"]
            for v in (0, 1, p // 2, p - 1):
                c = cuts[int(off[v]):int(off[v + 1])]
                b = np.searchsorted(c, X[:, v], side="right")          # bin = #cutpoints <= x
                lo = int(off[v]) + v
                for l in range(nleaf):
                    m = slot == l
                    want_c = np.bincount(b[m], minlength=len(c) + 1)
                    want_s = np.bincount(b[m], weights=r[m], minlength=len(c) + 1)
                    assert np.array_equal(cnt[l, lo:lo + len(c) + 1], want_c)
                    assert np.allclose(sm[l, lo:lo + len(c) + 1], want_s, rtol=0, atol=tol)
    s.close()
