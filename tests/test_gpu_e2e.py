"""End-to-end on the GPU through the public front ends: bcf() (the call of R/bcf_iv.R:89), bcf_iv(), bcf_itt(), and
the two multi-GPU modes.  Level-2 parity of BASELINE.json: posterior estimands against the model's known truth and
against the CPU oracle, within Monte-Carlo error."""
import os
import socket
import sys

import numpy as np

import helpers
import pytest

from bcf_iv_b200 import _lib, bayesiv, bcf, datasets
from helpers import make_problem, oracle_sampler

pytestmark = pytest.mark.gpu


def test_bcf_front_end_matches_oracle_posterior_means():
    """bcf(...)$tau -> colMeans, in the caller's row order and y units, against the oracle's run of the same chain"""
    d = datasets.generate_dataset_wide(n=600, p=8, seed=2)
    pihat = 0.05 * np.random.default_rng(5).standard_normal(600)
    fit = bcf.bcf(d["y"], d["z"], d["X"], d["X"], pihat, nburn=100, nsim=60, random_seed=17)
    assert fit["tau"].shape == (60, 600) and fit["sigma"].shape == (60,)
    assert np.allclose(fit["tau"].mean(axis=0), fit["tau_mean"], rtol=1e-9, atol=1e-9)
    from bcf_iv_b200 import prep
    P = helpers.prepare(d["y"], d["z"], d["X"], d["X"], pihat)
    class Pr: pass
    pr = Pr(); pr.P = P
    o = oracle_sampler(pr, seed=17, max_leaves=128)
    out = o.run(100, 60, 1)
    assert np.allclose(fit["tau_mean"], P.sdy * out["tau_mean"][P.inv_perm], rtol=1e-7, atol=1e-7)
    assert np.allclose(fit["sigma"], P.sdy * out["sigma"], rtol=1e-7)
    assert np.array_equal(fit["perm"], P.perm + 1)


def test_chains_are_independent_and_stack_like_bcf2():
    d = datasets.generate_dataset_wide(n=400, p=6, seed=4)
    pihat = np.zeros(400) + 0.01 * np.arange(400) / 400
    two = bcf.bcf(d["y"], d["z"], d["X"], d["X"], pihat, nburn=100, nsim=20, random_seed=3, n_chains=2)
    one = bcf.bcf(d["y"], d["z"], d["X"], d["X"], pihat, nburn=100, nsim=20, random_seed=3, n_chains=1)
    assert two["tau"].shape == (40, 400)
    assert np.array_equal(two["tau"][:20], one["tau"])            # chain 0 alone == chain 0 among two
    assert not np.allclose(two["tau"][20:], one["tau"])            # chain 1 has its own Philox key


def test_bcf_iv_recovers_the_readme_subgroups():
    """README.md:79-97 workload (n=1000, p=10, compliance .75, effect 2): the discovered tree splits on x1 and x2 and
    the leaf CCACEs are +2 / -2 in the two effect cells (R/generate_dataset.R:52-55)."""
    d = datasets.generate_dataset(n=1000, p=10, compliance=0.75, seed=7)
    tab = bayesiv.bcf_iv(d["y"], d["w"], d["z"], d["X"], n_burn=400, n_sim=400, max_depth=2, seed=42)
    assert list(tab.columns) == bayesiv.COLUMNS and tab.index[0] == 1
    import pandas as pd
    assert pd.isna(tab.loc[1, "node"]) and 0.5 < float(tab.loc[1, "Pi_compliers"]) < 1.0
    rules = " ".join(str(r) for r in tab["node"].dropna())
    assert "x1" in rules and "x2" in rules
    leaves = tab.dropna(subset=["Adj_pvalue"])
    eff = leaves["CCACE"].astype(float)
    assert eff.max() > 1.2 and eff.min() < -1.2
    # the two effect leaves are significant after Holm, the complier share is about .75 + the never-assigned half
    assert (leaves["Adj_pvalue"].astype(float) < 0.05).sum() >= 2


@pytest.mark.parametrize("pic_model", ["bartc", "bcf"])
def test_bcf_iv_binary_outcome_runs_through_the_gpu_sampler(pic_model):
    """BASELINE config 4 in small: binary outcome, low compliance, correlated covariates.  The reference fits both the
    complier proportion and the ITT with bartc there (R/bcf_iv.R:78, 198): probit BART on the GPU (pic_model = "bartc",
    the default); "bcf" keeps the linear-probability stand-in.  The effect cells (x1 = x2 = 0: y more often 1 under
    assignment; x1 = x2 = 1: less often) keep their signs."""
    import warnings
    d = datasets.generate_dataset_wide(n=4000, p=10, rho=0.5, compliance=0.6, seed=3, binary_outcome=True)
    assert set(np.unique(d["y"])) <= {0.0, 1.0}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)               # ("y appears to be discrete": the bcf stand-in's preamble)
        det = {}
        tab = bayesiv.bcf_iv(d["y"], d["w"], d["z"], d["X"], binary=True, n_burn=300, n_sim=300, max_depth=2, seed=42,
                             pic_model=pic_model, details=det)
    assert list(tab.columns) == bayesiv.COLUMNS
    if pic_model == "bartc":
        # the subgroup CART grew on the device, over the covariates bartc's response model holds binned there (its leading
        # columns), and is the host restatement of rpart's tree
        assert det["cart_on_device"] is True
        nd_h, wh_h = bayesiv.cart(d["X"][det["disc"]], det["tauhat"], maxdepth=2, cp=0.01, minsplit=10)
        assert sorted(nd_h) == sorted(int(i) for i in tab.index if tab.loc[i, "node"] is not None)
        assert sorted({nd.var for nd in nd_h.values() if nd.var is not None}) == det["split_vars"]
    rules = " ".join(str(r) for r in tab["node"].dropna())
    assert "x1" in rules or "x2" in rules
    eff = tab.dropna(subset=["Adj_pvalue"])["CCACE"].astype(float)
    assert eff.max() > 0.15 and eff.min() < -0.15


def test_bcf_itt_prints_and_returns_none(capsys):
    d = datasets.generate_dataset(n=600, p=10, seed=9)
    assert bayesiv.bcf_itt(d["y"], d["w"], d["z"], d["X"], n_burn=150, n_sim=150) is None
    out = capsys.readouterr().out
    assert "The effect on the overall sample is" in out and "P-value Weak-Instrument Test" in out


# ---- observation shards over 2 GPUs (needs 2 devices: run with `gpurun --gpus 2`) ---------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_worker(rank, world, port, ret, engine, n):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from bcf_iv_b200 import multi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr = make_problem(n=n, p=8, seed=12)
    P = pr.P
    # shards must be contiguous in the sampler's input order; any order is fine since the library permutes per shard
    kw = dict(lambda_=P.lambda_, sigma_init=P.sigma_init, con_sd=P.con_sd, mod_sd=P.mod_sd, seed=5, ntree_con=30,
              ntree_mod=10, device=rank, engine=engine)
    tau, ms, scal = multi.run_sharded(kw, P.yscale, P.z, P.x_c, P.x_m,
                                      (pr.cuts_c_flat, pr.off_c, pr.cuts_m_flat, pr.off_m), 6, 6)
    ret[rank] = (tau, scal)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("engine", [_lib.ENGINE_GRAPH, _lib.ENGINE_AUTO])
def test_observation_shards_walk_the_single_gpu_chain(engine, world):
    """2 / 4 / 8 shards == the same chain on one GPU, BIT FOR BIT: identical decisions, sigma and posterior mean of tau
    (the exchanged statistics are integers, so they do not depend on the partition).  Ragged shards
    (n is not a multiple of the rank count, of the 256-row tile or of anything else).
    ENGINE_GRAPH: one kernel per tree + ncclAllReduce of its leaf statistics (the checked baseline);
    ENGINE_AUTO: the sweep kernel exchanging each batch's words itself through peer memory (NVLink)."""
    if _lib.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import torch.multiprocessing as mp
    n = 30011
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    mp.spawn(_shard_worker, args=(world, _free_port(), ret, engine, n), nprocs=world, join=True)
    pr = make_problem(n=n, p=8, seed=12)
    from helpers import gpu_sampler
    s = gpu_sampler(pr, seed=5, ntree_con=30, ntree_mod=10, engine=1)
    s.run(6, 6)
    ref, sc = s.tau_mean(), s.scalars()
    for r in range(world):
        tau, scal = ret[r]
        assert np.array_equal(tau, ref)
        assert all(scal[k] == sc[k] for k in ("sigma", "mscale", "bscale1", "bscale0", "tau_con", "sweep"))
        assert scal["engine_in_use"][1] == (1 if engine == _lib.ENGINE_GRAPH else 2)    # NCCL per tree / in-kernel mailboxes


def test_cart_on_the_device_equals_the_host_cart():
    """the subgroup tree on tauhat grown from the all-candidate histogram kernel (K2') over the sampler's binned
    covariates == the host restatement of rpart (same candidates: midpoints of few-valued columns, same stopping rules)."""
    rng = np.random.default_rng(5)
    n, p = 6000, 12
    X = rng.integers(0, 2, (n, p)).astype(float)
    X[:, 5] = rng.integers(0, 4, n)                              # a four-valued column: three candidate cutpoints
    tau = 1.5 * (X[:, 0] == 0) * (X[:, 1] == 0) - 1.5 * (X[:, 0] == 1) * (X[:, 1] == 1) + 0.4 * (X[:, 5] >= 2) + 0.3 * rng.standard_normal(n)
    pr_y = rng.standard_normal(n)
    z = rng.integers(0, 2, n)
    fit = bcf.bcf(pr_y, z, X, X, 0.05 * rng.standard_normal(n), nburn=2, nsim=1, random_seed=1, keep_sampler=True)
    s = fit["sampler"]
    try:
        for depth, cp, minsplit in ((2, 0.01, 10), (3, 0.001, 20), (1, 0.01, 10), (4, 0.0, 40)):
            nd_h, wh_h = bayesiv.cart(X, tau, maxdepth=depth, cp=cp, minsplit=minsplit)
            nd_d, wh_d = bayesiv.cart_device(s, X, tau, maxdepth=depth, cp=cp, minsplit=minsplit)
            assert sorted(nd_h) == sorted(nd_d)
            assert np.array_equal(wh_h, wh_d)
            for k in nd_h:
                assert (nd_h[k].var, nd_h[k].cut, nd_h[k].n) == (nd_d[k].var, nd_d[k].cut, nd_d[k].n)
    finally:
        s.close()
    # and bcf_iv takes that path on the reference's binary designs
    d = datasets.generate_dataset(n=1000, p=10, compliance=0.75, seed=7)
    det = {}
    bayesiv.bcf_iv(d["y"], d["w"], d["z"], d["X"], n_burn=100, n_sim=100, max_depth=2, seed=42, details=det)
    assert det["cart_on_device"] is True and det["split_vars"] == [0, 1]


@pytest.mark.parametrize("nshards,engine", [(2, _lib.ENGINE_AUTO), (3, _lib.ENGINE_AUTO), (2, _lib.ENGINE_BATCHED), (4, _lib.ENGINE_BATCHED)])
def test_in_process_shards_on_one_gpu_walk_the_single_gpu_chain(nshards, engine):
    """The whole sharded path -- contiguous row shards, the in-kernel mailbox exchange of every batch's leaf statistics and of
    the sweep's moments, the agreement on the kernel -- on ONE device: bcfgpu_set_shard_local makes the handles of one
    process the shards of one chain (no NCCL, no IPC), each on its share of the SMs and driven by its own host thread.
    Bit for bit the chain of the unsharded sampler, with the pipelined kernel's tagged-word exchange
    (ENGINE_AUTO, two shards) and the batched kernel's flag protocol.
    The shards' sweep kernels wait for each other, so they must be resident at the same time; the hardware scheduler
    usually, but not always, starts the second one beside the first (observed: never a third pipelined kernel).  When it
    does not, the waiting kernel gives up after 4 s and says so: the case is then SKIPPED, not failed -- it runs in its
    own process because a kernel that gives up ends the CUDA context."""
    import json
    import subprocess
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inprocess_shards_worker.py")
    out = None
    for attempt in range(3):
        r = subprocess.run([sys.executable, worker, str(nshards), str(engine)], capture_output=True, text=True, timeout=600)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        assert lines, r.stderr[-2000:]
        out = json.loads(lines[-1])
        if "error" not in out:
            break
    if "error" in out:
        msg = " | ".join(out["error"])
        if "gave up" in msg or "launches started" in msg or "launch failure" in msg:
            pytest.skip("the shards' kernels were not scheduled side by side on this device: " + msg[-300:])
        pytest.fail(msg)
    assert out["tau_identical"] and out["scalars_identical"] and out["moves_identical"], out
    # the sweep kernel itself exchanges: k_pipe's tagged words for two shards, the batched kernel's mailboxes + flags else
    want = [4 if engine == _lib.ENGINE_AUTO and nshards <= 2 else 3, 2]
    assert all(e == want for e in out["engine_in_use"]), out
