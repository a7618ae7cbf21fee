"""N > 1 host logic on CPU: world_size-2 gloo process groups.  The data path itself needs GPUs; what is
checked here is everything around it -- row sharding, the NCCL-id hand-over, chain averaging, and the property the
sharded mode rests on: fixed-point limb sums all-reduced over shards are bit-identical to the single-process sums."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from bcf_iv_b200 import multi


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _shard_sums(rank, world):
    y = np.random.default_rng(0).standard_normal(10007) * 3.0
    lo, hi = multi.shard_range(len(y), rank, world)
    red = multi.allreduce_limbs(multi.fx_limbs(y[lo:hi]))
    return red, float(np.sum(y[lo:hi]))


def test_sharded_fixed_point_sums_are_partition_independent():
    y = np.random.default_rng(0).standard_normal(10007) * 3.0
    whole = multi.fx_limbs(y)
    out = _spawn(_shard_sums, 2)
    for red, _ in out:
        assert np.array_equal(red, whole)                      # identical bits on every rank, whatever the split
    assert multi.limbs_value(whole) == pytest.approx(y.sum(), abs=len(y) * 2.0 ** -45)
    assert out[0][1] + out[1][1] != 0                          # (the plain fp64 partial sums are what is NOT portable)


def _ids(rank, world):
    uid = multi.exchange_unique_id(lambda: bytes(range(128)))
    lo, hi = multi.shard_range(1001, rank, world)
    rows = multi.gather_rows(np.arange(lo, hi))
    means = multi.combine_chain_means({"tau": np.full(5, float(rank)), "mu": np.arange(5.0) * (rank + 1)})
    return uid, (lo, hi), rows, means


def test_unique_id_rows_and_chain_average():
    out = _spawn(_ids, 2)
    assert out[0][0] == out[1][0] == bytes(range(128))
    assert out[0][1] == (0, 500) and out[1][1] == (500, 1001)
    for r in out:
        assert np.array_equal(r[2], np.arange(1001))
        assert np.allclose(r[3]["tau"], 0.5) and np.allclose(r[3]["mu"], np.arange(5.0) * 1.5)


def test_shard_range_covers_everything():
    for n, w in ((10, 3), (1, 2), (10_000_019, 8)):
        edges = [multi.shard_range(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
    with pytest.raises(ValueError):
        multi.shard_range(10, 2, 2)
