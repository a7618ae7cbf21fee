"""Multi-GPU plumbing: one process per GPU under torchrun, torch.distributed for the host-side exchange.

Two ways the path partitions (BASELINE.json north_star):
  * chains  -- one independent chain per GPU on replicated data, NO data-path collective; the only exchange is the
               final averaging / concatenation of per-chain posterior summaries (`combine_chain_means`).
  * shards  -- for n >= 1e7: contiguous row shards; per tree ONE all-reduce (sum) of the leaf sufficient statistics,
               done inside libbcfgpu with NCCL on int64 fixed-point limbs (so every rank sees identical bits and takes
               the identical Metropolis-Hastings decision).  This module only hands the ranks their rows and the NCCL id.

`fx_limbs` / `limbs_value` restate the device's fixed-point representation on the host (bcf_device.cuh: value * 2^44 in
three 21-bit limbs, summed per limb in int64): integer sums are associative, which is what makes the sharded sums
independent of the partition.
"""
from __future__ import annotations

import numpy as np

FX_SHIFT, LIMB_BITS = 44, 21


def shard_range(n: int, rank: int, world: int):
    """rows [lo, hi) of rank `rank`: contiguous, balanced, covering 0..n-1"""
    if not (0 <= rank < world):
        raise ValueError("bad rank")
    return rank * n // world, (rank + 1) * n // world


def fx_limbs(x):
    """(count, limb0, limb1, limb2) int64 sums of round(x * 2^44) split in 21-bit limbs (top limb signed)"""
    x = np.asarray(x, dtype=np.float64)
    if not np.all(np.abs(x) < 32768.0):
        raise ValueError("|x| must be < 2^15 (pass the standardised outcome)")
    v = np.rint(x * float(1 << FX_SHIFT)).astype(np.int64)
    m = (1 << LIMB_BITS) - 1
    return np.array([x.size, int((v & m).sum()), int(((v >> LIMB_BITS) & m).sum()), int((v >> (2 * LIMB_BITS)).sum())],
                    dtype=np.int64)


def limbs_value(l):
    """exact value of limb sums, rounded once to double (the device's limbs_to_double)"""
    v = int(l[1]) + (int(l[2]) << LIMB_BITS) + (int(l[3]) << (2 * LIMB_BITS))
    return v / float(1 << FX_SHIFT)


def allreduce_limbs(l, group=None):
    """sum of int64 limb vectors over the ranks (gloo / nccl); identical bits on every rank"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(l, dtype=np.int64).copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def exchange_unique_id(make_id, group=None) -> bytes:
    """rank 0 makes the 128-byte NCCL unique id (bcfgpu_nccl_unique_id), every rank gets it"""
    import torch.distributed as dist
    box = [make_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    if not isinstance(box[0], (bytes, bytearray)) or len(box[0]) != 128:
        raise RuntimeError("NCCL unique id must be 128 bytes")
    return bytes(box[0])


def combine_chain_means(local: dict, group=None) -> dict:
    """average per-observation posterior means over the chains (one per rank); every rank gets the result"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = {}
    for k in sorted(local):
        t = torch.from_numpy(np.ascontiguousarray(local[k], dtype=np.float64).copy())
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out[k] = t.cpu().numpy() / world
    return out


def gather_rows(local_rows: np.ndarray, group=None) -> np.ndarray:
    """concatenate per-shard vectors in rank order (shards are contiguous row ranges)"""
    import torch.distributed as dist
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, np.asarray(local_rows), group=group)
    return np.concatenate(parts)


def run_sharded(sampler_kwargs, y, z, x_con, x_mod, cuts, nburn, nsim, nthin=1, group=None):
    """observation-sharded chain: every rank passes the FULL host arrays and keeps its own rows; returns the full-length
    posterior mean of tau (rank order = row order) and this rank's device time."""
    import torch.distributed as dist
    from . import _lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(len(y), rank, world)
    s = _lib.Sampler(**sampler_kwargs)
    s.set_shard(rank, world, exchange_unique_id(_lib.nccl_unique_id, group) if world > 1 else None)
    s.set_data(y[lo:hi], z[lo:hi], x_con[lo:hi], x_mod[lo:hi], *cuts)
    s.run(nburn, nsim, nthin)
    tau = gather_rows(s.tau_mean(), group)
    ms = s.last_run_ms
    scal = s.scalars()
    scal["engine_in_use"] = s.engine_in_use
    s.close()
    return tau, ms, scal
