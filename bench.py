#!/usr/bin/env python
"""bench.py -- BCF Gibbs sweeps/sec at the headline shape (BASELINE.json: n=1M, p=100, 200 mu + 50 tau trees).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm: libbcfgpu.so through its C-ABI
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU arm: the oracle restatement of bcf

A STEP is one Gibbs sweep: one birth/death + leaf-redraw update of every one of the 250 trees, the
parameter-expansion scale draws and the sigma draw.  N > 1 (torchrun, one process per GPU)
runs ONE INDEPENDENT CHAIN PER GPU on the replicated data with no data-path collective (SURVEY 8e "chains"):
weak scaling, value = sweeps of all chains / max-over-ranks time.  The same run then ALSO measures the other way the
path partitions (BASELINE config 5): ONE chain with the observations sharded over the N GPUs (n = 1e7, p = 200), the
leaf statistics exchanged inside the sweep kernel over NVLink; it is attached to the JSON line as `"sharded": {...}`
together with an in-line proof that the N-shard chain equals the 1-GPU chain bit for bit (`shard_parity`).
`--mode sharded` makes the sharded chain (at --nobs / --ncov) the main line instead.

Timed region of `value`: inputs binned and resident in HBM, chain burnt in (tree sizes stationary), W warm-up
sweeps, then exactly K sweeps bracketed by barrier + synchronize, timed with CUDA events on the library's stream,
max over ranks.  `e2e`: the whole bcf()-shaped call through the C-ABI with HOST buffers -- create, set_data (H2D
of y, z, X and binning), run(K sweeps), colMeans(tau) back to the host -- by wall clock.

The oracle (oracle/) is used here ONLY as the timed CPU baseline (`cpu_baseline`, `--impl reference`); the measured
product path never touches it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bcf_gibbs_sweeps_per_sec"
UNIT = "sweeps/s"
T_CON, T_MOD = 200, 50


def algorithmic_bytes_per_sweep(n, T, bx=1):
    """per (obs, tree) read r 8 + write r 8 + leaf slot 1 + proposed variable's bin bx; 48 B/obs epilogue."""
    return n * (T * (17 + bx) + 48)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic_per_sweep(n, engine_name):
    """dram bytes per sweep from the committed ncu --set full capture of the same workload, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            tr = json.load(f)
        e = tr.get(f"{engine_name}:n={n}")
        return float(e["dram_bytes_per_sweep"]) if e else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.th = threading.Thread(target=pump, daemon=True)
        self.th.start()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        if self.proc is None:
            return None
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.15] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    """every host core this process may use (torchrun exports OMP_NUM_THREADS=1: the oracle takes the count explicitly)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_numa(index):
    """Pin this process (and so the first touch of its pinned buffers and the library's host threads) to the CPUs next to
    its GPU: with one rank per GPU on a two-socket box the host-to-device copies otherwise cross the socket link.
    Best effort: returns the cpulist used or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpulist
    except Exception:
        return None


def pinned_like(a):
    """copy of a numpy array in pinned host memory (what the e2e leg hands to the C-ABI), keeping Fortran order"""
    import torch
    if a.ndim == 2 and a.flags.f_contiguous:
        t = torch.empty((a.shape[1], a.shape[0]), dtype=torch.from_numpy(np.zeros(1, a.dtype)).dtype, pin_memory=True)
        v = t.numpy().T
    else:
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.zeros(1, a.dtype)).dtype, pin_memory=True)
        v = t.numpy()
    v[...] = a
    return v, t


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of bcf's sweep, all host threads, on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------------------
def oracle_sampler_from(w, rows, nthreads):
    from oracle import orc
    z = w["z"][rows]
    order = np.argsort(-z, kind="stable")                      # bcf: perm = order(z, decreasing = TRUE)
    rows = rows[order]
    p_con, p_mod = w["x_con"].shape[1], w["x_mod"].shape[1]
    cc = [w["cuts_con"][w["off_con"][v]:w["off_con"][v + 1]] for v in range(p_con)]
    cm = [w["cuts_mod"][w["off_mod"][v]:w["off_mod"][v + 1]] for v in range(p_mod)]
    return orc.OracleSampler(w["y"][rows], int(z.sum()), w["x_con"][rows], w["x_mod"][rows], cc, cm,
                             lambda_=w["lambda_"], sigma_init=w["sigma_init"], nthreads=nthreads)


def time_oracle(w, n_full, n_sample, sweeps, warm, nthreads):
    """sweeps/s of the CPU oracle scaled to the full n: cost is O(n) per tree pass (4 passes/tree)."""
    rows = np.arange(n_sample, dtype=np.int64)                 # rows are iid: the first n_sample are a random sample
    o = oracle_sampler_from(w, rows, nthreads)
    if warm:
        o.sweeps(warm)
    t0 = time.perf_counter()
    o.sweeps(sweeps)
    dt = time.perf_counter() - t0
    o.close()
    return (sweeps / dt) * (n_sample / n_full), dt


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from oracle import orc
    from bcf_iv_b200 import workload
    orc.build()
    nth = host_threads()
    n, p = args.n, args.p
    # calibrate on 20k rows, then size the sample so that warmup + steps sweeps take about args.cpu_budget seconds
    n_cal = min(n, 20000)
    w = workload.synthetic_sampler_inputs(min(n, 200000), p, seed=args.data_seed)
    _, dt_cal = time_oracle(w, n, n_cal, 1, 1, nth)
    per_row_sweep = dt_cal / n_cal
    n_sample = int(min(n, len(w["y"]), max(5000, args.cpu_budget / (per_row_sweep * (args.steps + args.warmup)))))
    val, dt = time_oracle(w, n, n_sample, args.steps, args.warmup, nth)
    sample = (f"{args.steps} timed sweeps (+{args.warmup} warm-up) of the oracle on the first {n_sample} rows of the "
              f"same synthetic workload, {nth} OpenMP threads over observations; sweeps/s scaled by n_sample/n "
              f"(cost is O(n) per tree pass)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world, "cpu"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nth, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the un-vendored bcf package's sweep (oracle/bcf_oracle.c), not bcf itself: "
                "the reference cannot be installed or run here (no R, bcf sources absent; SURVEY 8c)",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world, where):
    return {"workload": f"bcf sampler called directly: n_s={args.n}, p={args.p} binary covariates + pihat "
                        f"(p_con={args.p + 1}), {T_CON} mu + {T_MOD} tau trees (BASELINE.json metric shape)",
            "n": args.n, "p": args.p, "ntree_control": T_CON, "ntree_moderate": T_MOD,
            "parallelism": (f"chains x{world} (one independent chain per GPU, data replicated)" if args.mode == "chains"
                            else f"observation shards x{world} (one exchange of the leaf statistics per batch of trees)"),
            "where": where}


# ------------------------------------------------------------------------------------------------------------------
# Observation shards (BASELINE config 5): ONE chain over all ranks, every rank holds a contiguous row range,
# the leaf statistics of a batch of trees are exchanged inside the sweep kernel through peer memory over NVLink.
# ------------------------------------------------------------------------------------------------------------------
def sharded_record(args, torch, dist, rank, local_rank, world, n, p, K, W, burnin, engine):
    """Runs on every rank; rank 0 gets the record.  (1) bit-parity of the N-shard chain with the same chain on one GPU
    at a small n; (2) sweeps/s of the N-shard chain at (n, p), device-timed, max over ranks; (3) the same rows per GPU
    run WITHOUT the exchange (each rank an independent chain on its rows) -- the difference is what the exchange costs."""
    from bcf_iv_b200 import _lib, multi, workload
    dev = torch.device("cuda", local_rank)
    T = T_CON + T_MOD

    def maxr(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def shard_sampler(lam, sig0, seed):
        s = _lib.Sampler(engine=engine, lambda_=lam, sigma_init=sig0, seed=seed, chain=0, device=local_rank)
        s.set_shard(rank, world, multi.exchange_unique_id(_lib.nccl_unique_id))
        return s

    # ---- (1) parity: decisions, sigma and the posterior mean of tau bit for bit, shards vs one GPU ----
    n_par, p_par, sweeps_par = args.shard_parity_n, 20, 12
    wp = workload.synthetic_sampler_inputs(n_par, p_par, seed=args.data_seed + 1)
    lo, hi = multi.shard_range(n_par, rank, world)
    s = shard_sampler(wp["lambda_"], wp["sigma_init"], args.seed)
    s.set_data(wp["y"][lo:hi], wp["z"][lo:hi], wp["x_con"][lo:hi], wp["x_mod"][lo:hi], wp["cuts_con"], wp["off_con"],
               wp["cuts_mod"], wp["off_mod"])
    s.run(0, sweeps_par)
    tau_sh = multi.gather_rows(s.tau_mean())
    tr_sh, lr_sh = s.trace()
    sc_sh, cnt_sh, eng_par = s.scalars(), s.counters(), s.engine_in_use
    s.close()
    parity = None
    if rank == 0:
        s1 = _lib.Sampler(lambda_=wp["lambda_"], sigma_init=wp["sigma_init"], seed=args.seed, chain=0, device=local_rank)
        s1.set_data(wp["y"], wp["z"], wp["x_con"], wp["x_mod"], wp["cuts_con"], wp["off_con"], wp["cuts_mod"], wp["off_mod"])
        s1.run(0, sweeps_par)
        tr1, lr1 = s1.trace()
        sc1 = s1.scalars()
        parity = {"n": n_par, "p": p_par, "sweeps": sweeps_par,
                  "moves_identical": bool(np.array_equal(tr1, tr_sh)) and s1.counters() == cnt_sh,
                  "log_ratios_identical": bool(np.array_equal(lr1, lr_sh, equal_nan=True)),
                  "scalars_identical": all(sc1[k] == sc_sh[k] for k in ("sigma", "mscale", "bscale1", "bscale0", "tau_con")),
                  "tau_mean_identical": bool(np.array_equal(s1.tau_mean(), tau_sh)),
                  "exchange": {0: "none", 1: "ncclAllReduce per tree", 2: "in-kernel peer-memory mailboxes (NVLink)"}[eng_par[1]]}
        s1.close()
    # ---- (2) the timed shape ----
    def gather_sum(v):
        parts = [None] * world
        dist.all_gather_object(parts, np.asarray(v))
        return np.sum(parts, axis=0)                          # same order on every rank: identical bits
    lo, hi = multi.shard_range(n, rank, world)
    t0 = time.perf_counter()
    w = workload.synthetic_shard_inputs(n, p, lo, hi, seed=args.data_seed, allreduce=gather_sum)
    gen_s = time.perf_counter() - t0
    s = shard_sampler(w["lambda_"], w["sigma_init"], args.seed)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
    set_s = maxr(time.perf_counter() - t0)
    s.run(burnin, 0)
    s.run(W, 0)
    l0 = s.kernel_launches
    dist.barrier(); torch.cuda.synchronize()
    s.run(0, K)
    torch.cuda.synchronize()
    ms = maxr(s.last_run_ms)
    launches = s.kernel_launches - l0
    in_use = s.engine_in_use
    sizes = [len(s.get_tree(t)[0]) for t in range(0, T, 5)]
    sig = s.scalars()["sigma"]
    s.close()
    # ---- (3) the same rows per GPU without the exchange ----
    s = _lib.Sampler(engine=engine, lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=args.seed, chain=1 + rank, device=local_rank)
    s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
    s.run(burnin, 0)
    s.run(W, 0)
    dist.barrier(); torch.cuda.synchronize()
    s.run(0, K)
    ms_local = maxr(s.last_run_ms)
    eng_local = s.engine_in_use
    s.close()
    _lib.release_cached_memory()
    if rank != 0:
        return None
    value = K / (ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    b_sweep = algorithmic_bytes_per_sweep(n, T)
    achieved = b_sweep * value / 1e9
    names = {0: "graph", 1: "persistent_hbm", 2: "persistent", 3: "batched", 4: "pipelined", 5: "batched_hbm"}
    kern = {"pipelined": "k_pipe", "batched": "k_batched<resident>", "batched_hbm": "k_batched<hbm>", "persistent": "k_resident",
            "persistent_hbm": "k_persistent", "graph": "k_tree_step x T + ncclAllReduce per tree"}
    exch = {0: "none", 1: "ncclAllReduce per tree", 2: "in-kernel peer-memory mailboxes (NVLink)"}[in_use[1]]
    ok = bool(parity and parity["moves_identical"] and parity["log_ratios_identical"] and parity["scalars_identical"]
              and parity["tau_mean_identical"])
    return {
        "workload": f"ONE chain, observations sharded over {world} GPUs: n_s={n} rows in all ({hi - lo} per GPU), p={p} binary "
                    f"covariates + pihat, {T_CON} mu + {T_MOD} tau trees (BASELINE.json config 5 shape)",
        "n": n, "p": p, "n_gpus": world, "rows_per_gpu": hi - lo, "value": value, "unit": UNIT, "ms_per_step": ms / K,
        "steps": K, "warmup": W, "burnin_sweeps": burnin, "scaling": "strong",
        "engine": names[in_use[0]], "kernel": kern[names[in_use[0]]], "shard_exchange": exch,
        "exchange_detail": "per batch of trees: every rank stores its reduced leaf statistics and cross counts (<= 6 KB of "
                           "int64 words) into its mailbox on every peer over NVLink, releases a sequence number, and "
                           "waits for the peers'; one more exchange per sweep for the scale / sigma moments",
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                     "algorithmic_bytes_per_sweep": b_sweep, "peak_source": peak_src + f" x {world} GPUs"},
        "same_rows_without_exchange": {"value": K / (ms_local * 1e-3), "ms_per_step": ms_local / K, "engine": names[eng_local[0]],
                                       "note": "each GPU an independent chain on its own rows: what the shard costs per "
                                               "sweep with no exchange; the difference to ms_per_step is the exchange "
                                               "(stores + wait for the slowest rank)"},
        "exchange_ms_per_step": ms / K - ms_local / K,
        "shard_parity": ok, "parity": parity,
        "gpu_launches": int(launches), "mean_nodes_per_tree": float(np.mean(sizes)), "sigma": sig,
        "set_data_s": set_s, "data_generation_s": gen_s,
    }


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from bcf_iv_b200 import _lib, workload

    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    if _lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (libbcfgpu has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's version banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n, p, K, W = args.n, args.p, args.steps, args.warmup
    T = T_CON + T_MOD
    sharded = args.mode == "sharded" and world > 1
    w = workload.synthetic_sampler_inputs(n, p, seed=args.data_seed)
    if sharded:
        lo, hi = rank * n // world, (rank + 1) * n // world
        rows = slice(lo, hi)
    else:
        rows = slice(0, n)
    host = {k: w[k][rows] if k in ("y", "z", "x_con", "x_mod") else w[k] for k in w}
    # pinned host copies: what the e2e call is handed (R would hand pageable memory; pinned is the contract's case)
    keep = []
    for k in ("y", "z", "x_con", "x_mod"):
        a = host[k]
        if k.startswith("x_"):
            a = np.asfortranarray(a)
        host[k], t = pinned_like(a)
        keep.append(t)
    if not sharded and np.shares_memory(w["x_mod"], w["x_con"]):
        # BayesIV's call is bcf(y, z, x, x, pihat): x_moderate is the leading columns of x_control = [x | pihat]; handed
        # over as such (same buffer) the library bins and copies the matrix once
        host["x_mod"] = host["x_con"][:, :w["x_mod"].shape[1]]

    engine = {"auto": 0, "graph": 1, "persistent": 2, "persistent_hbm": 3, "batched": 4, "pipelined": 5}[args.engine]

    def make_sampler():
        s = _lib.Sampler(engine=engine, lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=args.seed,
                         chain=0 if sharded else rank, device=local_rank)
        if sharded:
            uid = [_lib.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            s.set_shard(rank, world, uid[0])
        return s

    def set_data(s):
        s.set_data(host["y"], host["z"], host["x_con"], host["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"],
                   w["off_mod"])

    # ---- device-resident leg: `value` ----
    s = make_sampler()
    set_data(s)
    s.run(args.burnin, 0)                       # untimed burn-in: tree sizes (and so the sweep cost) become stationary
    for _ in range(1):
        s.run(W, 0)                             # W warm-up sweeps
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.3)
    l0 = s.kernel_launches
    barrier()
    t0 = time.time()
    s.run(0, K)                                 # exactly K sweeps, every one saved into the posterior accumulators
    torch.cuda.synchronize()
    t1 = time.time()
    dev_ms = s.last_run_ms                      # CUDA events on the library's stream around the K sweeps
    barrier()
    launches = s.kernel_launches - l0
    time.sleep(0.25)
    clocks.stop()
    ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks((t1 - t0) * 1e3)
    total_launches = int(sum_over_ranks(launches))
    units = K * (1 if sharded else world)       # sweeps of all chains (sharded: one chain over all ranks)
    value = units / (ms * 1e-3)
    in_use = s.engine_in_use
    sizes = [len(s.get_tree(t)[0]) for t in range(0, T, 5)]
    acc = s.counters()
    sig = s.scalars()["sigma"]
    stats0 = s.engine_stats
    # ---- other regimes of the same shape (N = 1 only; extra keys of the line, not the headline) ----
    regimes = None
    if world == 1 and not args.no_regimes:
        peak_gbs, _ = measured_peak_gbs()

        def timed(smp, bx):
            smp.run(W, 0)
            torch.cuda.synchronize()
            smp.run(0, K)
            v = K / (smp.last_run_ms * 1e-3)
            nodes = [len(smp.get_tree(t)[0]) for t in range(T)]
            return {"value": v, "unit": UNIT, "ms_per_step": smp.last_run_ms / K, "steps": K,
                    "mean_nodes_per_tree": float(np.mean(nodes)), "max_nodes_in_a_tree": int(np.max(nodes)),
                    "engine_stats_since_create": smp.engine_stats,
                    "roofline_frac": algorithmic_bytes_per_sweep(n, T, bx) * v / 1e9 / peak_gbs}
        regimes = {}
        # (1) the chain far into burn-in: trees at their stationary sizes (the headline burns in args.burnin sweeps)
        s.run(args.long_burnin - args.burnin - W - K, 0)
        regimes[f"after_{args.long_burnin}_burnin_sweeps"] = dict(timed(s, 1), note="same chain, same data as the headline")
        # (3) K2' (SURVEY 8a-4), the all-candidate histogram over the same binned covariates: what bcf_iv's subgroup CART calls
        # once per tree level; for two-valued columns an exact int8 contraction on tcgen05 (k_hist_mma).  Device time of the
        # kernel (events inside the library), algorithmic bytes n * (9 + p); not part of the sweeps/s figures.
        rr = np.random.default_rng(args.data_seed + 7).standard_normal(n)
        k2 = {}
        for nleaf in (1, 4, 8):
            slot = (np.arange(n) % nleaf).astype(np.int32)
            us = min(s.op_hist_all(1, slot, nleaf, rr)[2] for _ in range(3)) * 1e3
            gbs = n * (9 + p) / 1e9 / (us * 1e-6)
            k2[f"leaves_{nleaf}"] = {"us": us, "GBps": gbs, "roofline_frac": gbs / peak_gbs}
        regimes["k2prime_all_candidate_histogram"] = dict(k2, kernel="k_hist_mma (tcgen05.mma.kind::i8, TMA, tensor memory)",
                                                          algorithmic_bytes=n * (9 + p), columns=p)
    s.close()
    if regimes is not None:
        # (2) SURVEY 8d's stress variant: continuous covariates -> 10 000-cutpoint grids, uint16 bin columns
        wc = workload.synthetic_sampler_inputs(n, p, seed=args.data_seed, continuous=True, analytic_cuts=True)
        sc_ = _lib.Sampler(engine=engine, lambda_=wc["lambda_"], sigma_init=wc["sigma_init"], seed=args.seed, device=local_rank)
        sc_.set_data(wc["y"], wc["z"], wc["x_con"], wc["x_mod"], wc["cuts_con"], wc["off_con"], wc["cuts_mod"], wc["off_mod"])
        sc_.run(args.burnin, 0)
        regimes["continuous_x"] = dict(timed(sc_, 2), burnin_sweeps=args.burnin,
                                       note=f"X = Phi(raw) on {p - 2} of the {p} covariates: 10 000 cutpoints each, uint16 bins; "
                                            "roofline_frac on SURVEY 8d's b_x = 2 bytes")
        sc_.close()
        del wc

    # ---- end-to-end leg through the C-ABI with host buffers ----
    # The call a user makes: create + set_data (H2D of y, z, X from pinned host memory, binning) + run(K) + colMeans(tau)
    # back to the host.  Timed twice: `cold` right after the library's cached device blocks were released (every buffer
    # comes from cudaMalloc), then again in the state every fit but a process's first one sees (bcf_iv() makes two fits
    # per call, bcf(n_chains) many): blocks of the previous fit are reused.  `e2e.value` is the second; both are reported.
    def e2e_once():
        barrier()
        a = time.perf_counter()
        s2 = make_sampler()
        set_data(s2)
        s2.run(0, K)
        tau_ = s2.tau_mean()
        torch.cuda.synchronize()
        b_ = time.perf_counter()
        h2d_ = s2.h2d_bytes
        s2.close()
        return max_over_ranks(b_ - a), tau_, h2d_

    _lib.release_cached_memory()
    e2e_cold_s, _, _ = e2e_once()
    e2e_s, tau, h2d = e2e_once()
    d2h = tau.nbytes
    e2e_val = units / e2e_s
    if not np.all(np.isfinite(tau)):
        raise SystemExit("bench.py: non-finite posterior mean")

    shard_rec = None
    if world > 1 and not sharded and not args.no_sharded:
        host.clear(); keep.clear()                  # the pinned copies of the chains legs (0.8 GB per rank)
        shard_rec = sharded_record(args, torch, dist, rank, local_rank, world, args.shard_n, args.shard_p,
                                   min(K, args.shard_steps), min(W, 5), args.burnin, engine)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    n_local = n if not sharded else (hi - lo)
    b_sweep = algorithmic_bytes_per_sweep(n_local, T)
    peak, peak_src = measured_peak_gbs()
    sweeps_per_launch = K / max(launches, 1)
    launch_ms = dev_ms / max(launches, 1)
    achieved = b_sweep * sweeps_per_launch / (launch_ms * 1e-3) / 1e9
    eng_name = {0: "graph", 1: "persistent_hbm", 2: "persistent", 3: "batched", 4: "pipelined", 5: "batched_hbm"}[in_use[0]]
    exchange = {0: "none", 1: "ncclAllReduce per tree", 2: "in-kernel peer-memory mailboxes (NVLink)"}[in_use[1]]
    traffic_sweep = ncu_traffic_per_sweep(n_local, eng_name)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
        "vs_baseline": None, "dtype": "f64 (fixed-point int64 residual sums, exact)", "data": "synthetic",
        "config": dict(workload_config(args, world, "B200"), engine=eng_name, shard_exchange=exchange, burnin_sweeps=args.burnin,
                       l2="no flush: per-sweep working set (250 MB leaf-slot cache + 101 MB bins + 40 MB state) "
                          "exceeds the 126 MB L2",
                       mean_nodes_per_tree=float(np.mean(sizes)), sigma=sig, moves=acc, cpu_binding=numa_cpus,
                       wall_ms_per_step=wall_ms / K),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic_sweep * sweeps_per_launch if traffic_sweep is not None else None),
                     "peak_source": peak_src, "kernel": {"pipelined": "k_pipe", "batched": "k_batched<resident>", "batched_hbm": "k_batched<hbm>", "persistent": "k_resident", "persistent_hbm": "k_persistent", "graph": "k_tree_step x T (CUDA graph)"}.get(eng_name, eng_name),
                     "algorithmic_bytes_per_sweep": b_sweep, "sweeps_per_launch": sweeps_per_launch,
                     "launch_ms": launch_ms},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                "call": "bcfgpu_create + set_data(pinned host y, z, X_control, X_moderate: H2D + binning) + "
                        f"run({K} sweeps) + get_tau_mean (D2H), wall clock; bytes amortised over the {K} sweeps; device "
                        "blocks reused from the library's caching allocator (second fit of the process); `cold`: the same "
                        "call with the cache emptied first (all buffers from cudaMalloc)",
                "seconds": e2e_s, "cold": {"value": units / e2e_cold_s, "seconds": e2e_cold_s}},
        "gpu_launches": total_launches,
        "clocks": clocks.summary(t0, t1),
    }
    if shard_rec is not None:
        line["sharded"] = shard_rec
    if regimes is not None:
        line["regimes"] = regimes
    line["config"]["engine_stats"] = stats0
    if world == 1 and not args.no_cpu_baseline:
        from oracle import orc
        orc.build()
        nth = host_threads()
        n_s = min(n, args.cpu_sample)
        val, dt = time_oracle(w, n, n_s, args.cpu_sweeps, 1, nth)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": nth, "kind": "port",
                                "sample": f"{args.cpu_sweeps} timed sweeps (+1 warm-up) of the oracle (CPU restatement of "
                                          f"bcf's sweep) on the first {n_s} rows of the same workload, {nth} OpenMP "
                                          f"threads over observations; scaled by n_sample/n ({dt:.1f} s wall)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="chains", choices=["chains", "sharded"])
    ap.add_argument("--engine", default="auto", choices=["auto", "graph", "persistent", "persistent_hbm", "batched", "pipelined"])
    ap.add_argument("--nobs", "--n", dest="n", type=int, default=1_000_000)     # (--n / --p clash with torchrun abbreviations)
    ap.add_argument("--ncov", "--p", dest="p", type=int, default=100)
    ap.add_argument("--burnin", type=int, default=100)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--data-seed", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--cpu-sweeps", type=int, default=8)
    ap.add_argument("--cpu-budget", type=float, default=120.0, help="seconds of CPU work for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-regimes", action="store_true", help="N = 1: skip the long-burn-in and continuous-X records")
    ap.add_argument("--long-burnin", type=int, default=2000)
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the observation-sharded record")
    ap.add_argument("--shard-n", type=int, default=10_000_000, help="rows of the observation-sharded record (config 5)")
    ap.add_argument("--shard-p", type=int, default=200)
    ap.add_argument("--shard-steps", type=int, default=50)
    ap.add_argument("--shard-parity-n", type=int, default=60_000)
    args = ap.parse_args()
    if args.steps < 1 or args.warmup < 0:
        raise SystemExit("need --steps >= 1 and --warmup >= 0")
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
