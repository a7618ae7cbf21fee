"""Repeatability stress: the same chain, many times, every engine; any run that differs from the first is a race."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcf_iv_b200 import _lib, workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
w = workload.synthetic_sampler_inputs(n, 100, seed=0)
ref = None
for engine in (5, 4, 2):
    bad = 0
    for rep in range(reps):
        s = _lib.Sampler(engine=engine, lambda_=w["lambda_"], sigma_init=w["sigma_init"], seed=1)
        s.set_data(w["y"], w["z"], w["x_con"], w["x_mod"], w["cuts_con"], w["off_con"], w["cuts_mod"], w["off_mod"])
        s.run(sweeps // 2, sweeps - sweeps // 2)
        got = (s.scalars()["sigma"], tuple(s.counters().values()), s.trace()[0].tobytes(), s.tau_mean().tobytes(), s.sigma().tobytes())
        mism = s.verify_leaf_cache()
        s.close()
        if ref is None: ref = got
        if got != ref or mism:
            bad += 1
            print(f"engine {engine} rep {rep}: DIFFERENT sigma {got[0]!r} vs {ref[0]!r} counters {got[1]} vs {ref[1]} cache mismatches {mism}", flush=True)
    print(f"engine {engine}: {bad} of {reps} runs differ", flush=True)


# probit BART (latent drawn inside k_pipe<true>, counterfactual pass in tree order) and K2' on the tensor cores: run to run, bit for bit
from bcf_iv_b200 import bart, prep
rng = np.random.default_rng(2)
nb_ = min(n, 300_000)
X = rng.integers(0, 2, (nb_, 30)).astype(np.float64)
zb = rng.integers(0, 2, nb_).astype(np.int32)
yb = (rng.random(nb_) < 0.3 + 0.3 * X[:, 0] * zb).astype(np.float64)
xc = np.asfortranarray(np.column_stack([X, zb.astype(np.float64)]))
cuts, off = prep.pack_cuts([bart.bart_cutpoints(xc[:, v]) for v in range(xc.shape[1])])
refb, bad = None, 0
for rep in range(reps):
    b = _lib.Sampler(**bart.bart_config(seed=4, treatment_col=30))
    b.set_data(yb, zb, xc, None, cuts, off)
    b.run(12, 8)
    slot = (np.arange(nb_) % 5).astype(np.int32)
    cnt, sm, _ = b.op_hist_all(0, slot, 5, rng.standard_normal(nb_) if rep < 0 else np.sin(np.arange(nb_) * 0.37))
    got = (b.trace()[0].tobytes(), b.tau_mean().tobytes(), b.fits()[0].tobytes(), cnt.tobytes(), sm.tobytes(), tuple(b.engine_stats.values()))
    b.close()
    if refb is None: refb = got
    if got != refb:
        bad += 1
        print(f"probit BART / K2' rep {rep}: DIFFERENT", [a == c for a, c in zip(got, refb)], flush=True)
print(f"probit BART + K2' (k_hist_mma): {bad} of {reps} runs differ; engine stats {refb[5]}", flush=True)
