"""Probit BART (bartc's sampler) on a BASELINE config 4 shaped problem: sweeps/s of one chain and the share of each kernel.
usage: python scripts/bart_probe.py [n] [p] [sweeps] [engine]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcf_iv_b200 import _lib, bart, datasets, prep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ENGINE = int(sys.argv[4]) if len(sys.argv) > 4 else 0
d = datasets.generate_dataset_wide(n=n, p=p, rho=0.5, compliance=0.3, seed=1, binary_outcome=True)
X, z, y = d["X"], d["z"].astype(np.int32), d["y"]
xc = np.asfortranarray(np.column_stack([X, 0.5 + 0.01 * np.random.default_rng(0).standard_normal(n), z.astype(float)]))
cuts, off = prep.pack_cuts([bart.bart_cutpoints(xc[:, v]) for v in range(xc.shape[1])])
s = _lib.Sampler(engine=ENGINE, **bart.bart_config(seed=1, treatment_col=xc.shape[1] - 1))
t0 = time.perf_counter(); s.set_data(y, z, xc, None, cuts, off); t_set = time.perf_counter() - t0
s.run(K, 0)
ms_burn = s.last_run_ms
t0 = time.perf_counter(); s.run(0, K); wall = time.perf_counter() - t0
print(f"n={n} p={p}: set_data {t_set*1e3:.1f} ms; {K} burn-in sweeps {ms_burn:.1f} ms ({1e3*K/ms_burn:.0f} sweeps/s); "
      f"{K} saved sweeps {s.last_run_ms:.1f} ms device / {wall*1e3:.1f} ms wall ({1e3*K/s.last_run_ms:.0f} sweeps/s); "
      f"engine {s.engine_in_use}, leaves>8: {np.mean([(len(s.get_tree(t)[0]) + 1) // 2 > 7 for t in range(s.T)]):.2f} of trees, ITE mean {s.tau_mean().mean():.4f}, nodes/tree {np.mean([len(s.get_tree(t)[0]) for t in range(s.T)]):.2f}")
s.close()
