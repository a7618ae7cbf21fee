"""Wall time of the dense pre-steps on the device at the named shapes, next to the CPU restatement where it
finishes in seconds.  One JSON line per shape."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bcf_iv_b200 import _lib  # noqa: E402


def main():
    shapes = [(100_000, 50), (500_000, 100), (1_000_000, 100), (1_000_000, 200)]
    rng = np.random.default_rng(0)
    for n, p in shapes:
        X = (rng.random((n, p)) < 0.5).astype(np.float64)
        b = rng.normal(size=p) * 0.1
        z = (rng.random(n) < 1.0 / (1.0 + np.exp(-(X @ b - 0.3)))).astype(np.int32)
        y = X[:, 0] + 0.5 * z + rng.standard_normal(n)
        Xf = np.asfortranarray(X)
        rec = {"n": n, "p": p}
        for name, M in (("row_major", X), ("col_major", Xf)):
            _lib.glm_logit(z, M)
            t = []
            for _ in range(3):
                t0 = time.perf_counter(); eta, info = _lib.glm_logit(z, M, full=True); t.append(time.perf_counter() - t0)
            rec[f"glm_logit_{name}_s"] = round(min(t), 4)
            rec["glm_iterations"] = info["iterations"]
            t = []
            for _ in range(3):
                t0 = time.perf_counter(); _lib.lm_sigma(y, z, M); t.append(time.perf_counter() - t0)
            rec[f"lm_sigma_{name}_s"] = round(min(t), 4)
        if n <= 500_000:
            # the host path this replaces (NumPy / LAPACK normal equations, what round 1 shipped)
            t0 = time.perf_counter()
            A = np.column_stack([np.ones(n), X]); eta = np.zeros(n)
            for _ in range(info["iterations"]):
                mu = 1 / (1 + np.exp(-eta)); w = np.clip(mu * (1 - mu), 1e-10, None)
                AtW = A.T * w
                beta = np.linalg.solve(AtW @ A, AtW @ (eta + (z - mu) / w)); eta = A @ beta
            rec["numpy_irls_s"] = round(time.perf_counter() - t0, 3)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
