"""Small chains through every engine, for compute-sanitizer (memcheck / racecheck): n = 700, 30 + 12 trees."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import gpu_sampler, make_problem  # noqa: E402

pr = make_problem(n=700, p=6, seed=5, continuous=True)
for engine in (5, 4, 2, 3, 1):
    s = gpu_sampler(pr, seed=2, engine=engine, ntree_con=30, ntree_mod=12)
    s.run(3, 2)
    print("engine", engine, s.engine_in_use, float(s.tau_mean().sum()), s.verify_leaf_cache(), flush=True)
    s.close()
