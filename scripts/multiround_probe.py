import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import gpu_sampler, make_problem
pr = make_problem(n=int(sys.argv[1]), p=8, seed=31)
res = {}
for eng in (5, 4):
    g = gpu_sampler(pr, engine=eng, seed=3, ntree_con=20, ntree_mod=8)
    g.run(4, 4)
    res[eng] = (g.trace()[0].tobytes(), g.tau_mean().tobytes(), g.engine_in_use)
    print(eng, g.engine_in_use, g.scalars()["sigma"], g.verify_leaf_cache(), flush=True)
print("identical", res[5][:2] == res[4][:2])
