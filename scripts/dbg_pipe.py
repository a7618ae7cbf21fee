import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import make_problem, gpu_sampler
pr = make_problem(n=500, p=10, seed=3)
for (tc, tm) in ((200, 50), (200, 0), (100, 0), (40, 0), (60, 20)):
    for k in (2, 4, 6, 8, 10, 12):
        a = gpu_sampler(pr, seed=11, engine=4, ntree_con=tc, ntree_mod=tm); b = gpu_sampler(pr, seed=11, engine=5, ntree_con=tc, ntree_mod=tm)
        a.run(k, 0); b.run(k, 0)
        ok = a.scalars() == b.scalars() and np.array_equal(a.trace()[0], b.trace()[0])
        sizes = max(len(b.get_tree(t)[0]) for t in range(b.T))
        print(f"T=({tc},{tm}) one launch of {k} sweeps: {'same' if ok else 'DIFFERENT'} max nodes {sizes}", flush=True)
        a.close(); b.close()
        if not ok: break
