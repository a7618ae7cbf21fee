import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import make_problem, gpu_sampler
from bcf_iv_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
pr = make_problem(n=n, p=10, seed=3)
g = gpu_sampler(pr, seed=7, engine=_lib.ENGINE_PIPELINED)
g.run(200, 0)
g.run(5, 0)
g.close()
