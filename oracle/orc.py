"""ctypes binding of the CPU oracle (oracle/bcf_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of bcf_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product package.  PARITY UNPINNED (the reference has no golden vectors).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbcf_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc -O2 -fopenmp)."""
    src = os.path.join(_HERE, "bcf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class OrcConfig(C.Structure):
    _fields_ = [
        ("ntree_con", C.c_int32), ("ntree_mod", C.c_int32),
        ("alpha_con", C.c_double), ("beta_con", C.c_double),
        ("alpha_mod", C.c_double), ("beta_mod", C.c_double),
        ("con_sd", C.c_double), ("mod_sd", C.c_double),
        ("nu", C.c_double), ("lambda_", C.c_double), ("sigma_init", C.c_double),
        ("use_muscale", C.c_int32), ("use_tauscale", C.c_int32),
        ("min_leaf", C.c_int32), ("max_leaves", C.c_int32),
        ("seed", C.c_uint64), ("chain", C.c_uint32),
        ("nthreads", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        up = C.POINTER(C.c_uint32)
        L.orc_config_size.restype = C.c_int
        assert L.orc_config_size() == C.sizeof(OrcConfig), "orc_config layout mismatch"
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig), C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, ip, dp, ip]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_sweeps.argtypes = [C.c_void_p, C.c_int]
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [dp] * 9
        L.orc_get_scalars.argtypes = [C.c_void_p, dp]
        L.orc_get_fits.argtypes = [C.c_void_p, dp, dp]
        L.orc_get_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_get_trace.argtypes = [C.c_void_p, ip, dp]
        L.orc_tree_size.restype = C.c_int
        L.orc_tree_size.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_tree.restype = C.c_int
        L.orc_get_tree.argtypes = [C.c_void_p, C.c_int, up, ip, ip, dp]
        L.orc_set_tree.restype = C.c_int
        L.orc_set_tree.argtypes = [C.c_void_p, C.c_int, C.c_int, up, ip, ip, dp]
        L.orc_refit.argtypes = [C.c_void_p]
        L.orc_assign_leaves.argtypes = [C.c_void_p, C.c_int, up]
        L.orc_allsuff.restype = C.c_int
        L.orc_allsuff.argtypes = [C.c_void_p, C.c_int, dp, dp, up, dp, dp, dp]
        L.orc_getsuff_birth.restype = C.c_int
        L.orc_getsuff_birth.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_int, dp, dp, dp]
        L.orc_mh_logratio.restype = C.c_int
        L.orc_mh_logratio.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, dp, dp, dp]
        L.orc_set_y.argtypes = [C.c_void_p, dp]
        L.orc_get_allfit.argtypes = [C.c_void_p, dp]
        u8 = C.POINTER(C.c_uint8)
        L.orc_probit_enable.argtypes = [C.c_void_p, u8, ip, C.c_double, C.c_int]
        L.orc_probit_run.restype = C.c_int
        L.orc_probit_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp]
        L.orc_probit_sweeps.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_y.argtypes = [C.c_void_p, dp]
        L.orc_probit_latent_probe.restype = C.c_double
        L.orc_probit_latent_probe.argtypes = [C.c_double, C.c_int, C.c_double]
        L.orc_run_predict.restype = C.c_int
        L.orc_run_predict.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.orc_lil.restype = C.c_double
        L.orc_lil.argtypes = [C.c_double] * 3
        L.orc_lil_homosked.restype = C.c_double
        L.orc_lil_homosked.argtypes = [C.c_double] * 5
        L.orc_hist_all.argtypes = [C.c_void_p, C.c_int, ip, C.c_int, dp, C.POINTER(C.c_int64), dp]
        L.orc_philox4x32_10.argtypes = [up, up, up]
        L.orc_rng_probe.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp]
        L.orc_rng_gamma_probe.restype = C.c_double
        L.orc_rng_gamma_probe.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_up(c), _up(k), _up(o))
    return o


def rng_probe(seed, chain, sweep, tree, purpose, idx):
    o = np.zeros(3)
    lib().orc_rng_probe(seed, chain, sweep, tree, purpose, idx, _dp(o))
    return o  # ua, ub, normal


def rng_gamma(seed, chain, sweep, purpose, a):
    return lib().orc_rng_gamma_probe(seed, chain, sweep, purpose, a)


def pack_cuts(cut_list):
    """list of per-variable cutpoint vectors -> (flat cuts, offsets[p+1])."""
    off = np.zeros(len(cut_list) + 1, dtype=np.int32)
    for v, c in enumerate(cut_list):
        off[v + 1] = off[v] + len(c)
    flat = np.concatenate([np.asarray(c, dtype=np.float64) for c in cut_list]) if len(cut_list) else np.zeros(0)
    return np.ascontiguousarray(flat), off


class OracleSampler:
    """The reference-shaped sampler.  Data must already be in bcf's internal order (treated rows first,
    SURVEY A.1 `perm = order(z, decreasing=TRUE)`); y on the standardised scale."""

    def __init__(self, y, ntrt, x_con, x_mod, cuts_con, cuts_mod, *, ntree_con=200, ntree_mod=50,
                 alpha_con=0.95, beta_con=2.0, alpha_mod=0.25, beta_mod=3.0, con_sd=2.0, mod_sd=1.0 / 0.674,
                 nu=3.0, lambda_=1.0, sigma_init=1.0, use_muscale=True, use_tauscale=True, min_leaf=5,
                 max_leaves=0, seed=42, chain=0, nthreads=1):
        L = lib()
        self.n = int(len(y))
        self.ntrt = int(ntrt)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.xc = np.ascontiguousarray(x_con, dtype=np.float64)  # n x p row-major
        self.xm = np.ascontiguousarray(x_mod, dtype=np.float64)
        self.p_con = self.xc.shape[1]
        self.p_mod = self.xm.shape[1]
        self.cc, self.oc = pack_cuts(cuts_con)
        self.cm, self.om = pack_cuts(cuts_mod)
        self.cfg = OrcConfig(ntree_con, ntree_mod, alpha_con, beta_con, alpha_mod, beta_mod, con_sd, mod_sd, nu,
                             lambda_, sigma_init, int(use_muscale), int(use_tauscale), min_leaf, max_leaves, seed,
                             chain, nthreads)
        self.T = ntree_con + ntree_mod
        self.h = L.orc_create(C.byref(self.cfg), self.n, self.ntrt, self.p_con, self.p_mod, _dp(self.y),
                              _dp(self.xc), _dp(self.xm), _dp(self.cc), _ip(self.oc), _dp(self.cm), _ip(self.om))

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sweeps(self, k=1):
        lib().orc_sweeps(self.h, int(k))

    def run(self, nburn, nsim, nthin=1, draws=False):
        n = self.n
        out = {"tau_mean": np.zeros(n), "mu_mean": np.zeros(n), "yhat_mean": np.zeros(n),
               "sigma": np.zeros(nsim), "msd": np.zeros(nsim), "bsd": np.zeros(nsim)}
        td = md = yd = None
        if draws:
            td = np.zeros((nsim, n)); md = np.zeros((nsim, n)); yd = np.zeros((nsim, n))
            out.update(tau=td, mu=md, yhat=yd)
        saved = lib().orc_run(self.h, nburn, nsim, nthin, _dp(td), _dp(md), _dp(yd), _dp(out["tau_mean"]),
                              _dp(out["mu_mean"]), _dp(out["yhat_mean"]), _dp(out["sigma"]), _dp(out["msd"]),
                              _dp(out["bsd"]))
        out["saved"] = saved
        return out

    def run_predict(self, nburn, nsim, nthin, x_con_new, x_mod_new, draws=False):
        """run the chain and evaluate every saved sweep's forests at new rows (row-major n_new x p)"""
        xc = np.ascontiguousarray(x_con_new, dtype=np.float64); xm = np.ascontiguousarray(x_mod_new, dtype=np.float64)
        m = xc.shape[0]
        out = {"tau_mean": np.zeros(m), "mu_mean": np.zeros(m)}
        td = np.zeros((nsim, m)) if draws else None
        out["saved"] = lib().orc_run_predict(self.h, int(nburn), int(nsim), int(nthin), m, _dp(xc), _dp(xm), _dp(out["tau_mean"]),
                                             _dp(out["mu_mean"]), _dp(td))
        if draws:
            out["tau"] = td
        return out

    def scalars(self):
        o = np.zeros(8)
        lib().orc_get_scalars(self.h, _dp(o))
        return dict(sigma=o[0], mscale=o[1], bscale1=o[2], bscale0=o[3], delta_con=o[4], tau_con=o[5],
                    tau_mod=o[6], sweep=int(o[7]))

    def fits(self):
        a = np.zeros(self.n); b = np.zeros(self.n)
        lib().orc_get_fits(self.h, _dp(a), _dp(b))
        return a, b

    def counters(self):
        o = np.zeros(4, dtype=np.int64)
        lib().orc_get_counters(self.h, o.ctypes.data_as(C.POINTER(C.c_int64)))
        return dict(birth_prop=int(o[0]), birth_acc=int(o[1]), death_prop=int(o[2]), death_acc=int(o[3]))

    def trace(self):
        t = np.zeros((self.T, 5), dtype=np.int32); lr = np.zeros(self.T)
        lib().orc_get_trace(self.h, _ip(t), _dp(lr))
        return t, lr

    def get_tree(self, tid):
        k = lib().orc_tree_size(self.h, tid)
        heap = np.zeros(k, dtype=np.uint32); var = np.zeros(k, dtype=np.int32); cut = np.zeros(k, dtype=np.int32)
        mu = np.zeros(k)
        lib().orc_get_tree(self.h, tid, _up(heap), _ip(var), _ip(cut), _dp(mu))
        return heap, var, cut, mu

    def set_tree(self, tid, heap, var, cut, mu):
        heap = np.ascontiguousarray(heap, dtype=np.uint32); var = np.ascontiguousarray(var, dtype=np.int32)
        cut = np.ascontiguousarray(cut, dtype=np.int32); mu = np.ascontiguousarray(mu, dtype=np.float64)
        rc = lib().orc_set_tree(self.h, tid, len(heap), _up(heap), _ip(var), _ip(cut), _dp(mu))
        if rc != 0:
            raise ValueError("bad serialised tree")

    def refit(self):
        lib().orc_refit(self.h)

    def assign_leaves(self, tid):
        o = np.zeros(self.n, dtype=np.uint32)
        lib().orc_assign_leaves(self.h, tid, _up(o))
        return o

    def allsuff(self, tid, phi, R):
        k = lib().orc_tree_size(self.h, tid)
        heap = np.zeros(k, dtype=np.uint32); n0 = np.zeros(k); sp = np.zeros(k); spr = np.zeros(k)
        phi = np.ascontiguousarray(phi, dtype=np.float64); R = np.ascontiguousarray(R, dtype=np.float64)
        nb = lib().orc_allsuff(self.h, tid, _dp(phi), _dp(R), _up(heap), _dp(n0), _dp(sp), _dp(spr))
        return heap[:nb], n0[:nb], sp[:nb], spr[:nb]

    def getsuff_birth(self, tid, leaf_heap, v, c, phi, R):
        o = np.zeros(6)
        phi = np.ascontiguousarray(phi, dtype=np.float64); R = np.ascontiguousarray(R, dtype=np.float64)
        rc = lib().orc_getsuff_birth(self.h, tid, int(leaf_heap), int(v), int(c), _dp(phi), _dp(R), _dp(o))
        if rc != 0:
            raise ValueError("no such leaf")
        return o

    def mh_logratio(self, tid, move, heap, v, c, phi, R):
        """(log alpha1, log likelihood ratio) of a named move on the tree as it stands: move 1 = birth at leaf `heap`
        under rule (v, c), move 2 = death of nog node `heap`"""
        o = np.zeros(2)
        phi = np.ascontiguousarray(phi, dtype=np.float64); R = np.ascontiguousarray(R, dtype=np.float64)
        if lib().orc_mh_logratio(self.h, tid, int(move), int(heap), int(v), int(c), _dp(phi), _dp(R), _dp(o)) != 0:
            raise ValueError("no such move")
        return o

    def set_y(self, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        assert y.size == self.n
        lib().orc_set_y(self.h, _dp(y))

    def allfit(self):
        o = np.zeros(self.n)
        lib().orc_get_allfit(self.h, _dp(o))
        return o

    # ---- probit BART (create with ntree_mod = 0, use_muscale = False, y = zeros) ----
    def probit_enable(self, ybin, rowid, offset=0.0, trt_col=-1):
        yb = np.ascontiguousarray(ybin, dtype=np.uint8); ri = np.ascontiguousarray(rowid, dtype=np.int32)
        assert yb.size == self.n and ri.size == self.n
        lib().orc_probit_enable(self.h, yb.ctypes.data_as(C.POINTER(C.c_uint8)), _ip(ri), float(offset), int(trt_col))

    def probit_sweeps(self, k=1):
        lib().orc_probit_sweeps(self.h, int(k))

    def probit_run(self, nburn, nsim, nthin=1):
        ite = np.zeros(self.n); prob = np.zeros(self.n)
        saved = lib().orc_probit_run(self.h, int(nburn), int(nsim), int(nthin), _dp(ite), _dp(prob))
        return dict(ite_mean=ite, prob_mean=prob, saved=saved)

    def get_y(self):
        o = np.zeros(self.n)
        lib().orc_get_y(self.h, _dp(o))
        return o

    def hist_all(self, forest, slot, nleaf, r):
        off = self.oc if forest == 0 else self.om
        p = self.p_con if forest == 0 else self.p_mod
        nb = int(off[p]) + p
        slot = np.ascontiguousarray(slot, dtype=np.int32); r = np.ascontiguousarray(r, dtype=np.float64)
        cnt = np.zeros((nleaf, nb), dtype=np.int64); s = np.zeros((nleaf, nb))
        lib().orc_hist_all(self.h, forest, _ip(slot), nleaf, _dp(r), cnt.ctypes.data_as(C.POINTER(C.c_int64)), _dp(s))
        return cnt, s


def lil(sum_phi, sum_phi_r, tau):
    return lib().orc_lil(sum_phi, sum_phi_r, tau)


def lil_homosked(n, sy, sy2, sigma, tau):
    return lib().orc_lil_homosked(n, sy, sy2, sigma, tau)


def max_threads():
    return lib().orc_max_threads()


def probit_latent(m, y, u):
    return lib().orc_probit_latent_probe(float(m), int(y), float(u))
