"""ctypes binding of libbcfgpu.so -- exactly the entry points declared in include/bcfgpu.h.

This is the Python stand-in for the R shim (bcf_iv_b200/R/BayesIVgpu/src/r_shim.c): same C-ABI, host buffers in, host
buffers out.  There is no CPU fallback: if the library is missing it is an ImportError-like failure, and if
there is no CUDA device every compute call raises BcfGpuError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbcfgpu.so")

MAX_LEAVES = 128
STATUS = {0: "OK", 1: "BAD_ARG", 2: "CUDA", 3: "NCCL", 4: "NUMERIC", 5: "OOM", 6: "STATE", 7: "INTERRUPTED"}
ENGINE_AUTO, ENGINE_GRAPH, ENGINE_PERSISTENT, ENGINE_PERSISTENT_HBM, ENGINE_BATCHED, ENGINE_PIPELINED = 0, 1, 2, 3, 4, 5
MODEL_BCF, MODEL_PROBIT_BART = 0, 1

# every symbol include/bcfgpu.h declares (tests check the library exports each of them)
SYMBOLS = [
    "bcfgpu_abi_version", "bcfgpu_device_count", "bcfgpu_last_error", "bcfgpu_config_init", "bcfgpu_create",
    "bcfgpu_destroy", "bcfgpu_release_cached_memory", "bcfgpu_cache_stats", "bcfgpu_nccl_unique_id", "bcfgpu_set_shard", "bcfgpu_set_shard_local", "bcfgpu_set_data", "bcfgpu_run", "bcfgpu_request_stop",
    "bcfgpu_fit_chains",
    "bcfgpu_last_run_ms", "bcfgpu_kernel_launches", "bcfgpu_h2d_bytes", "bcfgpu_engine_in_use", "bcfgpu_engine_stats", "bcfgpu_num_saved", "bcfgpu_get_tau_mean", "bcfgpu_get_tau_var",
    "bcfgpu_get_mu_mean", "bcfgpu_get_yhat_mean", "bcfgpu_get_sigma", "bcfgpu_get_scales", "bcfgpu_get_draws",
    "bcfgpu_num_saved_forests", "bcfgpu_predict", "bcfgpu_get_scalars", "bcfgpu_get_fits", "bcfgpu_get_counters", "bcfgpu_get_trace", "bcfgpu_tree_size",
    "bcfgpu_get_tree", "bcfgpu_set_tree", "bcfgpu_get_leaf_ids", "bcfgpu_verify_leaf_cache", "bcfgpu_op_bin_column",
    "bcfgpu_op_suffstats", "bcfgpu_op_hist_all", "bcfgpu_op_lil", "bcfgpu_op_rng_probe", "bcfgpu_op_probit_latent",
    "bcfgpu_glm_logit", "bcfgpu_lm_sigma", "bcfgpu_scan_columns",
]


class BcfGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbcfgpu: {STATUS.get(code, code)}: {msg}")
        self.code = code


B200_SMS = 148          # the library is written for B200 (sm_100a) only


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("ntree_con", C.c_int32), ("ntree_mod", C.c_int32),
        ("alpha_con", C.c_double), ("beta_con", C.c_double), ("alpha_mod", C.c_double), ("beta_mod", C.c_double),
        ("con_sd", C.c_double), ("mod_sd", C.c_double), ("nu", C.c_double), ("lambda_", C.c_double),
        ("sigma_init", C.c_double), ("use_muscale", C.c_int32), ("use_tauscale", C.c_int32),
        ("min_leaf", C.c_int32), ("max_leaves", C.c_int32), ("seed", C.c_uint64), ("chain", C.c_uint32),
        ("device", C.c_int32), ("engine", C.c_int32), ("keep_draws", C.c_int32), ("max_ctas", C.c_int32),
        ("model", C.c_int32), ("treatment_col", C.c_int32), ("reserved0", C.c_int32), ("probit_offset", C.c_double),
        ("reserved", C.c_int32 * 2),
    ]


POLL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)    # bcfgpu_fit_chains' interrupt hook

_lib = None


def load():
    """dlopen libbcfgpu.so (building it first if nvcc is there and it is missing).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build
        build.build_lib()
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} is missing: build it with `python -m bcf_iv_b200.build` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int32); up = C.POINTER(C.c_uint32); lp = C.POINTER(C.c_int64)
    vp = C.c_void_p
    L.bcfgpu_abi_version.restype = C.c_int
    L.bcfgpu_device_count.restype = C.c_int
    L.bcfgpu_last_error.restype = C.c_char_p
    L.bcfgpu_config_init.argtypes = [C.POINTER(Config)]
    L.bcfgpu_config_init.restype = None
    L.bcfgpu_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.bcfgpu_destroy.argtypes = [vp]
    L.bcfgpu_release_cached_memory.argtypes = []
    L.bcfgpu_cache_stats.argtypes = [lp]
    L.bcfgpu_nccl_unique_id.argtypes = [C.c_char_p]
    L.bcfgpu_set_shard.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.bcfgpu_set_shard_local.argtypes = [C.POINTER(vp), C.c_int]
    L.bcfgpu_set_data.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, dp, ip, dp, dp, dp, ip, dp, ip]
    L.bcfgpu_run.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    L.bcfgpu_request_stop.argtypes = [vp]
    L.bcfgpu_fit_chains.argtypes = [C.POINTER(Config), C.c_int32, ip, C.c_int32, C.c_int64, C.c_int32, C.c_int32, dp, ip, dp,
                                    dp, dp, ip, dp, ip, C.c_int32, C.c_int32, C.c_int32, POLL_FN, vp, C.POINTER(vp)]
    L.bcfgpu_last_run_ms.argtypes = [vp, dp]
    L.bcfgpu_kernel_launches.argtypes = [vp, lp]
    L.bcfgpu_h2d_bytes.argtypes = [vp, lp]
    L.bcfgpu_engine_in_use.argtypes = [vp, ip]
    L.bcfgpu_engine_stats.argtypes = [vp, lp]
    L.bcfgpu_num_saved.argtypes = [vp, ip]
    for nm in ("tau_mean", "tau_var", "mu_mean", "yhat_mean", "sigma"):
        getattr(L, "bcfgpu_get_" + nm).argtypes = [vp, dp]
    L.bcfgpu_get_scales.argtypes = [vp, dp, dp]
    L.bcfgpu_get_draws.argtypes = [vp, C.c_int32, dp]
    L.bcfgpu_num_saved_forests.argtypes = [vp, ip, lp]
    L.bcfgpu_predict.argtypes = [vp, C.c_int64, dp, dp, dp, dp, dp, dp, dp]
    L.bcfgpu_get_scalars.argtypes = [vp, dp]
    L.bcfgpu_get_fits.argtypes = [vp, dp, dp]
    L.bcfgpu_get_counters.argtypes = [vp, lp]
    L.bcfgpu_get_trace.argtypes = [vp, ip, dp]
    L.bcfgpu_tree_size.argtypes = [vp, C.c_int32, ip]
    L.bcfgpu_get_tree.argtypes = [vp, C.c_int32, up, ip, ip, dp]
    L.bcfgpu_set_tree.argtypes = [vp, C.c_int32, C.c_int32, up, ip, ip, dp]
    L.bcfgpu_get_leaf_ids.argtypes = [vp, C.c_int32, up]
    L.bcfgpu_verify_leaf_cache.argtypes = [vp, lp]
    L.bcfgpu_op_bin_column.argtypes = [dp, C.c_int64, dp, C.c_int32, C.POINTER(C.c_uint16)]
    L.bcfgpu_op_suffstats.argtypes = [vp, C.c_int32, dp, C.c_uint32, C.c_int32, C.c_int32, ip, up, lp, lp, dp, dp, dp]
    L.bcfgpu_op_hist_all.argtypes = [vp, C.c_int32, ip, C.c_int32, dp, lp, dp, dp]
    L.bcfgpu_op_lil.argtypes = [dp, dp, dp, C.c_int32, dp]
    L.bcfgpu_op_rng_probe.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp]
    L.bcfgpu_op_probit_latent.argtypes = [dp, ip, dp, C.c_int32, dp]
    L.bcfgpu_glm_logit.argtypes = [C.c_int32, C.c_int64, C.c_int32, dp, C.c_int32, ip, C.c_int32, C.c_double, dp, dp, dp, ip]
    L.bcfgpu_lm_sigma.argtypes = [C.c_int32, C.c_int64, C.c_int32, dp, C.c_int32, ip, dp, dp, dp, ip]
    L.bcfgpu_scan_columns.argtypes = [dp, C.c_int64, C.c_int32, C.c_int64, dp, dp, C.POINTER(C.c_uint8)]
    for nm in SYMBOLS:
        f = getattr(L, nm)
        if nm not in ("bcfgpu_last_error", "bcfgpu_config_init"):
            f.restype = C.c_int
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def check(code):
    if code != 0:
        raise BcfGpuError(code, load().bcfgpu_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return load().bcfgpu_device_count()


def default_config() -> Config:
    c = Config()
    load().bcfgpu_config_init(C.byref(c))
    return c


def release_cached_memory():
    check(load().bcfgpu_release_cached_memory())


def cache_stats():
    o = np.zeros(2, dtype=np.int64)
    check(load().bcfgpu_cache_stats(_lp(o)))
    return {"device_bytes": int(o[0]), "pinned_bytes": int(o[1])}


def set_shard_local(samplers):
    """make the samplers (one per shard, same process) the shards of one chain; then drive each from its own thread"""
    hs = (C.c_void_p * len(samplers))(*[s._h for s in samplers])
    check(load().bcfgpu_set_shard_local(hs, len(samplers)))


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(load().bcfgpu_nccl_unique_id(buf))
    return buf.raw


class Sampler:
    """Thin object wrapper over a bcfgpu_handle."""

    def __init__(self, **kw):
        L = load()
        c = default_config()
        for k, v in kw.items():
            if k == "lambda":
                k = "lambda_"
            if not hasattr(c, k):
                raise TypeError(f"unknown config field {k}")
            setattr(c, k, v)
        self.cfg = c
        self.T = c.ntree_con + c.ntree_mod
        self._h = C.c_void_p()
        check(L.bcfgpu_create(C.byref(c), C.byref(self._h)))
        self.n = 0
        self._keep = []

    @classmethod
    def _adopt(cls, handle, cfg, n, off_con, off_mod, p_con, p_mod, cuts_con=None, cuts_mod=None):
        """wrap a handle made by the library (bcfgpu_fit_chains)"""
        s = cls.__new__(cls)
        s.cfg, s.T, s._h, s.n, s._keep = cfg, cfg.ntree_con + cfg.ntree_mod, C.c_void_p(handle), n, []
        s.off_con, s.off_mod, s.p_con, s.p_mod = off_con, off_mod, p_con, p_mod
        s.cuts_con, s.cuts_mod = cuts_con, cuts_mod
        return s

    def cutpoints(self, forest):
        """the flat cutpoint vector the forest's columns were binned with (offsets: off_con / off_mod)"""
        return self.cuts_con if forest == 0 else self.cuts_mod

    def request_stop(self):
        """thread-safe: the bcfgpu_run in progress on another thread returns INTERRUPTED at its next launch boundary"""
        check(load().bcfgpu_request_stop(self._h))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().bcfgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, rank, nranks, unique_id: bytes | None):
        check(load().bcfgpu_set_shard(self._h, rank, nranks, unique_id))

    def set_data(self, y, z, x_con, x_mod, cuts_con, off_con, cuts_mod=None, off_mod=None):
        """x_con / x_mod: n x p arrays (any layout; converted to column-major float64 without copying when they
        already are Fortran-ordered)."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.int32)
        xc = np.asfortranarray(x_con, dtype=np.float64)
        n = y.size
        if self.cfg.ntree_mod > 0:
            xm = np.asfortranarray(x_mod, dtype=np.float64)
            cm = np.ascontiguousarray(cuts_mod, dtype=np.float64); om = np.ascontiguousarray(off_mod, dtype=np.int32)
            p_mod = xm.shape[1]
        else:
            xm = cm = om = None
            p_mod = 0
        cc = np.ascontiguousarray(cuts_con, dtype=np.float64); oc = np.ascontiguousarray(off_con, dtype=np.int32)
        if xc.shape[0] != n or z.size != n or (xm is not None and xm.shape[0] != n):
            raise ValueError("row count mismatch")
        if oc.size != xc.shape[1] + 1 or (om is not None and om.size != p_mod + 1):
            raise ValueError("cut offsets must have p+1 entries")
        check(load().bcfgpu_set_data(self._h, n, xc.shape[1], p_mod, _dp(y), _ip(z), _dp(xc), _dp(xm), _dp(cc),
                                     _ip(oc), _dp(cm), _ip(om)))
        self.n = n
        self.off_con, self.off_mod = oc, om
        self.cuts_con, self.cuts_mod = cc, cm
        self.p_con, self.p_mod = xc.shape[1], p_mod
        v = C.c_int64()
        check(load().bcfgpu_h2d_bytes(self._h, C.byref(v)))
        self.h2d_bytes = v.value          # what the library copied (x_mod that IS x_con's leading columns is not copied twice)

    def run(self, nburn, nsim, nthin=1):
        check(load().bcfgpu_run(self._h, nburn, nsim, nthin))

    @property
    def last_run_ms(self):
        v = C.c_double()
        check(load().bcfgpu_last_run_ms(self._h, C.byref(v)))
        return v.value

    @property
    def kernel_launches(self):
        v = C.c_int64()
        check(load().bcfgpu_kernel_launches(self._h, C.byref(v)))
        return v.value

    @property
    def engine_in_use(self):
        """(sweep kernel, shard exchange) of the last run: kernel 0 graph, 1 k_persistent, 2 k_resident, 3 k_batched,
        4 k_pipe; exchange 0 none, 1 ncclAllReduce per tree, 2 in-kernel peer-memory mailboxes"""
        o = np.zeros(2, dtype=np.int32)
        check(load().bcfgpu_engine_in_use(self._h, _ip(o)))
        return int(o[0]), int(o[1])

    @property
    def engine_stats(self):
        """sweeps run by the pipelined kernel / handed to the batched kernel (a tree wider than 8 keys), since create"""
        o = np.zeros(2, dtype=np.int64)
        check(load().bcfgpu_engine_stats(self._h, _lp(o)))
        return {"pipelined_sweeps": int(o[0]), "fallback_sweeps": int(o[1])}

    @property
    def num_saved(self):
        v = C.c_int32()
        check(load().bcfgpu_num_saved(self._h, C.byref(v)))
        return v.value

    def _vec(self, name):
        o = np.zeros(self.n)
        check(getattr(load(), "bcfgpu_get_" + name)(self._h, _dp(o)))
        return o

    def tau_mean(self):
        return self._vec("tau_mean")

    def tau_var(self):
        return self._vec("tau_var")

    def mu_mean(self):
        return self._vec("mu_mean")

    def yhat_mean(self):
        return self._vec("yhat_mean")

    def sigma(self):
        o = np.zeros(max(self.num_saved, 1))
        check(load().bcfgpu_get_sigma(self._h, _dp(o)))
        return o[: self.num_saved]

    def scales(self):
        a = np.zeros(max(self.num_saved, 1)); b = np.zeros(max(self.num_saved, 1))
        check(load().bcfgpu_get_scales(self._h, _dp(a), _dp(b)))
        return a[: self.num_saved], b[: self.num_saved]

    def draws(self, which):
        idx = {"tau": 0, "mu": 1, "yhat": 2}[which]
        o = np.zeros((self.num_saved, self.n))
        check(load().bcfgpu_get_draws(self._h, idx, _dp(o)))
        return o

    @property
    def saved_forests(self):
        """(draws whose forests are kept on the device, nodes they hold)"""
        a = C.c_int32(); b = C.c_int64()
        check(load().bcfgpu_num_saved_forests(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def predict(self, x_con_new, x_mod_new=None, draws=False):
        """tau(x), mu(x) of the saved forests at new rows (standardised scale): dict(tau_mean, tau_var, mu_mean[, tau, mu])"""
        xc = np.asfortranarray(x_con_new, dtype=np.float64)
        xm = np.asfortranarray(x_mod_new, dtype=np.float64) if x_mod_new is not None else None
        m = xc.shape[0]
        if xc.shape[1] != self.p_con or (self.p_mod > 0 and (xm is None or xm.shape != (m, self.p_mod))):
            raise ValueError("new rows must have the training columns")
        out = {k: np.zeros(m) for k in ("tau_mean", "tau_var", "mu_mean")}
        td = md = None
        if draws:
            ns = self.saved_forests[0]
            td = np.zeros((ns, m)); md = np.zeros((ns, m))
            out.update(tau=td, mu=md)
        check(load().bcfgpu_predict(self._h, m, _dp(xc), _dp(xm), _dp(out["tau_mean"]), _dp(out["tau_var"]), _dp(out["mu_mean"]),
                                    _dp(td), _dp(md)))
        return out

    def scalars(self):
        o = np.zeros(8)
        check(load().bcfgpu_get_scalars(self._h, _dp(o)))
        return dict(sigma=o[0], mscale=o[1], bscale1=o[2], bscale0=o[3], delta_con=o[4], tau_con=o[5], tau_mod=o[6],
                    sweep=int(o[7]))

    def fits(self):
        a = np.zeros(self.n); b = np.zeros(self.n)
        check(load().bcfgpu_get_fits(self._h, _dp(a), _dp(b)))
        return a, b

    def counters(self):
        o = np.zeros(4, dtype=np.int64)
        check(load().bcfgpu_get_counters(self._h, _lp(o)))
        return dict(birth_prop=int(o[0]), birth_acc=int(o[1]), death_prop=int(o[2]), death_acc=int(o[3]))

    def trace(self):
        t = np.zeros((self.T, 5), dtype=np.int32); lr = np.zeros(self.T)
        check(load().bcfgpu_get_trace(self._h, _ip(t), _dp(lr)))
        return t, lr

    def get_tree(self, tid):
        k = C.c_int32()
        check(load().bcfgpu_tree_size(self._h, tid, C.byref(k)))
        k = k.value
        heap = np.zeros(k, dtype=np.uint32); var = np.zeros(k, dtype=np.int32); cut = np.zeros(k, dtype=np.int32)
        mu = np.zeros(k)
        check(load().bcfgpu_get_tree(self._h, tid, _up(heap), _ip(var), _ip(cut), _dp(mu)))
        return heap, var, cut, mu

    def set_tree(self, tid, heap, var, cut, mu):
        heap = np.ascontiguousarray(heap, dtype=np.uint32); var = np.ascontiguousarray(var, dtype=np.int32)
        cut = np.ascontiguousarray(cut, dtype=np.int32); mu = np.ascontiguousarray(mu, dtype=np.float64)
        check(load().bcfgpu_set_tree(self._h, tid, len(heap), _up(heap), _ip(var), _ip(cut), _dp(mu)))

    def leaf_ids(self, tid):
        o = np.zeros(self.n, dtype=np.uint32)
        check(load().bcfgpu_get_leaf_ids(self._h, tid, _up(o)))
        return o

    def verify_leaf_cache(self):
        v = C.c_int64()
        check(load().bcfgpu_verify_leaf_cache(self._h, C.byref(v)))
        return v.value

    def op_suffstats(self, tid, r, birth_leaf_heap=0, var=0, cut=0):
        r = np.ascontiguousarray(r, dtype=np.float64)
        nl = C.c_int32()
        heap = np.zeros(MAX_LEAVES, dtype=np.uint32)
        ct = np.zeros(MAX_LEAVES, dtype=np.int64); cc = np.zeros(MAX_LEAVES, dtype=np.int64)
        st = np.zeros(MAX_LEAVES); sc = np.zeros(MAX_LEAVES); ob = np.zeros(8)
        check(load().bcfgpu_op_suffstats(self._h, tid, _dp(r), int(birth_leaf_heap), int(var), int(cut), C.byref(nl),
                                         _up(heap), _lp(ct), _lp(cc), _dp(st), _dp(sc), _dp(ob)))
        k = nl.value
        return dict(heap=heap[:k], cnt_trt=ct[:k], cnt_ctl=cc[:k], sum_trt=st[:k], sum_ctl=sc[:k], birth=ob)

    def op_hist_all(self, forest, slot, nleaf, r):
        off = self.off_con if forest == 0 else self.off_mod
        p = self.p_con if forest == 0 else self.p_mod
        nb = int(off[p]) + p
        slot = np.ascontiguousarray(slot, dtype=np.int32); r = np.ascontiguousarray(r, dtype=np.float64)
        cnt = np.zeros((nleaf, nb), dtype=np.int64); s = np.zeros((nleaf, nb)); ms = C.c_double()
        check(load().bcfgpu_op_hist_all(self._h, forest, _ip(slot), nleaf, _dp(r), _lp(cnt), _dp(s), C.byref(ms)))
        return cnt, s, ms.value


def op_bin_column(x, cuts):
    x = np.ascontiguousarray(x, dtype=np.float64); cuts = np.ascontiguousarray(cuts, dtype=np.float64)
    out = np.zeros(x.size, dtype=np.uint16)
    check(load().bcfgpu_op_bin_column(_dp(x), x.size, _dp(cuts), cuts.size, out.ctypes.data_as(C.POINTER(C.c_uint16))))
    return out


def op_lil(sum_phi, sum_phi_r, tau):
    a = np.ascontiguousarray(sum_phi, dtype=np.float64); b = np.ascontiguousarray(sum_phi_r, dtype=np.float64)
    t = np.ascontiguousarray(tau, dtype=np.float64); o = np.zeros(a.size)
    check(load().bcfgpu_op_lil(_dp(a), _dp(b), _dp(t), a.size, _dp(o)))
    return o


def op_rng_probe(seed, chain, sweep, tree, purpose, idx):
    o = np.zeros(3)
    check(load().bcfgpu_op_rng_probe(seed, chain, sweep, tree, purpose, idx, _dp(o)))
    return o


def fit_chains(cfg_kw, n_chains, devices, y, z, x_con, x_mod, cuts_con, off_con, cuts_mod, off_mod, nburn, nsim, nthin=1,
               poll=None):
    """bcfgpu_fit_chains: n_chains chains in one native call (one host thread per device inside the library).
    Returns a list of Sampler objects wrapping the finished chains (close them)."""
    L = load()
    c = default_config()
    for k, v in cfg_kw.items():
        if k == "lambda":
            k = "lambda_"
        if not hasattr(c, k):
            raise TypeError(f"unknown config field {k}")
        setattr(c, k, v)
    y = np.ascontiguousarray(y, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.int32)
    xc = np.asfortranarray(x_con, dtype=np.float64)
    n = y.size
    if c.ntree_mod > 0:
        xm = np.asfortranarray(x_mod, dtype=np.float64)
        cm = np.ascontiguousarray(cuts_mod, dtype=np.float64); om = np.ascontiguousarray(off_mod, dtype=np.int32)
        p_mod = xm.shape[1]
    else:
        xm = cm = om = None
        p_mod = 0
    cc = np.ascontiguousarray(cuts_con, dtype=np.float64); oc = np.ascontiguousarray(off_con, dtype=np.int32)
    if xc.shape[0] != n or z.size != n or (xm is not None and xm.shape[0] != n):
        raise ValueError("row count mismatch")
    dv = np.ascontiguousarray(devices if devices is not None else [], dtype=np.int32)
    hs = (C.c_void_p * n_chains)()
    cb = POLL_FN(lambda _ctx: int(bool(poll()))) if poll is not None else C.cast(None, POLL_FN)
    check(L.bcfgpu_fit_chains(C.byref(c), n_chains, _ip(dv) if dv.size else None, dv.size, n, xc.shape[1], p_mod, _dp(y),
                              _ip(z), _dp(xc), _dp(xm), _dp(cc), _ip(oc), _dp(cm), _ip(om), nburn, nsim, nthin, cb, None, hs))
    out = []
    for k in range(n_chains):
        ck = Config.from_buffer_copy(c)
        ck.chain = c.chain + k
        out.append(Sampler._adopt(hs[k], ck, n, oc, om, xc.shape[1], p_mod, cc, cm))
    return out


def op_probit_latent(m, y, u):
    m = np.ascontiguousarray(m, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.int32)
    u = np.ascontiguousarray(u, dtype=np.float64); o = np.zeros(m.size)
    check(load().bcfgpu_op_probit_latent(_dp(m), _ip(y), _dp(u), m.size, _dp(o)))
    return o


def _matrix_arg(X):
    """(array kept alive, row_major flag) of an n x p float64 matrix handed over without a host-side transpose"""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    if X.flags.f_contiguous:
        return X, 0
    return np.ascontiguousarray(X), 1


def glm_logit(z, X, device=0, max_iter=25, epsilon=1e-8, full=False):
    """stats::glm(z ~ X, family = binomial) + predict(type = "link") on the device (R/bcf_iv.R:72-75, R/bcf_itt.R:58-61):
    the linear predictor (log-odds) of every row.  full=True also returns coefficients (NaN = aliased), deviance, info."""
    Xa, rm = _matrix_arg(X)
    n, p = Xa.shape
    zi = np.ascontiguousarray(np.asarray(z).ravel(), dtype=np.int32)
    if zi.size != n:
        raise ValueError("z and X must have the same number of rows")
    eta = np.empty(n)
    beta = np.empty(p + 1)
    dev = np.zeros(1)
    info = np.zeros(4, dtype=np.int32)
    check(load().bcfgpu_glm_logit(device, n, p, _dp(Xa), rm, _ip(zi), max_iter, epsilon, _dp(eta), _dp(beta), _dp(dev), _ip(info)))
    if full:
        return eta, {"beta": beta, "deviance": float(dev[0]), "iterations": int(info[0]), "converged": bool(info[1]),
                     "rank": int(info[2]), "launches": int(info[3])}
    return eta


def scan_columns(X):
    """(lo, hi, flags) per column of a column-major float64 matrix in one threaded host pass (bcfgpu_scan_columns):
    flags bit 0 = two-valued or constant column, bit 1 = NaN / Inf present.  Any column stride >= n rows is taken as it is."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2 or X.shape[1] == 0 or X.shape[0] == 0:
        raise ValueError("scan_columns needs a non-empty matrix")
    n, p = X.shape
    if not (X.strides[0] == 8 and (p == 1 or (X.strides[1] % 8 == 0 and X.strides[1] >= 8 * n))):
        X = np.asfortranarray(X)
    ld = n if p == 1 else X.strides[1] // 8
    lo = np.empty(p); hi = np.empty(p); fl = np.zeros(p, dtype=np.uint8)
    check(load().bcfgpu_scan_columns(_dp(X), n, p, ld, _dp(lo), _dp(hi), fl.ctypes.data_as(C.POINTER(C.c_uint8))))
    return lo, hi, fl


def lm_sigma(y, z, X, device=0, full=False):
    """summary(lm(y ~ z + X))$sigma on the device (bcf's default sighat); z may be None."""
    Xa, rm = _matrix_arg(X)
    n, p = Xa.shape
    ya = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
    zi = None if z is None else np.ascontiguousarray(np.asarray(z).ravel(), dtype=np.int32)
    if ya.size != n or (zi is not None and zi.size != n):
        raise ValueError("y, z and X must have the same number of rows")
    sig = np.zeros(1)
    beta = np.empty(1 + (zi is not None) + p)
    info = np.zeros(2, dtype=np.int32)
    check(load().bcfgpu_lm_sigma(device, n, p, _dp(Xa), rm, _ip(zi), _dp(ya), _dp(sig), _dp(beta), _ip(info)))
    if full:
        return float(sig[0]), {"beta": beta, "rank": int(info[0]), "launches": int(info[1])}
    return float(sig[0])
