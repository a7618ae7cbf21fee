"""The drop-in boundary: libbcfgpu.so loads, exports every symbol include/bcfgpu.h declares, validates its
arguments, and has NO CPU fallback (without a CUDA device every compute entry point fails with BCFGPU_ERR_CUDA)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bcf_iv_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "bcfgpu.h")).read()
    return sorted(set(re.findall(r"BCFGPU_API\s+[\w\s\*]+?\b(bcfgpu_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    decl = _declared_symbols()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/bcfgpu.h but not exported"
    assert sorted(_lib.SYMBOLS) == decl, "ctypes binding and header disagree"


def test_nothing_else_is_exported():
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert syms and all(s.startswith("bcfgpu_") for s in syms), syms


def test_config_defaults_are_bcf_defaults():
    c = _lib.default_config()
    assert c.struct_size == C.sizeof(_lib.Config)
    assert (c.ntree_con, c.ntree_mod) == (200, 50)
    assert (c.alpha_con, c.beta_con, c.alpha_mod, c.beta_mod) == (0.95, 2.0, 0.25, 3.0)
    assert c.con_sd == 2.0 and c.mod_sd == pytest.approx(1 / 0.674)
    assert c.nu == 3.0 and c.min_leaf == 5 and c.use_muscale == 1 and c.use_tauscale == 1
    assert _lib.load().bcfgpu_abi_version() == 1


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(_lib.BcfGpuError) as e:
        _lib.Sampler()
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.BcfGpuError) as e:
        _lib.op_bin_column(np.zeros(4), np.array([0.5]))
    assert e.value.code == 2
    with pytest.raises(_lib.BcfGpuError):
        _lib.op_lil(np.ones(2), np.ones(2), np.ones(2))
    with pytest.raises(_lib.BcfGpuError):
        _lib.op_rng_probe(1, 0, 0, 0, 0, 0)


def test_bad_arguments_are_rejected_before_touching_the_device():
    L = _lib.load()
    h = C.c_void_p()
    assert L.bcfgpu_create(None, C.byref(h)) == 1
    c = _lib.default_config()
    c.struct_size = 8
    assert L.bcfgpu_create(C.byref(c), C.byref(h)) == 1
    assert b"struct_size" in L.bcfgpu_last_error()
    for field, bad in (("ntree_con", 0), ("alpha_con", 1.5), ("lambda_", -1.0), ("max_leaves", 1), ("min_leaf", 0)):
        c = _lib.default_config()
        setattr(c, field, bad)
        assert L.bcfgpu_create(C.byref(c), C.byref(h)) == 1, field
    assert L.bcfgpu_run(None, 1, 1, 1) == 1
    assert L.bcfgpu_destroy(None) == 0
