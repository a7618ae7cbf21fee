"""Builds libbcfgpu.so (the C-ABI library, include/bcfgpu.h) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbcfgpu.so")
SOURCES = ["bcfgpu.cu", "bcf_host.cpp"]
HEADERS = ["bcf_device.cuh", "bcf_step.cuh", "bcf_kernels.cuh", "bcf_persistent.cuh", "bcf_batched.cuh", "bcf_pipe.cuh", "bcf_hist_mma.cuh", "bcf_host.h", "bcf_glm.cuh", "bcfgpu_glm.inl", os.path.join("..", "..", "include", "bcfgpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # no FMA contraction: every kernel (and both engines, and any sharding) rounds identically
              "-fmad=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libbcfgpu.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(p) > t for p in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("BCFGPU_NVCC_EXTRA", "").split()      # e.g. -DBCF_PROF: compile the phase timers / batch timeline of k_pipe in (BCFGPU_PROFILE, BCFGPU_TRACE); they cost ~6 % of the sweep
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
