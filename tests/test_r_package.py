"""The R side of the boundary cannot be executed here (no R in the image): check what can be checked -- the .Call
shim compiles against a stub of R's headers and binds only symbols include/bcfgpu.h declares; the R sources export
the reference's three functions (NAMESPACE:3-5 of fbargaglistoffi/BCF-IV) with the reference's formal arguments."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bcf_iv_b200", "R", "BayesIVgpu")


def test_shim_compiles_against_header():
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "tests", "r_stub"),
                           "-I" + os.path.join(ROOT, "include"), "-Wno-cast-function-type",
                           os.path.join(PKG, "src", "r_shim.c")])


def test_shim_uses_only_declared_entry_points():
    hdr = open(os.path.join(ROOT, "include", "bcfgpu.h")).read()
    declared = set(re.findall(r"\b(bcfgpu_\w+)\s*\(", hdr))
    used = set(re.findall(r"\b(bcfgpu_(?!R_)\w+)\s*\(", open(os.path.join(PKG, "src", "r_shim.c")).read()))
    assert used and used <= declared, used - declared


def _formals(path, fn):
    src = open(path).read()
    m = re.search(fn + r"\s*<-\s*function\s*\((.*?)\)\s*\{", src, re.S)
    assert m, fn
    return [a.split("=")[0].strip() for a in re.split(r",(?![^()]*\))", m.group(1))]


def test_r_functions_keep_the_reference_signatures():
    ns = open(os.path.join(PKG, "NAMESPACE")).read()
    for f in ("bcf_iv", "bcf_itt", "generate_dataset", "bcf"):
        assert f"export({f})" in ns
    assert _formals(os.path.join(PKG, "R", "bcf_iv.R"), "bcf_iv") == [
        "y", "w", "z", "x", "binary", "n_burn", "n_sim", "inference_ratio", "max_depth", "cp", "minsplit", "adj_method", "seed"]
    assert _formals(os.path.join(PKG, "R", "bcf_itt.R"), "bcf_itt") == [
        "y", "w", "z", "x", "binary", "max_depth", "n_burn", "n_sim", "inference_ratio", "seed"]      # R/bcf_itt.R:33-34
    assert _formals(os.path.join(PKG, "R", "generate_dataset.R"), "generate_dataset")[:6] == [
        "n", "p", "rho", "null", "effect_size", "compliance"]
    b = _formals(os.path.join(PKG, "R", "bcf.R"), "bcf")
    assert b[:7] == ["y", "z", "x_control", "x_moderate", "pihat", "nburn", "nsim"]
    for a in ("ntree_control", "ntree_moderate", "base_control", "power_control", "nu", "lambda", "sigq", "sighat",
              "include_pi", "use_muscale", "use_tauscale"):
        assert a in b
