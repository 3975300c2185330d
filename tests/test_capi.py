"""The C-ABI shared library: it loads, exports every symbol include/
monty_b200.h declares, and refuses to run without a GPU (no CPU fallback).
No compute calls here: this file runs on the CPU-only build box."""
import ctypes
import os
import re

import pytest

from oracle_lib import ROOT

HEADER = os.path.join(ROOT, "include", "monty_b200.h")
LIB = os.path.join(ROOT, "monty_b200", "libmonty_b200.so")


def declared_symbols():
    txt = open(HEADER).read()
    names = set(re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", txt))
    for macro, args in re.findall(r"^MB200_DECL_SAMPLER\d\((\w+)(.*?)\)", txt, flags=re.M):
        names.add("mb200_random_" + macro)
    names.discard("mb200_random_")
    return sorted(n for n in names if not n.endswith("##name"))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 45
    for must in ("mb200_rng_create", "mb200_rng_state", "mb200_rng_set_state", "mb200_rng_jump",
                 "mb200_rng_long_jump", "mb200_rng_sample", "mb200_rng_sample_dev",
                 "mb200_rng_multinomial", "mb200_density", "mb200_density_dev",
                 "mb200_random_binomial", "mb200_random_truncated_normal", "mb200_random_real"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build it: make -C monty_b200/csrc"
    lib = ctypes.CDLL(LIB)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_no_cpu_fallback():
    import monty_b200 as mb
    assert b"monty_b200" in mb._lib.lib.mb200_version()
    if mb._lib.lib.mb200_device_count() > 0:
        pytest.skip("a GPU is present; the refusal path is for CPU-only boxes")
    with pytest.raises(mb.MontyError, match="no usable CUDA device"):
        mb.monty_rng_create(4, 42)
    with pytest.raises(mb.MontyError, match="no usable CUDA device"):
        mb.density.density_poisson([1, 2], [1.0, 2.0], True)


def test_boundary_errors_need_no_gpu():
    """Argument checks that happen before any device work."""
    import monty_b200 as mb
    with pytest.raises(mb.MontyError, match="Invalid type for 'seed'"):
        mb.monty_rng_create(1, seed="nope")
    with pytest.raises(mb.MontyError, match="Expected raw vector of length as multiple of 32 for 'seed'"):
        mb.monty_rng_create(1, seed=bytes(31))
    with pytest.raises(mb.MontyError, match="Unknown normal algorithm"):
        mb.monty_random_normal(0, 1, None, algorithm="wrong")


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under monty_b200/ may import,
    link or execute it."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "monty_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle_lib|libmonty_oracle|libmonty_ref|oracle/", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_examples_library_and_host_headers():
    """The consumers of the device-side API build into their own library; the
    host prng class compiles as plain C++ against the C-ABI header."""
    import subprocess
    ex_header = os.path.join(ROOT, "examples", "examples.h")
    ex_lib = os.path.join(ROOT, "monty_b200", "libmonty_b200_examples.so")
    assert os.path.exists(ex_lib), "build it: make -C monty_b200/csrc"
    ctypes.CDLL(LIB, mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(ex_lib)
    for name in re.findall(r"\b(mb200_example_[a-z0-9_]+)\s*\(", open(ex_header).read()):
        assert hasattr(lib, name), name
    src = '#include "include/monty_b200/prng.hpp"\nint main() { return sizeof(monty_b200::prng) ? 0 : 1; }\n'
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", ROOT, "-"],
                       input=src, text=True, capture_output=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr


def test_device_facade_declares_the_reference_names():
    """include/monty_b200/random.cuh mirrors the C++ names odin2 generates
    (reference R/dsl-distributions.R:57-340)."""
    txt = open(os.path.join(ROOT, "include", "monty_b200", "random.cuh")).read()
    for name in ("uniform", "exponential_rate", "exponential_mean", "binomial", "poisson", "gamma_scale",
                 "gamma_rate", "beta", "negative_binomial_prob", "negative_binomial_mu", "hypergeometric",
                 "truncated_normal", "cauchy", "weibull", "log_normal", "beta_binomial_prob",
                 "beta_binomial_ab", "zi_poisson", "zi_negative_binomial_prob", "zi_negative_binomial_mu"):
        assert re.search(r"MONTY_B200_SAMPLER\(%s," % name, txt), name
    for name in ("random_real", "random_normal", "normal", "seed", "jump", "long_jump", "generator"):
        assert re.search(r"\b%s\b" % name, txt), name
