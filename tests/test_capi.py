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
