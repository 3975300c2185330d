"""The cpp11 binding of the reference's R package over the C-ABI (r_glue/).

R and cpp11 are not installed here, so the glue cannot be linked or run; what
can be checked on this box is (1) that every file compiles -- `g++
-fsyntax-only` against a declarations-only stand-in for the cpp11 / R API
(tests/cpp11_stub/) and the real include/monty_b200.h, so every C-ABI call is
type-checked -- and (2) that the 77 routines the reference registers
(src/cpp11.cpp:552-638: 49 samplers + 2 multinomial, 4 state routines, the
allocator, 21 densities... listed below) are all defined, with the argument
counts R/cpp11.R passes."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "r_glue")

# name -> number of arguments, as registered by the reference (src/cpp11.cpp:552-630)
SAMPLERS = {"real": 0, "exponential_rate": 1, "exponential_mean": 1, "poisson": 1, "beta": 2, "binomial": 2,
            "cauchy": 2, "gamma_scale": 2, "gamma_rate": 2, "negative_binomial_prob": 2,
            "negative_binomial_mu": 2, "normal_box_muller": 2, "normal_polar": 2, "normal_ziggurat": 2,
            "uniform": 2, "weibull": 2, "log_normal": 2, "zi_poisson": 2, "beta_binomial_prob": 3,
            "beta_binomial_ab": 3, "hypergeometric": 3, "zi_negative_binomial_prob": 3,
            "zi_negative_binomial_mu": 3, "truncated_normal": 4, "multinomial": 2}
DENSITIES = {"beta": 4, "beta_binomial_ab": 5, "beta_binomial_prob": 5, "binomial": 4, "cauchy": 4,
             "exponential_mean": 3, "exponential_rate": 3, "gamma_rate": 4, "gamma_scale": 4,
             "hypergeometric": 5, "log_normal": 4, "negative_binomial_mu": 5, "negative_binomial_prob": 4,
             "normal": 4, "poisson": 3, "truncated_normal": 6, "uniform": 4, "weibull": 4,
             "zi_negative_binomial_mu": 5, "zi_negative_binomial_prob": 5, "zi_poisson": 4}


def expected_routines():
    r = {"monty_rng_alloc": 3, "cpp_monty_rng_state": 1, "cpp_monty_rng_set_state": 2, "cpp_monty_rng_jump": 2,
         "cpp_monty_rng_long_jump": 2, "test_xoshiro_run": 1}
    for name, k in SAMPLERS.items():
        r["cpp_monty_random_" + name] = k + 1            # parameters + ptr
        r["cpp_monty_random_n_" + name] = k + 2          # n_samples + parameters + ptr
    for name, k in DENSITIES.items():
        r["density_" + name] = k
    return r


def preprocessed(path):
    out = subprocess.run(["g++", "-std=c++17", "-E", "-P", "-I", os.path.join(ROOT, "tests", "cpp11_stub"),
                          "-I", os.path.join(ROOT, "include"), path], capture_output=True, text=True, check=True)
    return out.stdout


@pytest.mark.parametrize("name", ["random.cpp", "density.cpp", "test_rng.cpp"])
def test_glue_compiles_against_the_c_abi(name):
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-attributes",
                          "-I", os.path.join(ROOT, "tests", "cpp11_stub"), "-I", os.path.join(ROOT, "include"),
                          os.path.join(GLUE, name)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_every_registered_routine_is_defined():
    want = expected_routines()
    assert len(want) == 77
    text = "".join(preprocessed(os.path.join(GLUE, f)) for f in ("random.cpp", "density.cpp", "test_rng.cpp"))
    found = {}
    for m in re.finditer(r"\[\[cpp11::register\]\]\s*[\w:<>\s]+?\s(\w+)\s*\(([^)]*)\)\s*\{", text):
        args = [a for a in m.group(2).split(",") if a.strip()]
        found[m.group(1)] = len(args)
    missing = sorted(set(want) - set(found))
    assert not missing, missing
    assert not sorted(set(found) - set(want))
    wrong = {k: (found[k], want[k]) for k in want if found[k] != want[k]}
    assert not wrong, wrong
