"""Pins the C oracle (oracle/monty_oracle.c) -- the checker of the CUDA path.

(1) against the reference's own golden files (tests/golden/xoshiro-ref, copied
    from tests/testthat/xoshiro-ref; protocol src/test_rng.cpp:21-38),
(2) against vectors generated from the unmodified reference headers
    (tests/golden/reference_vectors.npz, tools/make_golden.py),
(3) bit-for-bit against the compiled reference itself when oracle/_ref exists,
(4) through the reference test-suite's metamorphic identities
    (tests/testthat/test-rng.R, cited per test).
"""
import os

import numpy as np
import pytest

from cases import density_cases, sampler_cases
from oracle_lib import GEN_NAMES, OracleError

HERE = os.path.dirname(os.path.abspath(__file__))
N_STREAMS, N_SAMPLES, SEED = 64, 5, 20240607


def golden_file(name):
    with open(os.path.join(HERE, "golden", "xoshiro-ref", name)) as fh:
        return fh.read().split()


def as_golden_lines(name, vals, state):
    width = 8 if name.startswith("xoshiro128") else 16
    return [str(int(v)) for v in vals] + [format(int(w), "0%dx" % width) for w in state]


@pytest.mark.parametrize("gen", range(12))
def test_golden_files_port(port, gen):
    vals, st = port.xoshiro_run(gen)
    assert as_golden_lines(GEN_NAMES[gen], vals, st) == golden_file(GEN_NAMES[gen])


@pytest.mark.parametrize("gen", range(12))
def test_golden_files_reference(reference, gen):
    vals, st = reference.xoshiro_run(gen)
    assert as_golden_lines(GEN_NAMES[gen], vals, st) == golden_file(GEN_NAMES[gen])


def test_seeding_and_jumps_vs_fixtures(port, golden):
    assert np.array_equal(port.seed_states(42, 100), golden["seed_states_42_100"])
    st = port.seed_states(7, 33)
    port.jump(st, 3)
    assert np.array_equal(st, golden["jump3_seed7_33"])
    port.jump(st, 2, long=True)
    assert np.array_equal(st, golden["longjump2_after"])
    # seed vector: full states are consumed, then jump from the LAST supplied one
    # (ref: prng.hpp:45-51)
    assert np.array_equal(port.seed_states(0, 10, golden["raw_seed_words"]), golden["raw_seed_states_10"])
    for gen in range(3):
        s = port.seed_states(SEED, 5)
        assert np.array_equal(port.raw(gen, s, 11), golden["raw_gen%d" % gen])
        assert np.array_equal(s, golden["raw_gen%d_state" % gen])


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("name", sorted(sampler_cases(4)))
def test_samplers_vs_fixtures(port, golden, rt, name):
    pars = sampler_cases(N_STREAMS)[name]
    for det in (False, True):
        if det and name == "cauchy":
            continue
        s = port.seed_states(SEED, N_STREAMS)
        y = port.sample(name, s, N_SAMPLES, pars, rt, det)
        key = "sample/%s/%s/%s" % (rt, name, "det" if det else "sto")
        assert np.array_equal(y, golden[key], equal_nan=True), key
        assert np.array_equal(s, golden[key + "/state"]), key


@pytest.mark.parametrize("rt", ["f64", "f32"])
def test_densities_vs_fixtures(port, golden, rt):
    for name, (x, pars) in density_cases(129).items():
        for lg in (True, False):
            got = port.density(name, x, pars, lg, rt)
            assert np.array_equal(got, golden["density/%s/%s/%d" % (rt, name, lg)], equal_nan=True), (name, lg)


def test_multinomial_vs_fixture(port, golden):
    s = port.seed_states(SEED, 9)
    y = port.multinomial(s, 4, np.array([100.0]), np.array([0.1, 0.0, 0.4, 0.5, 0.2]))
    assert np.array_equal(y, golden["multinomial/shared"])
    assert np.array_equal(s, golden["multinomial/shared/state"])
    assert (y.sum(axis=-1) == 100).all()


# ---- live comparison with the compiled reference (larger, when available) ----
@pytest.mark.parametrize("rt", ["f64", "f32"])
def test_port_equals_reference_live(port, reference, rt):
    n = 2000
    for name, pars in sampler_cases(n).items():
        for det in (False, True):
            if det and name == "cauchy":
                continue
            a = port.seed_states(11, n)
            b = a.copy()
            x = port.sample(name, a, 6, pars, rt, det)
            y = reference.sample(name, b, 6, pars, rt, det)
            assert np.array_equal(x, y, equal_nan=True), (name, det)
            assert np.array_equal(a, b), (name, det)
    for name, (x, pars) in density_cases(3000).items():
        for lg in (True, False):
            assert np.array_equal(port.density(name, x, pars, lg, rt),
                                  reference.density(name, x, pars, lg, rt), equal_nan=True), name


def test_port_threads_equal_serial(port):
    n = 512
    pars = sampler_cases(n)["gamma_scale"]
    a = port.seed_states(3, n)
    b = a.copy()
    x = port.sample("gamma_scale", a, 9, pars, n_threads=1)
    y = port.sample("gamma_scale", b, 9, pars, n_threads=4)
    assert np.array_equal(x, y) and np.array_equal(a, b)


ERROR_CASES = [
    # ref: tests/testthat/test-rng.R:137-168, 261-268, 970-986, 1072-1083
    ("binomial", [np.array([10.0, 10, -1, 10]), np.array([0.5])],
     "Invalid call to binomial with n = -1, p = 0.5, q = 0.5", 2),
    ("binomial", [np.array([10.0]), np.array([0.5, 1.5, 0.1, 0.1])],
     "Invalid call to binomial with n = 10, p = 1.5, q = -0.5", 1),
    ("poisson", [np.array([1.0, np.inf, 1, 1])], "Invalid call to Poisson with lambda = inf", 1),
    ("poisson", [np.array([-1.0])], "Invalid call to Poisson with lambda = -1", 0),
    ("gamma_scale", [np.array([1.0, 2, -3, 1]), np.array([1.0])],
     "Invalid call to gamma with shape = -3, scale = 1", 2),
    ("negative_binomial_prob", [np.array([1.0, 2, 3, 1]), np.array([0.5, 0.5, 0, 0.5])],
     "Invalid call to negative_binomial with size = 3, prob = 0", 2),
    ("hypergeometric", [np.array([5.0]), np.array([5.0]), np.array([3.0, 11, 3, 3])],
     "Invalid call to hypergeometric with n1 = 5, n2 = 5, k = 11", 1),
    ("weibull", [np.array([-1.0]), np.array([1.0])], "Invalid call to Weibull with shape = -1, scale = 1", 0),
    ("log_normal", [np.array([1.0]), np.array([0.0])],
     "Invalid call to log_normal with meanlog = 1, sdlog = 0", 0),
    ("beta_binomial_ab", [np.array([5.0]), np.array([0.0]), np.array([1.0])],
     "Invalid call to beta_binomial with size = 5, a = 0, b = 1", 0),
    ("zi_poisson", [np.array([2.0]), np.array([1.0])], "Invalid call to zi_poisson with pi0 = 2, lambda = 1", 0),
    ("zi_negative_binomial_mu", [np.array([0.5]), np.array([-1.0]), np.array([1.0])],
     "Invalid call to zi_negative_binomial with pi0 = 0.5, size = -1, prob = -inf", 0),
]


@pytest.mark.parametrize("name,pars,message,offender", ERROR_CASES)
def test_error_text_and_partial_advance(port, name, pars, message, offender):
    """The loop throws at the first bad stream: earlier streams stay advanced,
    later ones are untouched (ref: src/random.cpp:211-228 + BEGIN_CPP11)."""
    st = port.seed_states(5, 4)
    before = st.copy()
    with pytest.raises(OracleError) as e:
        port.sample(name, st, 3, pars)
    assert str(e.value) == message
    changed = (st != before).any(axis=1)
    assert changed[:offender].all() and not changed[offender:].any()


def test_binomial_nan_edge_cases_allowed(port):
    # ref: tests/testthat/test-rng.R:158-168 -- binomial(0, NaN) == 0
    st = port.seed_states(1, 1)
    assert port.sample("binomial", st, 1, [np.array([0.0]), np.array([np.nan])])[0, 0] == 0


# ---- metamorphic identities of the reference's test-suite, on the oracle ------
def _draw(port, name, pars, n=8, ns=50, seed=1, rt="f64"):
    st = port.seed_states(seed, n)
    return port.sample(name, st, ns, pars, rt), st


def test_uniform01_is_real(port):                       # test-rng.R:41-52
    a, sa = _draw(port, "real", [])
    b, sb = _draw(port, "uniform", [np.array([0.0]), np.array([1.0])])
    assert np.array_equal(a, b) and np.array_equal(sa, sb)


@pytest.mark.parametrize("algo", ["normal_box_muller", "normal_polar", "normal_ziggurat"])
def test_normal_scaling(port, algo):                    # test-rng.R:321-333
    z, _ = _draw(port, algo, [np.array([0.0]), np.array([1.0])])
    y, _ = _draw(port, algo, [np.array([3.0]), np.array([2.0])])
    assert np.array_equal(y, z * 2.0 + 3.0)


@pytest.mark.parametrize("size,prob", [(10.0, 0.3), (1000.0, 0.3)])
def test_binomial_complement(port, size, prob):         # test-rng.R:115-134
    a, _ = _draw(port, "binomial", [np.array([size]), np.array([prob])])
    b, _ = _draw(port, "binomial", [np.array([size]), np.array([1 - prob])])
    assert np.array_equal(a, size - b)


def test_exponential_mean_is_rate(port):                # test-rng.R:357-364
    a, _ = _draw(port, "exponential_mean", [np.array([4.0])])
    b, _ = _draw(port, "exponential_rate", [np.array([0.25])])
    assert np.array_equal(a, b)


def test_gamma_identities(port):                        # test-rng.R:1008-1029
    a, _ = _draw(port, "gamma_scale", [np.array([1.0]), np.array([0.5])])
    b, _ = _draw(port, "exponential_mean", [np.array([0.5])])
    assert np.array_equal(a, b)
    c, _ = _draw(port, "gamma_rate", [np.array([2.5]), np.array([4.0])])
    d, _ = _draw(port, "gamma_scale", [np.array([2.5]), np.array([0.25])])
    assert np.array_equal(c, d)


def test_beta_is_ratio_of_gammas(port):                 # test-rng.R:1271-1281
    st = port.seed_states(1, 1)
    b = port.sample("beta", st, 20, [np.array([2.5]), np.array([1.5])])[0]
    st = port.seed_states(1, 1)
    out = []
    for _ in range(20):
        x = port.sample("gamma_scale", st, 1, [np.array([2.5]), np.array([1.0])])[0, 0]
        y = port.sample("gamma_scale", st, 1, [np.array([1.5]), np.array([1.0])])[0, 0]
        out.append(x / (x + y))
    assert np.array_equal(b, np.array(out))


def test_nbinom_mu_is_prob(port):                       # test-rng.R:1098-1110
    size, prob = 20.0, 0.3
    mu = size * (1 - prob) / prob
    a, _ = _draw(port, "negative_binomial_prob", [np.array([size]), np.array([prob])])
    b, _ = _draw(port, "negative_binomial_mu", [np.array([size]), np.array([mu])])
    assert np.array_equal(a, b)


def test_hypergeometric_symmetries(port):               # test-rng.R:932-949
    n1, n2, k = 70.0, 100.0, 80.0
    a, _ = _draw(port, "hypergeometric", [np.array([n1]), np.array([n2]), np.array([k])])
    b, _ = _draw(port, "hypergeometric", [np.array([n2]), np.array([n1]), np.array([k])])
    assert np.array_equal(a, k - b)
    c, _ = _draw(port, "hypergeometric", [np.array([n1]), np.array([n2]), np.array([n1 + n2 - k])])
    assert np.array_equal(a, n1 - c)


def test_truncated_normal_identities(port):             # test-rng.R:1334-1354
    a, _ = _draw(port, "truncated_normal", [np.array([1.0]), np.array([2.0]), np.array([-np.inf]), np.array([np.inf])])
    b, _ = _draw(port, "normal_box_muller", [np.array([1.0]), np.array([2.0])])
    assert np.array_equal(a, b)
    lo, _ = _draw(port, "truncated_normal", [np.array([0.0]), np.array([1.0]), np.array([-np.inf]), np.array([-1.0])])
    hi, _ = _draw(port, "truncated_normal", [np.array([0.0]), np.array([1.0]), np.array([1.0]), np.array([np.inf])])
    assert np.array_equal(lo, -hi)


def test_zi_poisson_pi0_zero_is_poisson(port):          # test-rng.R:1504-1517
    a, sa = _draw(port, "zi_poisson", [np.array([0.0]), np.array([7.0])])
    b, sb = _draw(port, "poisson", [np.array([7.0])])
    assert np.array_equal(a, b) and np.array_equal(sa, sb)


def test_no_draw_exits_leave_state_untouched(port):     # test-rng.R:271-277, 989-1005, 1520-1534
    for name, pars in [("poisson", [np.array([0.0])]),
                       ("hypergeometric", [np.array([0.0]), np.array([5.0]), np.array([3.0])]),
                       ("hypergeometric", [np.array([5.0]), np.array([5.0]), np.array([0.0])]),
                       ("zi_poisson", [np.array([1.0]), np.array([5.0])]),
                       ("binomial", [np.array([0.0]), np.array([0.5])]),
                       ("binomial", [np.array([10.0]), np.array([0.0])]),
                       ("binomial", [np.array([10.0]), np.array([1.0])]),
                       ("gamma_scale", [np.array([0.0]), np.array([1.0])])]:
        st = port.seed_states(2, 3)
        before = st.copy()
        port.sample(name, st, 4, pars)
        assert np.array_equal(st, before), name


def test_stream_k_is_k_jumps(port):                     # test-rng.R:11-38, 380-391
    many = port.seed_states(42, 6)
    one = port.seed_states(42, 1)
    for k in range(6):
        assert np.array_equal(many[k], one[0])
        port.jump(one, 1)


def test_moments_quick(port):                           # test-rng.R:64-112 (reduced n)
    st = port.seed_states(1, 4)
    z = port.sample("normal_ziggurat", st, 50000, [np.array([0.0]), np.array([1.0])]).ravel()
    assert abs(z.mean()) < 1e-2 and abs(z.var() - 1) < 2e-2
    b = port.sample("binomial", st, 50000, [np.array([100.0]), np.array([0.3])]).ravel()
    assert abs(b.mean() - 30) < 0.1 and abs(b.var() - 21) < 0.5


def test_generator_vectors_are_the_reference(reference):
    """tests/golden/generator_vectors.npz (the reference's samplers on its other generators,
    what tests/test_gpu_generators.py holds the device-side API to) is what the live reference
    build produces (tools/make_golden_generators.py)."""
    import os
    vec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generator_vectors.npz"))
    from cases import sampler_cases
    n, ns, seed = 48, 6, 20240607
    for gen, rt in (("xoshiro128plus", "f32"), ("xoshiro128plus", "f64"), ("xoshiro256plus", "f64"),
                    ("xoroshiro128plusplus", "f64"), ("xoshiro512starstar", "f64")):
        for name, pars in sampler_cases(n).items():
            y, st = reference.generator_draw(gen, name, seed, n, ns, pars, rt, jumps=1, long_jumps=1)
            assert np.array_equal(st, vec["%s/%s/%s/state" % (gen, rt, name)]), (gen, rt, name)
            assert np.array_equal(y, vec["%s/%s/%s" % (gen, rt, name)], equal_nan=True), (gen, rt, name)
    # generator<double> on xoshiro256+ through this path equals the bulk path's oracle
    st0 = reference.seed_states(seed, n)
    reference.jump(st0, 1)
    reference.jump(st0, 1, long=True)
    exp = reference.sample("normal_ziggurat", st0, ns, sampler_cases(n)["normal_ziggurat"])
    assert np.array_equal(exp, vec["xoshiro256plus/f64/normal_ziggurat"])
    assert np.array_equal(st0, vec["xoshiro256plus/f64/normal_ziggurat/state"])
