"""The device-side monty::random API (include/monty_b200/random.cuh) and the
host prng-style access to the state array, exercised through the example
consumers (examples/*.cu -> monty_b200/libmonty_b200_examples.so).

  * every sampler called inline from a kernel produces exactly the draws (and
    leaves exactly the states) of the bulk C-ABI path, which the oracle tests
    pin to the reference;
  * the fused SIR particle filter (the reference test-suite's own model,
    tests/testthat/helper-sir-filter.R) matches the same model run over the
    unmodified reference's C++ samplers: compartments and generator states
    bit for bit, the log-likelihood to 1e-12.
"""
import numpy as np
import pytest

from cases import sampler_cases
from gpu_util import states_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import monty_b200
    return monty_b200


@pytest.fixture(scope="module")
def ex():
    import monty_b200.examples
    return monty_b200.examples


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("name", sorted(sampler_cases(4)))
def test_inline_draw_equals_bulk_path(mb, ex, name, rt):
    n = 3000
    pars = sampler_cases(n)[name]
    a = mb.monty_rng_create(n, 21, real_type=rt)
    b = mb.monty_rng_create(n, 21, real_type=rt)
    for _ in range(3):
        inline = ex.inline_draw(a, name, pars)
        bulk = np.asarray(mb.random._sample(b, name, None, pars))
        assert np.array_equal(inline, bulk, equal_nan=True), name
    assert np.array_equal(states_of(mb, a), states_of(mb, b))


def test_inline_multinomial_equals_bulk_path(mb, ex):
    """monty::random::multinomial (hdr/multinomial.hpp:53-82) called inside a
    kernel: the counts and the states of mb200_rng_multinomial (which
    test_gpu_samplers.py pins to the reference), shared and per-stream
    probabilities, zero probabilities included; rows sum to size."""
    n = 2000
    rs = np.random.RandomState(4)
    shared = np.array([0.1, 0.0, 0.4, 0.5, 0.2, 0.0, 1.3])
    per_stream = np.abs(rs.randn(n, 6)) * (rs.rand(n, 6) > 0.2)
    per_stream[:, 3] += 0.01
    size = np.floor(rs.rand(n) * 400)
    for prob, sz in [(shared, 100.0), (per_stream, size), (shared, size)]:
        a = mb.monty_rng_create(n, 77)
        b = mb.monty_rng_create(n, 77)
        for _ in range(2):
            inline = ex.inline_multinomial(a, sz, prob)
            bulk = np.asarray(mb.monty_random_multinomial(sz, prob.T if prob.ndim == 2 else prob, b))
            assert np.array_equal(inline, bulk.T)
            assert np.array_equal(inline.sum(axis=1), np.broadcast_to(sz, (n,)))
        assert np.array_equal(states_of(mb, a), states_of(mb, b))
    # invalid probabilities: a row of NaN instead of the reference's fatal error
    bad = np.tile(np.array([0.5, 0.25, 0.25]), (n, 1))
    bad[7] = [0.5, -0.1, 0.6]
    bad[9] = 0.0
    c = mb.monty_rng_create(n, 78)
    y = ex.inline_multinomial(c, 10.0, bad)
    assert np.isnan(y[7]).all() and np.isnan(y[9]).all()
    assert not np.isnan(np.delete(y, [7, 9], axis=0)).any()


SIR = dict(N=1000.0, I0=10.0, beta=0.2, gamma=0.1, exp_noise=1e6)
TIMES = np.arange(4.0, 81.0, 4.0)
INCIDENCE = np.array([1, 3, 5, 7, 11, 14, 19, 22, 25, 24, 21, 18, 14, 10, 8, 5, 4, 2, 1, 1], dtype=np.float64)


@pytest.mark.parametrize("resample", [False, True])
@pytest.mark.parametrize("n_particles", [100, 4096])
def test_sir_filter_matches_reference_samplers(mb, ex, reference, n_particles, resample):
    system = mb.monty_rng_create(n_particles, 42)
    filt = mb.monty_rng_create(1, 43)
    sys_state = reference.seed_states(42, n_particles)
    fil_state = reference.seed_states(43, 1)
    ll, final = ex.sir_filter(system, filt, SIR, TIMES, INCIDENCE, resample=resample, return_state=True)
    ll_ref, final_ref = reference.sir_filter(sys_state, fil_state, SIR, TIMES, INCIDENCE, resample=resample)
    assert np.array_equal(states_of(mb, system), sys_state)
    assert np.array_equal(states_of(mb, filt), fil_state)
    assert np.array_equal(final, final_ref)
    assert abs(ll - ll_ref) <= 1e-12 * abs(ll_ref)
    # a sanity check on the model itself: the population is conserved
    assert np.all(final[0] + final[1] + final[2] == SIR["N"])


def test_sir_filter_estimates_a_likelihood(mb, ex):
    """Independent replicates of the filter agree to Monte-Carlo accuracy."""
    lls = []
    for seed in range(8):
        system = mb.monty_rng_create(8192, 100 + seed)
        filt = mb.monty_rng_create(1, 200 + seed)
        lls.append(ex.sir_filter(system, filt, SIR, TIMES, INCIDENCE))
    lls = np.array(lls)
    assert np.all(np.isfinite(lls)) and lls.std() < 1.0


def test_zero_inflated_densities_of_the_device_api(mb, oracle):
    """monty::density::zi_poisson / zi_negative_binomial_prob / _mu
    (ref: inst/include/monty/random/density.hpp:355-409) through random.cuh,
    against the bulk density kernels and the reference."""
    from monty_b200 import examples
    from cases import density_cases
    n = 5000
    for which, name in (("poisson", "zi_poisson"), ("negative_binomial_prob", "zi_negative_binomial_prob"),
                        ("negative_binomial_mu", "zi_negative_binomial_mu")):
        x, pars = density_cases(n)[name]
        pars = [np.asarray(p, dtype=np.float64) for p in pars]
        pars[0][:3] = [0.0, 1.0, 1.0]            # the pi0 == 0 and pi0 == 1 branches
        for log in (True, False):
            got = examples.zi_density(which, x.astype(np.float64), pars[0], pars[1], pars[2] if len(pars) > 2 else None, log)
            bulk = mb.density._density(name, x, pars, log)
            assert np.array_equal(got, bulk, equal_nan=True), name
            exp = oracle.density(name, x, pars, log)
            assert np.allclose(got, exp, rtol=1e-12, atol=1e-300, equal_nan=True), name


def test_examples_library_is_built_without_fma_contraction():
    """The device-side API's arithmetic contract (-fmad=false): checked at run
    time by monty::random::contraction_is_off()."""
    from monty_b200 import examples
    assert examples.math_probe("contraction_is_off", np.zeros(4))[0] == 1.0
