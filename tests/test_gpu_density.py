"""GPU parity of the (log-)densities (ref: src/density.cpp, density.hpp;
tests/testthat/test-densities.R) and of the fused device-side sum."""
import numpy as np
import pytest
from scipy import stats

from cases import density_cases
from gpu_util import ulp_distance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import monty_b200
    return monty_b200


def gpu_density(mb, name, x, pars, log, rt="f64", **kw):
    return mb.density._density(name, x, pars, log, rt, **kw)


@pytest.mark.parametrize("name", sorted(density_cases(4)))
@pytest.mark.parametrize("log", [True, False])
def test_double_parity(mb, oracle, name, log):
    """The north-star's bar is log-densities within 4 ulp of the reference.  All 21
    are BIT-IDENTICAL, in log and in non-log form: each is assembled from log, exp,
    pow, lgamma, log1p, erfc (glibc's bits on the device,
    include/monty_b200/detail/glibc_math.cuh) and IEEE arithmetic only."""
    x, pars = density_cases(20000)[name]
    got = gpu_density(mb, name, x, pars, log)
    exp = oracle.density(name, x, pars, log)
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    assert np.array_equal(np.isinf(got), np.isinf(exp))
    fin = np.isfinite(exp)
    d = ulp_distance(got[fin], exp[fin])
    assert d.max() == 0, (name, log, d.max())


def test_fixture_parity(mb, golden):
    for name, (x, pars) in density_cases(129).items():
        got = gpu_density(mb, name, x, pars, True)
        exp = golden["density/f64/%s/1" % name]
        assert np.array_equal(got, exp, equal_nan=True), name


def test_float_variant(mb, oracle):
    """density_negative_binomial_mu(..., is_float = TRUE), src/density.cpp:28-45;
    the reference tests it at tolerance sqrt(sqrt(eps)) (test-densities.R:62-67)."""
    x, pars = density_cases(5000)["negative_binomial_mu"]
    got = mb.density.density_negative_binomial_mu(x, pars[0], pars[1], True, is_float=True)
    exp = oracle.density("negative_binomial_mu", x, pars, True, "f32")
    assert np.allclose(got, exp, rtol=2e-4, atol=2e-4)
    same = (got == exp) | (np.isnan(got) & np.isnan(exp))          # logf / lgammaf / log1pf are glibc's bits
    assert same.all(), float(1 - same.mean())


@pytest.mark.parametrize("name", sorted(density_cases(4)))
def test_float_parity(mb, oracle, name):
    """real_type = float (the reference's `<float>` instantiations of density.hpp): logf, expf,
    powf, lgammaf, log1pf return glibc's bits on the device, so the float densities are
    bit-identical as well, in log and in non-log form."""
    x, pars = density_cases(20000)[name]
    for log in (True, False):
        got = gpu_density(mb, name, x, pars, log, "f32")
        exp = oracle.density(name, x, pars, log, "f32")
        same = (got == exp) | (np.isnan(got) & np.isnan(exp))
        assert same.all(), (name, log, float(1 - same.mean()), got[~same][:3], exp[~same][:3])


def test_known_answers_against_scipy(mb):
    """ref: tests/testthat/test-densities.R compares with base R d*() at 1.5e-8."""
    rs = np.random.RandomState(1)
    x = rs.poisson(20, 50).astype(np.int32)
    lam = rs.uniform(1, 40, 50)
    assert np.allclose(mb.density.density_poisson(x, lam, True), stats.poisson.logpmf(x, lam), rtol=1.5e-8)
    size = rs.uniform(1, 30, 50)
    mu = rs.uniform(1, 30, 50)
    assert np.allclose(mb.density.density_negative_binomial_mu(x, size, mu, True),
                       stats.nbinom.logpmf(x, size, size / (size + mu)), rtol=1.5e-8)
    n = (x + rs.poisson(10, 50)).astype(np.int32)
    p = rs.uniform(0.1, 0.9, 50)
    assert np.allclose(mb.density.density_binomial(x, n, p, True), stats.binom.logpmf(x, n, p), rtol=1.5e-8)
    a, b = rs.uniform(0.5, 5, 50), rs.uniform(0.5, 5, 50)
    assert np.allclose(mb.density.density_beta_binomial_ab(x, n, a, b, True),
                       stats.betabinom.logpmf(x, n, a, b), rtol=1.5e-8)
    xd = rs.normal(0, 2, 50)
    assert np.allclose(mb.density.density_normal(xd, np.full(50, 0.5), np.full(50, 2.0), False),
                       stats.norm.pdf(xd, 0.5, 2.0), rtol=1.5e-8)
    xg = rs.gamma(2, 2, 50)
    assert np.allclose(mb.density.density_gamma_rate(xg, np.full(50, 2.0), np.full(50, 0.5), True),
                       stats.gamma.logpdf(xg, 2.0, scale=2.0), rtol=1.5e-8)


def test_corner_cases(mb):
    """ref: tests/testthat/test-densities.R:10-17, 47-56, 92-147, 215-237."""
    d = mb.density
    assert d.density_binomial([0], [0], [0.5], True)[0] == 0
    assert d.density_poisson([0], [0.0], True)[0] == 0
    assert d.density_negative_binomial_mu([0], [0.0], [1.0], True)[0] == 0
    assert d.density_negative_binomial_mu([1], [0.0], [1.0], True)[0] == -np.inf
    assert d.density_negative_binomial_mu([0], [1.0], [0.0], True)[0] == 0
    assert d.density_negative_binomial_mu([3], [1.0], [0.0], False)[0] == 0
    assert d.density_normal([1.0], [1.0], [0.0], True)[0] == np.inf
    assert d.density_normal([1.5], [1.0], [0.0], True)[0] == -np.inf
    assert d.density_exponential_rate([-1.0], [2.0], True)[0] == -np.inf
    assert d.density_uniform([3.0], [0.0], [2.0], False)[0] == 0
    assert d.density_beta_binomial_ab([5], [3], [1.0], [1.0], True)[0] == -np.inf
    # mu << size: the Poisson-limit branch (density.hpp:113-117)
    got = d.density_negative_binomial_mu([3], [1e12], [1e-8], True)[0]
    assert abs(got - stats.poisson.logpmf(3, 1e-8)) < 1e-6


def test_device_sum_matches_vector_sum(mb, oracle):
    """The fused block/grid reduction (config C5): same value as summing the
    per-observation vector, to 1e-12 relative (summation order differs)."""
    n = 1 << 20
    x, pars = density_cases(n)["poisson"]
    vec, total = mb.density._density("poisson", x, pars, True, want_out=True, want_sum=True)
    only = mb.density.log_likelihood("poisson", x, pars)
    assert total == only
    assert abs(total - np.sum(vec)) <= 1e-12 * abs(np.sum(vec))
    exp = oracle.density("poisson", x[:50000], [pars[0][:50000]], True)
    assert np.allclose(vec[:50000], exp, rtol=1e-13)
    # reproducible run to run
    assert mb.density.log_likelihood("poisson", x, pars) == only


def test_device_resident_sum_and_bench_inputs(mb, oracle):
    """monty_b200.distributed.density_sum_dev on torch tensors, with the C5
    inputs bench.py generates on the device (same counter hash as the CPU)."""
    import os, sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from monty_b200.synthetic import density_inputs_torch
    from monty_b200.distributed import density_sum_dev
    dev = torch.device("cuda", 0)
    n = 200000
    for name in ("poisson", "negative_binomial_mu", "beta_binomial_prob"):
        x, pars = density_inputs_torch(name, n, 1000, dev, chunk=65536)
        from oracle_lib import counter_hash
        idx = np.arange(1000, 1000 + n)
        assert np.array_equal(x.cpu().numpy(), np.floor(200 * counter_hash(idx, 0) ** 2).astype(np.int32))
        total = density_sum_dev(name, x, pars).item()
        ref = oracle.density(name, x.cpu().numpy(), [p.cpu().numpy() for p in pars], True)
        assert np.isfinite(ref).all()
        assert abs(total - ref.sum()) <= 1e-11 * abs(ref.sum())
