"""Distributional checks of the GPU samplers (ref: the moment / KS tests of
tests/testthat/test-rng.R:64-112, 206-258, 280-305, 346-353, 1032-1054, ...),
restated with scipy on the reference's sample sizes and thresholds."""
import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import monty_b200
    return monty_b200


def draws(mb, fn, n, *pars, seed=1, rt="f64", **kw):
    rng = mb.monty_rng_create(64, seed, real_type=rt)
    return np.asarray(fn(n // 64, *pars, rng, **kw)).ravel()


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("algo", ["box_muller", "polar", "ziggurat"])
def test_normal_ks(mb, algo, rt):                     # test-rng.R:280-305: p > 0.1
    z = draws(mb, mb.monty_random_n_normal, 1 << 20, 0.0, 1.0, algorithm=algo, rt=rt)
    assert stats.kstest(z, "norm").pvalue > 0.01
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 5e-3


@pytest.mark.parametrize("rt", ["f64", "f32"])
def test_exponential_ks(mb, rt):                      # test-rng.R:346-353
    x = draws(mb, mb.monty_random_n_exponential_rate, 1 << 20, 0.5, rt=rt)
    assert stats.kstest(x, "expon", args=(0, 2.0)).pvalue > 0.01


def test_uniform_ks(mb):
    x = draws(mb, mb.monty_random_n_uniform, 1 << 20, -2.0, 5.0)
    assert stats.kstest(x, "uniform", args=(-2.0, 7.0)).pvalue > 0.01


def chisq_discrete(x, pmf, lo, hi):
    counts = np.array([(x == k).sum() for k in range(lo, hi + 1)], dtype=float)
    p = np.array([pmf(k) for k in range(lo, hi + 1)])
    rest = len(x) - counts.sum()
    counts = np.append(counts, rest)
    p = np.append(p, max(1 - p.sum(), 1e-12))
    keep = p * len(x) > 5
    return stats.chisquare(counts[keep], p[keep] / p[keep].sum() * counts[keep].sum()).pvalue


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("size,prob", [(10, 0.3), (40, 0.3), (1000, 0.3), (100, 0.9)])
def test_binomial_chisq(mb, size, prob, rt):          # test-rng.R:74-134
    x = draws(mb, mb.monty_random_n_binomial, 1 << 19, float(size), prob, rt=rt)
    d = stats.binom(size, prob)
    lo, hi = int(d.ppf(1e-6)), int(d.ppf(1 - 1e-6))
    assert chisq_discrete(x, d.pmf, lo, hi) > 1e-3
    assert abs(x.mean() - size * prob) < 4 * np.sqrt(size * prob * (1 - prob) / len(x)) + 1e-9


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("lam", [0.5, 5.0, 20.0, 200.0, 5000.0])
def test_poisson_chisq(mb, lam, rt):                  # test-rng.R:206-258
    x = draws(mb, mb.monty_random_n_poisson, 1 << 19, lam, rt=rt)
    d = stats.poisson(lam)
    lo, hi = int(d.ppf(1e-6)), int(d.ppf(1 - 1e-6))
    assert chisq_discrete(x, d.pmf, lo, hi) > 1e-3
    assert abs(x.mean() - lam) < 5 * np.sqrt(lam / len(x))


def test_poisson_large_lambda_moments(mb):            # test-rng.R:236-258 (Cauchy envelope)
    for lam in (1e9, 1e12):
        x = draws(mb, mb.monty_random_n_poisson, 1 << 18, lam)
        assert abs(x.mean() / lam - 1) < 1e-3 / np.sqrt(lam / 1e9) + 1e-6
        assert abs(x.var() / lam - 1) < 2e-2


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("shape,scale", [(0.3, 2.0), (1.0, 0.5), (5.0, 1.0), (50.0, 0.1)])
def test_gamma_ks(mb, shape, scale, rt):              # test-rng.R:1032-1054
    x = draws(mb, mb.monty_random_n_gamma_scale, 1 << 19, shape, scale, rt=rt)
    assert stats.kstest(x, "gamma", args=(shape, 0, scale)).pvalue > 1e-3


@pytest.mark.parametrize("rt", ["f64", "f32"])
def test_beta_ks(mb, rt):                             # test-rng.R:1261-1268
    x = draws(mb, mb.monty_random_n_beta, 1 << 19, 2.5, 1.5, rt=rt)
    assert stats.kstest(x, "beta", args=(2.5, 1.5)).pvalue > 1e-3


@pytest.mark.parametrize("rt", ["f64", "f32"])
def test_negative_binomial_chisq(mb, rt):             # test-rng.R:1086-1095
    x = draws(mb, mb.monty_random_n_negative_binomial_prob, 1 << 19, 20.0, 0.3, rt=rt)
    d = stats.nbinom(20, 0.3)
    assert chisq_discrete(x, d.pmf, int(d.ppf(1e-6)), int(d.ppf(1 - 1e-6))) > 1e-3


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("n1,n2,k", [(7, 10, 8), (70, 100, 80), (700, 1000, 800)])
def test_hypergeometric_chisq(mb, n1, n2, k, rt):     # test-rng.R:879-929 parameter sets
    x = draws(mb, mb.monty_random_n_hypergeometric, 1 << 19, float(n1), float(n2), float(k), rt=rt)
    d = stats.hypergeom(n1 + n2, n1, k)
    assert chisq_discrete(x, d.pmf, int(d.ppf(1e-7)), int(d.ppf(1 - 1e-7))) > 1e-3


def test_all_values_appear(mb):                       # test-rng.R:1237-1243 regression
    x = draws(mb, mb.monty_random_n_hypergeometric, 1 << 16, 5.0, 5.0, 5.0)
    assert set(np.unique(x)) == set(range(6))


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("lo,hi", [(-1.0, 2.0), (2.0, 4.0), (1.5, np.inf), (-np.inf, -0.5)])
def test_truncated_normal_ks(mb, lo, hi, rt):         # test-rng.R:1300-1401
    x = draws(mb, mb.monty_random_n_truncated_normal, 1 << 18, 0.0, 1.0, lo, hi, rt=rt)
    assert x.min() >= lo and x.max() <= hi
    assert stats.kstest(x, "truncnorm", args=(lo, hi)).pvalue > 1e-3


def test_cauchy_weibull_lognormal_ks(mb):             # test-rng.R:1212-1219, 1404-1462
    x = draws(mb, mb.monty_random_n_cauchy, 1 << 18, 1.0, 2.0)
    assert stats.kstest(x, "cauchy", args=(1.0, 2.0)).pvalue > 1e-3
    x = draws(mb, mb.monty_random_n_weibull, 1 << 18, 1.7, 2.0)
    assert stats.kstest(x, "weibull_min", args=(1.7, 0, 2.0)).pvalue > 1e-3
    x = draws(mb, mb.monty_random_n_log_normal, 1 << 18, 0.3, 0.8)
    assert stats.kstest(x, "lognorm", args=(0.8, 0, np.exp(0.3))).pvalue > 1e-3
