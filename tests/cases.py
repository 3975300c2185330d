"""Seed-defined parameter sets shared by the golden-vector generator, the
oracle tests and the GPU parity tests (the counter hash h(i, k))."""
import numpy as np

from oracle_lib import counter_hash

INTEGER_VALUED = {
    "binomial", "poisson", "negative_binomial_prob", "negative_binomial_mu",
    "hypergeometric", "beta_binomial_prob", "beta_binomial_ab", "zi_poisson",
    "zi_negative_binomial_prob", "zi_negative_binomial_mu",
}


def sampler_cases(n, first=0):
    """Per-stream parameter vectors spanning every regime of every sampler."""
    idx = np.arange(first, first + n)
    h = lambda k: counter_hash(idx, k)  # noqa: E731
    n1 = np.floor(2000 * h(0))
    n2 = np.floor(2000 * h(1))
    return {
        "real": [],
        "uniform": [h(0) * 10 - 5, h(1) * 3 + 6],
        "exponential_rate": [h(0) * 5 + 0.1],
        "exponential_mean": [h(0) * 5 + 0.1],
        "normal_box_muller": [h(0) * 10 - 5, 0.5 + h(1)],
        "normal_polar": [h(0) * 10 - 5, 0.5 + h(1)],
        "normal_ziggurat": [h(0) * 10 - 5, 0.5 + h(1)],
        # size 1..1e6 x prob: inversion (n q < 10) and BTRS regimes, both mirrors
        "binomial": [np.floor(10 ** (6 * h(0))), h(1)],
        # Knuth (< 10), Hormann (< 1e8), every 16th stream Cauchy (>= 1e8)
        "poisson": [np.where(idx % 16 == 0, 1e8 * (1 + h(0)), 10 ** (-1 + 5 * h(2)))],
        "gamma_scale": [np.where(idx % 11 == 0, 1.0, 10 ** (-1 + 2.5 * h(0))), 0.5 + h(1)],
        "gamma_rate": [10 ** (-1 + 2.5 * h(0)), 0.5 + h(1)],
        "beta": [0.2 + 20 * h(0), 0.2 + 20 * h(1)],
        "negative_binomial_prob": [0.5 + 100 * h(0), 0.05 + 0.9 * h(1)],
        "negative_binomial_mu": [0.5 + 100 * h(0), 100 * h(1)],
        # HIN, H2PE-recursive and H2PE-squeeze
        "hypergeometric": [n1, n2, np.floor((n1 + n2) * h(2))],
        "truncated_normal": [h(3) - 0.5, 0.5 + h(2),
                             np.where(idx % 16 == 0, -np.inf, -3 + 5 * h(0)),
                             np.where(idx % 8 == 1, np.inf, -3 + 5 * h(0) + 0.1 + 4 * h(1))],
        "cauchy": [h(0), h(1) + 0.1],
        "weibull": [h(0) * 3 + 0.2, h(1) + 0.1],
        "log_normal": [h(0), h(1) + 0.1],
        "beta_binomial_prob": [np.floor(100 * h(0)), 0.05 + 0.9 * h(1), 0.01 + 0.49 * h(2)],
        "beta_binomial_ab": [np.floor(100 * h(0)), 0.2 + 5 * h(1), 0.2 + 5 * h(2)],
        "zi_poisson": [np.where(idx % 7 == 0, 0.0, np.where(idx % 7 == 1, 1.0, h(0))),
                       10 ** (-1 + 3 * h(1))],
        "zi_negative_binomial_prob": [h(0), 0.5 + 50 * h(1), 0.05 + 0.9 * h(2)],
        "zi_negative_binomial_mu": [h(0), 0.5 + 50 * h(1), 100 * h(2)],
    }


def standard_cases(n):
    """Well-conditioned parameters (location 0 / scale 1) on which the ulp
    distance of continuous draws is meaningful: a shifted draw z*sd+mean near
    zero turns a 1-ulp difference in z into an arbitrary ulp distance."""
    one, zero = np.array([1.0]), np.array([0.0])
    h = lambda k: counter_hash(np.arange(n), k)  # noqa: E731
    return {
        "exponential_rate": [one], "exponential_mean": [one],
        "normal_box_muller": [zero, one], "normal_polar": [zero, one],
        "normal_ziggurat": [zero, one], "cauchy": [zero, one],
        "weibull": [1.0 + 2 * h(0), one], "log_normal": [zero, one],
        "gamma_scale": [2.0 + 10 * h(0), one], "beta": [2.0 + 5 * h(0), 2.0 + 5 * h(1)],
    }


def density_cases(n):
    idx = np.arange(n)
    h = lambda k: counter_hash(idx, k)  # noqa: E731
    xi = np.floor(200 * h(0) ** 2).astype(np.int32)
    sz = (xi + np.floor(100 * h(1))).astype(np.int32)
    xd = h(4) * 5
    return {
        "binomial": (xi, [sz, h(2)]),
        "normal": (xd, [h(1), h(2) + 0.1]),
        "negative_binomial_mu": (xi, [0.5 + 50 * h(1), 0.1 + 100 * h(2)]),
        "negative_binomial_prob": (xi, [0.5 + 50 * h(1), 0.05 + 0.9 * h(2)]),
        "beta_binomial_prob": (xi, [sz, 0.05 + 0.9 * h(2), 0.01 + 0.49 * h(3)]),
        "beta_binomial_ab": (xi, [sz, 0.2 + 5 * h(2), 0.2 + 5 * h(3)]),
        "poisson": (xi, [0.5 + 100 * h(1)]),
        "uniform": (xd, [h(1), h(2) * 6]),
        "exponential_rate": (xd, [h(1) + 0.1]),
        "exponential_mean": (xd, [h(1) + 0.1]),
        "gamma_rate": (xd, [h(1) * 4 + 0.1, h(2) + 0.1]),
        "gamma_scale": (xd, [h(1) * 4 + 0.1, h(2) + 0.1]),
        "beta": (h(4), [h(1) * 4 + 0.1, h(2) * 3 + 0.1]),
        "hypergeometric": (np.minimum(xi, 10).astype(np.int32),
                           [(sz + 10).astype(np.int32), (sz + 20).astype(np.int32),
                            np.full(n, 10, np.int32)]),
        "log_normal": (xd, [h(1), h(2) + 0.1]),
        "truncated_normal": (xd, [h(1), h(2) + 0.5, np.full(n, -1.0), np.full(n, 6.0)]),
        "cauchy": (xd, [h(1), h(2) + 0.1]),
        "weibull": (xd, [h(1) * 3 + 0.2, h(2) + 0.1]),
        "zi_poisson": (xi, [h(3), 0.5 + 100 * h(1)]),
        "zi_negative_binomial_mu": (xi, [h(3), 0.5 + 50 * h(1), 0.1 + 100 * h(2)]),
        "zi_negative_binomial_prob": (xi, [h(3), 0.5 + 50 * h(1), 0.05 + 0.9 * h(2)]),
    }
