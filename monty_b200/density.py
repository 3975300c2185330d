"""The vectorised (log-)densities of the reference's src/density.cpp, on B200.

Each `density_<dist>(x, params..., log)` mirrors the R-internal function of the
same name (ref: R/cpp11.R:3-85, src/density.cpp:6-254): every argument is a
vector indexed by i (no recycling), integer-valued arguments of the count
distributions are passed as int32 exactly where the reference takes
`cpp11::integers`, the result is a float64 vector of the same length.

`log_likelihood(...)` is the fused reduce-only form used for config C5: the
per-observation values never leave the GPU, only their sum does.
"""
import ctypes as C

import numpy as np

from ._lib import DENS, DENS_N_PARAMS, check, lib

# which arguments the reference takes as cpp11::integers (x, then parameters)
_INT_ARGS = {
    "binomial": (True, [True, False]),
    "negative_binomial_mu": (True, [False, False]),
    "negative_binomial_prob": (True, [False, False]),
    "beta_binomial_prob": (True, [True, False, False]),
    "beta_binomial_ab": (True, [True, False, False]),
    "poisson": (True, [False]),
    "hypergeometric": (True, [True, True, True]),
    "zi_poisson": (True, [False, False]),
    "zi_negative_binomial_mu": (True, [False, False, False]),
    "zi_negative_binomial_prob": (True, [False, False, False]),
}


def _prep(a, as_int):
    a = np.asarray(a)
    if as_int:
        return np.ascontiguousarray(a, dtype=np.int32), 1
    return np.ascontiguousarray(a, dtype=np.float64), 0


def _density(name, x, params, log, real_type="f64", device=0, want_out=True, want_sum=False):
    d = DENS[name]
    x_int, p_int = _INT_ARGS.get(name, (False, [False] * DENS_N_PARAMS[d]))
    xa, xi = _prep(x, x_int)
    n = xa.size
    arrs = [_prep(p, i) for p, i in zip(params, p_int)]
    for a, _ in arrs:
        if a.size < n:
            raise ValueError("density arguments must all have the length of x (no recycling)")
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a, _ in arrs] + [None] * (4 - len(arrs)))
    isi = (C.c_int * 4)(*[i for _, i in arrs] + [0] * (4 - len(arrs)))
    out = np.empty(n, dtype=np.float64) if want_out else None
    total = C.c_double(0.0)
    rc = lib.mb200_density(d, 1 if real_type == "f32" else 0, n,
                           C.c_void_p(xa.ctypes.data), xi, ptrs, isi, 1 if log else 0,
                           C.c_void_p(out.ctypes.data) if want_out else None,
                           C.byref(total) if want_sum else None, device)
    check(rc)
    if want_out and want_sum:
        return out, total.value
    return out if want_out else total.value


def log_likelihood(name, x, params, real_type="f64", device=0):
    """Sum of log densities, reduced on the device (SURVEY config C5)."""
    return _density(name, x, params, True, real_type, device, want_out=False, want_sum=True)


def density_binomial(x, size, prob, log):
    return _density("binomial", x, [size, prob], log)


def density_normal(x, mu, sd, log):
    return _density("normal", x, [mu, sd], log)


def density_negative_binomial_mu(x, size, mu, log, is_float=False):
    return _density("negative_binomial_mu", x, [size, mu], log, "f32" if is_float else "f64")


def density_negative_binomial_prob(x, size, prob, log):
    return _density("negative_binomial_prob", x, [size, prob], log)


def density_beta_binomial_prob(x, size, prob, rho, log):
    return _density("beta_binomial_prob", x, [size, prob, rho], log)


def density_beta_binomial_ab(x, size, a, b, log):
    return _density("beta_binomial_ab", x, [size, a, b], log)


def density_poisson(x, lambda_, log):
    return _density("poisson", x, [lambda_], log)


def density_uniform(x, min, max, log):
    return _density("uniform", x, [min, max], log)


def density_exponential_rate(x, rate, log):
    return _density("exponential_rate", x, [rate], log)


def density_exponential_mean(x, mean, log):
    return _density("exponential_mean", x, [mean], log)


def density_gamma_rate(x, shape, rate, log):
    return _density("gamma_rate", x, [shape, rate], log)


def density_gamma_scale(x, shape, scale, log):
    return _density("gamma_scale", x, [shape, scale], log)


def density_beta(x, a, b, log):
    return _density("beta", x, [a, b], log)


def density_hypergeometric(x, n1, n2, k, log):
    return _density("hypergeometric", x, [n1, n2, k], log)


def density_log_normal(x, mulog, sdlog, log):
    return _density("log_normal", x, [mulog, sdlog], log)


def density_truncated_normal(x, mu, sd, min, max, log):
    return _density("truncated_normal", x, [mu, sd, min, max], log)


def density_cauchy(x, location, scale, log):
    return _density("cauchy", x, [location, scale], log)


def density_weibull(x, shape, scale, log):
    return _density("weibull", x, [shape, scale], log)


def density_zi_poisson(x, pi0, lambda_, log):
    return _density("zi_poisson", x, [pi0, lambda_], log)


def density_zi_negative_binomial_mu(x, pi0, size, mu, log):
    return _density("zi_negative_binomial_mu", x, [pi0, size, mu], log)


def density_zi_negative_binomial_prob(x, pi0, size, prob, log):
    return _density("zi_negative_binomial_prob", x, [pi0, size, prob], log)
