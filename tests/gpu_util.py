"""Helpers for the GPU parity tests."""
import numpy as np


def states_of(mb, rng):
    return np.frombuffer(mb.monty_rng_state(rng), dtype=np.uint64).reshape(-1, 4).copy()


def ulp_distance(a, b):
    """Distance in units in the last place between two float64 arrays."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ia = a.view(np.int64).astype(np.int64)
    ib = b.view(np.int64).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2 ** 63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2 ** 63) - ib, ib)
    with np.errstate(over="ignore"):
        d = np.abs(ia - ib).astype(np.float64)      # exact: subtract as integers first
    d[(a == b) | (np.isnan(a) & np.isnan(b))] = 0
    return d


def gpu_sample(mb, name, n_streams, n_samples, pars, seed=7, rt="f64", **kw):
    """(draws as (n_streams, n_samples), final states, handle)."""
    rng = mb.monty_rng_create(n_streams, seed, real_type=rt, **kw)
    y = mb.random._sample(rng, name, n_samples, pars)
    y = np.ascontiguousarray(np.asarray(y).reshape(n_samples, n_streams).T) if n_streams > 1 \
        else np.asarray(y).reshape(1, n_samples)
    return y, states_of(mb, rng), rng
