"""ctypes binding of monty_b200/libmonty_b200_examples.so: consumers of the
device-side monty::random API (include/monty_b200/random.cuh), see
examples/examples.h.  Like the main library there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

from ._lib import DIST, GENERATORS, check, lib as _main  # noqa: F401  (loads libmonty_b200.so first)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmonty_b200_examples.so")
if not os.path.exists(LIB_PATH):
    raise ImportError("%s is missing: build it with `python __graft_entry__.py`" % LIB_PATH)
lib = C.CDLL(LIB_PATH)
lib.mb200_example_inline_draw.restype = C.c_int
lib.mb200_example_sir_filter.restype = C.c_int
lib.mb200_example_math_probe.restype = C.c_int
lib.mb200_example_generator_draw.restype = C.c_int
lib.mb200_example_generator_seed_vector.restype = C.c_int

lib.mb200_example_inline_multinomial.restype = C.c_int


def inline_multinomial(state, size, prob):
    """One multinomial draw per stream through monty::random::multinomial() inside
    a kernel.  prob: (prob_len,) shared or (n_streams, prob_len); returns
    (n_streams, prob_len)."""
    size = np.ascontiguousarray(np.atleast_1d(np.asarray(size, dtype=np.float64)))
    prob = np.ascontiguousarray(np.asarray(prob, dtype=np.float64))
    cols = 1 if prob.ndim == 1 else prob.shape[0]
    prob_len = prob.shape[-1]
    out = np.empty((state.n_streams, prob_len), dtype=np.float64)
    check(lib.mb200_example_inline_multinomial(
        state.handle, C.c_int(1 if state.real_type == "f32" else 0),
        size.ctypes.data_as(C.c_void_p), C.c_size_t(size.size),
        prob.ctypes.data_as(C.c_void_p), C.c_int(prob_len), C.c_size_t(cols),
        out.ctypes.data_as(C.c_void_p)))
    return out


MATH_FN = {"log": 0, "log_normal": 1, "cos_0_2pi": 2, "sqrt_normal": 3, "sqrt": 4, "contraction_is_off": 5, "lgamma": 6, "exp": 7, "pow": 8, "tan": 9, "log1p": 10, "erfc": 11, "tanf": 12}


def math_probe(fn, x):
    """The samplers' own math routines (include/monty_b200/detail/samplers.cuh) evaluated on the device."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    check(lib.mb200_example_math_probe(C.c_int(MATH_FN[fn]), x.ctypes.data_as(C.c_void_p),
                                       y.ctypes.data_as(C.c_void_p), C.c_size_t(x.size)))
    return y


def zi_density(which, x, pi0, a, b=None, log=True):
    """monty::density::zi_<which> evaluated through the device-side API."""
    idx = {"poisson": 0, "negative_binomial_prob": 1, "negative_binomial_mu": 2}[which]
    arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, pi0, a)]
    bb = np.ascontiguousarray(b, dtype=np.float64) if b is not None else None
    y = np.empty_like(arrs[0])
    check(lib.mb200_example_zi_density(C.c_int(idx), arrs[0].ctypes.data_as(C.c_void_p),
                                       arrs[1].ctypes.data_as(C.c_void_p), arrs[2].ctypes.data_as(C.c_void_p),
                                       bb.ctypes.data_as(C.c_void_p) if bb is not None else None,
                                       C.c_int(1 if log else 0), y.ctypes.data_as(C.c_void_p), C.c_size_t(y.size)))
    return y


def inline_draw(state, dist, params=()):
    """One draw per stream through monty::random::<dist>() inside a kernel."""
    arrs = [np.ascontiguousarray(np.atleast_1d(np.asarray(p, dtype=np.float64))) for p in params]
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs] + [None] * (4 - len(arrs)))
    lens = (C.c_size_t * 4)(*[a.size for a in arrs] + [0] * (4 - len(arrs)))
    out = np.empty(state.n_streams, dtype=np.float64)
    check(lib.mb200_example_inline_draw(state.handle, C.c_int(DIST[dist]),
                                        C.c_int(1 if state.real_type == "f32" else 0),
                                        ptrs, lens, out.ctypes.data_as(C.c_void_p)))
    return out


def sir_filter(system, filt, pars, data_time, data_incidence, resample=True, return_state=False):
    """Bootstrap particle filter of the SIR model of the reference's test-suite
    (tests/testthat/helper-sir-filter.R): `system` has one stream per particle,
    `filt` one stream.  pars: dict(N, I0, beta, gamma, exp_noise)."""
    t = np.ascontiguousarray(data_time, dtype=np.float64)
    y = np.ascontiguousarray(data_incidence, dtype=np.float64)
    ll = C.c_double(0.0)
    n = system.n_streams
    final = np.empty((4, n), dtype=np.float64) if return_state else None
    check(lib.mb200_example_sir_filter(
        system.handle, filt.handle, C.c_double(pars["N"]), C.c_double(pars["I0"]),
        C.c_double(pars["beta"]), C.c_double(pars["gamma"]), C.c_double(pars["exp_noise"]),
        t.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_size_t(t.size),
        C.c_int(1 if resample else 0), C.byref(ll),
        final.ctypes.data_as(C.c_void_p) if final is not None else None))
    return (ll.value, final) if return_state else ll.value


# state words per stream and their width, by generator family
GEN_LAYOUT = {"xoshiro256": (4, np.uint64), "xoshiro128": (4, np.uint32), "xoroshiro128": (2, np.uint64),
              "xoshiro512": (8, np.uint64)}


def generator_layout(gen):
    for scr in ("plusplus", "starstar", "plus"):
        if gen.endswith(scr):
            return GEN_LAYOUT[gen[:-len(scr)]]
    raise ValueError(gen)


def generator_draw(gen, dist, seed, n_streams, n_samples, params=(), real_type="f64", jumps=0, long_jumps=0):
    """The device-side API on another generator (examples/generator_draw.cu): streams seeded
    like the reference's prng<gen>(n_streams, seed), then jumps, then n_samples draws of
    `dist` per stream through monty::random::<dist><real_type>(state, ...).  Returns
    (draws (n_streams, n_samples), final states (n_streams, words))."""
    arrs = [np.ascontiguousarray(np.atleast_1d(np.asarray(p, dtype=np.float64))) for p in params]
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs] + [None] * (4 - len(arrs)))
    lens = (C.c_size_t * 4)(*[a.size for a in arrs] + [0] * (4 - len(arrs)))
    words, wt = generator_layout(gen)
    out = np.empty((n_streams, n_samples), dtype=np.float64)
    st = np.empty((n_streams, words), dtype=wt)
    check(lib.mb200_example_generator_draw(
        C.c_int(GENERATORS.index(gen)), C.c_int(DIST[dist]), C.c_int(1 if real_type == "f32" else 0),
        C.c_uint64(seed), C.c_size_t(n_streams), C.c_size_t(n_samples), C.c_int(jumps), C.c_int(long_jumps),
        ptrs, lens, out.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p)))
    return out, st


def generator_seed_vector(gen, words, n_streams, reimport_and_jump=False):
    """monty_b200::device_prng<gen>(n_streams, seed words): the reference's seed-vector
    constructor; optionally export -> import into a second container -> jump -> export."""
    n_words, wt = generator_layout(gen)
    w = np.ascontiguousarray(words, dtype=wt)
    st = np.empty((n_streams, n_words), dtype=wt)
    check(lib.mb200_example_generator_seed_vector(C.c_int(GENERATORS.index(gen)), w.ctypes.data_as(C.c_void_p),
                                                  C.c_size_t(w.size), C.c_size_t(n_streams),
                                                  C.c_int(1 if reimport_and_jump else 0), st.ctypes.data_as(C.c_void_p)))
    return st
