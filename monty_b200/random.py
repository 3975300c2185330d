"""The `monty_rng_*` / `monty_random_*` interface of the reference, on B200.

Every function has the name, argument order and semantics of its R original
(ref: R/random.R, cited per function); where R returns a matrix with
`dim = c(n_samples, n_streams)` this returns a Fortran-ordered numpy array of
that shape (same memory layout as R: each stream's draws are contiguous).
Errors raise `MontyError` with the reference's message text.

State lives on the GPU inside a `MontyRngState` (the analogue of the R
external pointer with its attributes, R/random.R:59-72).
"""
import ctypes as C
import math
import random as _pyrandom

import numpy as np

from . import _lib
from ._lib import DIST, MontyError, check, lib

__all__ = [
    "MontyRngState", "monty_rng_create", "monty_rng_state", "monty_rng_set_state",
    "monty_rng_jump", "monty_rng_long_jump", "monty_rng_distributed_state",
    "monty_random_real", "monty_random_n_real",
    "monty_random_exponential_rate", "monty_random_n_exponential_rate",
    "monty_random_exponential_mean", "monty_random_n_exponential_mean",
    "monty_random_poisson", "monty_random_n_poisson",
    "monty_random_beta", "monty_random_n_beta",
    "monty_random_binomial", "monty_random_n_binomial",
    "monty_random_cauchy", "monty_random_n_cauchy",
    "monty_random_gamma_scale", "monty_random_n_gamma_scale",
    "monty_random_gamma_rate", "monty_random_n_gamma_rate",
    "monty_random_negative_binomial_prob", "monty_random_n_negative_binomial_prob",
    "monty_random_negative_binomial_mu", "monty_random_n_negative_binomial_mu",
    "monty_random_normal", "monty_random_n_normal",
    "monty_random_uniform", "monty_random_n_uniform",
    "monty_random_beta_binomial_prob", "monty_random_n_beta_binomial_prob",
    "monty_random_beta_binomial_ab", "monty_random_n_beta_binomial_ab",
    "monty_random_hypergeometric", "monty_random_n_hypergeometric",
    "monty_random_truncated_normal", "monty_random_n_truncated_normal",
    "monty_random_weibull", "monty_random_n_weibull",
    "monty_random_log_normal", "monty_random_n_log_normal",
    "monty_random_zi_poisson", "monty_random_n_zi_poisson",
    "monty_random_zi_negative_binomial_prob", "monty_random_n_zi_negative_binomial_prob",
    "monty_random_zi_negative_binomial_mu", "monty_random_n_zi_negative_binomial_mu",
    "monty_random_multinomial", "monty_random_n_multinomial",
    "test_xoshiro_run",
]


class MontyRngState:
    """Handle on a device-resident multi-stream generator.

    ref: the `monty_rng_state` object of R/random.R:59-72 (external pointer +
    attributes n_streams, n_threads, n_deterministic,
    preserve_stream_dimension).
    """

    def __init__(self, handle, n_streams, n_threads, deterministic,
                 preserve_stream_dimension, real_type="f64"):
        self._h = handle
        self.n_streams = n_streams
        self.n_threads = n_threads
        self.n_deterministic = deterministic
        self.preserve_stream_dimension = preserve_stream_dimension
        self.real_type = real_type

    def __len__(self):                       # R/random.R:183-185
        return self.n_streams

    def __repr__(self):                      # R/random.R:175-180
        return ("<monty_rng_state>\n * %d random number stream%s\n * %d execution thread%s"
                % (self.n_streams, "" if self.n_streams == 1 else "s",
                   self.n_threads, "" if self.n_threads == 1 else "s"))

    @property
    def handle(self):
        if not self._h:
            # ref: src/random.cpp:12-19
            raise MontyError("Pointer has been serialised, cannot continue safely (handle)")
        return self._h

    @property
    def device(self):
        return lib.mb200_rng_device(self.handle)

    def state_ptr(self):
        """Device address of the live state array, uint64[n_streams][4]."""
        return lib.mb200_rng_state_dev(self.handle)

    def cuda_stream(self):
        return lib.mb200_rng_cuda_stream(self.handle)

    def sync(self):
        check(lib.mb200_rng_sync(self.handle))

    def close(self):
        if self._h:
            lib.mb200_rng_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self):
        # like an R external pointer, the handle does not survive serialisation
        d = dict(self.__dict__)
        d["_h"] = None
        return d

    # ---- device-resident sampling (no host copies): the dust2-style seam ----
    def sample_dev(self, dist, n_samples, params=(), out=None, real_type=None):
        """Draw into a torch CUDA tensor `out` of shape (n_streams, n_samples)
        (float64, or float32 with real_type "f32").  `params` are torch CUDA
        float64 tensors of 1 or n_streams elements.  Asynchronous; call
        `sync()` to wait and collect parameter errors.

        Ordering.  The launch runs on the handle's own (non-blocking) CUDA
        stream.  It is ordered AFTER everything queued so far on torch's current
        stream (so parameters produced by pending torch kernels are complete,
        and a freshly allocated `out` is no longer in use by earlier work), and
        torch's current stream is made to wait for it, so torch operations
        issued afterwards on `out` see the draws."""
        import torch
        d = DIST[dist] if isinstance(dist, str) else dist
        rt = real_type or self.real_type
        n = self.n_streams
        out_f32 = 0
        if out is None:
            out = torch.empty((n, n_samples), dtype=torch.float64, device="cuda:%d" % self.device)
        if out.dtype == torch.float32:
            out_f32 = 1
        elif out.dtype != torch.float64:
            raise TypeError("out must be float64 or float32")
        if out.numel() != n * n_samples or not out.is_contiguous():
            raise ValueError("out must be contiguous with n_streams * n_samples elements")
        ptrs = (C.c_void_p * 4)()
        lens = (C.c_size_t * 4)()
        keep = []
        for k, p in enumerate(params):
            if p.dtype != torch.float64 or not p.is_cuda or not p.is_contiguous():
                raise TypeError("parameters must be contiguous float64 CUDA tensors")
            keep.append(p)
            ptrs[k] = p.data_ptr()
            lens[k] = p.numel()
        self._keepalive = (keep, out)
        hs = torch.cuda.ExternalStream(self.cuda_stream(), device=out.device)
        cur = torch.cuda.current_stream(out.device)
        ordered = cur.cuda_stream != hs.cuda_stream
        if ordered:
            hs.wait_stream(cur)
        check(lib.mb200_rng_sample_dev(self.handle, d, 1 if rt == "f32" else 0, n_samples,
                                       ptrs, lens, C.c_void_p(out.data_ptr()), out_f32))
        if ordered:
            # torch's stream now waits for the draws, so stream order also protects the
            # tensors' memory: the caching allocator reuses a freed block on the stream it
            # was allocated on, i.e. after this wait
            cur.wait_stream(hs)
        return out


def _seed_args(seed):
    """ref: as_rng_seed, inst/include/monty/r/random.hpp:47-67."""
    if seed is None:
        # NULL: ceil(|unif_rand()| * SIZE_MAX) from the host language's RNG
        u = _pyrandom.random()
        return int(math.ceil(abs(u) * (2 ** 64 - 1))) & (2 ** 64 - 1), None
    if isinstance(seed, (bytes, bytearray, memoryview)):
        raw = np.frombuffer(bytes(seed), dtype=np.uint8)
        return 0, raw
    if isinstance(seed, np.ndarray) and seed.dtype == np.uint8:
        return 0, np.ascontiguousarray(seed)
    if isinstance(seed, (int, np.integer)):
        return int(seed) & (2 ** 64 - 1), None
    if isinstance(seed, (float, np.floating)):
        return int(seed) & (2 ** 64 - 1), None
    raise MontyError("Invalid type for 'seed'")


def monty_rng_create(n_streams=1, seed=None, n_threads=1, deterministic=False,
                     preserve_stream_dimension=False, device=0, stream_offset=0,
                     long_jumps=0, real_type="f64"):
    """ref: R/random.R:59-72, src/random.cpp:341-353.

    `device`, `stream_offset`, `long_jumps` and `real_type` are extensions for
    multi-GPU sharding and the float path; defaults reproduce the reference."""
    if not isinstance(deterministic, (bool, np.bool_)):
        raise MontyError("Expected 'deterministic' to be a scalar logical")
    if not isinstance(preserve_stream_dimension, (bool, np.bool_)):
        raise MontyError("Expected 'preserve_stream_dimension' to be a scalar logical")
    preserve = bool(preserve_stream_dimension) or n_streams > 1
    seed_int, raw = _seed_args(seed)
    h = C.c_void_p()
    if raw is None:
        rc = lib.mb200_rng_create(seed_int, None, 0, n_streams, int(deterministic),
                                  stream_offset, long_jumps, device, C.byref(h))
    else:
        rc = lib.mb200_rng_create(0, raw.ctypes.data_as(C.c_void_p), raw.size, n_streams,
                                  int(deterministic), stream_offset, long_jumps, device,
                                  C.byref(h))
    check(rc)
    return MontyRngState(h, int(n_streams), n_threads, bool(deterministic), preserve, real_type)


def monty_rng_state(state):
    """ref: R/random.R:96-98, src/random.cpp:356-364 -> bytes, 32 per stream."""
    buf = np.empty(32 * state.n_streams, dtype=np.uint8)
    check(lib.mb200_rng_state(state.handle, buf.ctypes.data_as(C.c_void_p), buf.size))
    return buf.tobytes()


def monty_rng_set_state(value, state):
    """ref: R/random.R:110-112, src/random.cpp:366-379."""
    raw = np.frombuffer(bytes(value), dtype=np.uint8)
    check(lib.mb200_rng_set_state(state.handle, raw.ctypes.data_as(C.c_void_p), raw.size))


def _assert_scalar_size(n, name="n"):
    if not isinstance(n, (int, np.integer)) or isinstance(n, bool) or n < 1:
        raise MontyError("'%s' must be at least 1" % name)


def monty_rng_jump(state, n=1):
    """ref: R/random.R:131-141."""
    if isinstance(state, (bytes, bytearray)):
        rng = monty_rng_create(seed=bytes(state), n_streams=len(state) // 32)
        return monty_rng_state(monty_rng_jump(rng, n))
    if not isinstance(state, MontyRngState):
        raise MontyError("Expected 'state' to be a 'monty_rng_state'")
    _assert_scalar_size(n)
    check(lib.mb200_rng_jump(state.handle, int(n)))
    return state


def monty_rng_long_jump(state, n=1):
    """ref: R/random.R:146-156."""
    if isinstance(state, (bytes, bytearray)):
        rng = monty_rng_create(seed=bytes(state), n_streams=len(state) // 32)
        return monty_rng_state(monty_rng_long_jump(rng, n))
    if not isinstance(state, MontyRngState):
        raise MontyError("Expected 'state' to be a 'monty_rng_state'")
    _assert_scalar_size(n)
    check(lib.mb200_rng_long_jump(state.handle, int(n)))
    return state


def monty_rng_distributed_state(seed=None, n_nodes=1):
    """ref: R/random.R:161-171: the long-jump ladder handed to worker nodes."""
    rng = monty_rng_create(seed=seed)
    ret = []
    for i in range(n_nodes):
        ret.append(monty_rng_state(rng))
        if i < n_nodes - 1:
            monty_rng_long_jump(rng)
    return ret


def _sample(state, dist, n_samples, params):
    n_streams = state.n_streams
    one = n_samples is None
    ns = 1 if one else int(n_samples)
    arrs = [np.ascontiguousarray(np.atleast_1d(np.asarray(p, dtype=np.float64))) for p in params]
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs] + [None] * (4 - len(arrs)))
    lens = (C.c_size_t * 4)(*[a.size for a in arrs] + [0] * (4 - len(arrs)))
    out = np.empty(ns * n_streams, dtype=np.float64)
    rc = lib.mb200_rng_sample(state.handle, DIST[dist], 1 if state.real_type == "f32" else 0,
                              ns, ptrs, lens, out.ctypes.data_as(C.c_void_p))
    check(rc)
    if one:
        return out                               # src/random.cpp:89-195
    if n_streams > 1 or state.preserve_stream_dimension:   # src/random.cpp:197-201
        return out.reshape((n_streams, ns)).T
    return out


def _make(dist, names):
    def one(*args):
        *params, state = args
        if len(params) != len(names):
            raise TypeError("expected arguments (%s, state)" % ", ".join(names))
        return _sample(state, dist, None, params)

    def many(n_samples, *args):
        *params, state = args
        if len(params) != len(names):
            raise TypeError("expected arguments (n_samples, %s, state)" % ", ".join(names))
        return _sample(state, dist, n_samples, params)

    one.__doc__ = "ref: monty_random_%s(%s, state), R/random.R" % (dist, ", ".join(names))
    many.__doc__ = "ref: monty_random_n_%s(n_samples, %s, state), R/random.R" % (dist, ", ".join(names))
    return one, many


monty_random_real, monty_random_n_real = _make("real", [])
monty_random_exponential_rate, monty_random_n_exponential_rate = _make("exponential_rate", ["rate"])
monty_random_exponential_mean, monty_random_n_exponential_mean = _make("exponential_mean", ["mean"])
monty_random_poisson, monty_random_n_poisson = _make("poisson", ["lambda"])
monty_random_beta, monty_random_n_beta = _make("beta", ["a", "b"])
monty_random_binomial, monty_random_n_binomial = _make("binomial", ["size", "prob"])
monty_random_cauchy, monty_random_n_cauchy = _make("cauchy", ["location", "scale"])
monty_random_gamma_scale, monty_random_n_gamma_scale = _make("gamma_scale", ["shape", "scale"])
monty_random_gamma_rate, monty_random_n_gamma_rate = _make("gamma_rate", ["shape", "rate"])
monty_random_negative_binomial_prob, monty_random_n_negative_binomial_prob = _make(
    "negative_binomial_prob", ["size", "prob"])
monty_random_negative_binomial_mu, monty_random_n_negative_binomial_mu = _make(
    "negative_binomial_mu", ["size", "mu"])
monty_random_uniform, monty_random_n_uniform = _make("uniform", ["min", "max"])
monty_random_beta_binomial_prob, monty_random_n_beta_binomial_prob = _make(
    "beta_binomial_prob", ["size", "prob", "rho"])
monty_random_beta_binomial_ab, monty_random_n_beta_binomial_ab = _make(
    "beta_binomial_ab", ["size", "a", "b"])
monty_random_hypergeometric, monty_random_n_hypergeometric = _make(
    "hypergeometric", ["n1", "n2", "k"])
monty_random_truncated_normal, monty_random_n_truncated_normal = _make(
    "truncated_normal", ["mean", "sd", "min", "max"])
monty_random_weibull, monty_random_n_weibull = _make("weibull", ["shape", "scale"])
monty_random_log_normal, monty_random_n_log_normal = _make("log_normal", ["meanlog", "sdlog"])
monty_random_zi_poisson, monty_random_n_zi_poisson = _make("zi_poisson", ["pi0", "lambda"])
monty_random_zi_negative_binomial_prob, monty_random_n_zi_negative_binomial_prob = _make(
    "zi_negative_binomial_prob", ["pi0", "size", "prob"])
monty_random_zi_negative_binomial_mu, monty_random_n_zi_negative_binomial_mu = _make(
    "zi_negative_binomial_mu", ["pi0", "size", "mu"])

_NORMAL = {"box_muller": "normal_box_muller", "polar": "normal_polar", "ziggurat": "normal_ziggurat"}


def monty_random_normal(mean, sd, state, algorithm="box_muller"):
    """ref: R/random.R:475-482."""
    if algorithm not in _NORMAL:
        raise MontyError("Unknown normal algorithm '%s'" % algorithm)
    return _sample(state, _NORMAL[algorithm], None, [mean, sd])


def monty_random_n_normal(n_samples, mean, sd, state, algorithm="box_muller"):
    """ref: R/random.R:486-496."""
    if algorithm not in _NORMAL:
        raise MontyError("Unknown normal algorithm '%s'" % algorithm)
    return _sample(state, _NORMAL[algorithm], n_samples, [mean, sd])


def _multinomial(state, n_samples, size, prob):
    n_streams = state.n_streams
    one = n_samples is None
    ns = 1 if one else int(n_samples)
    size = np.ascontiguousarray(np.atleast_1d(np.asarray(size, dtype=np.float64)))
    prob = np.asarray(prob, dtype=np.float64)
    if prob.ndim == 1:                       # src/random.cpp:52-55: no dim -> shared vector
        prob_len, prob_cols = prob.size, 1
        pf = np.ascontiguousarray(prob)
    elif prob.ndim == 2:                     # R matrix [len x cols], column-major
        prob_len, prob_cols = prob.shape
        pf = np.ascontiguousarray(prob.T)
    else:
        raise MontyError("Expected 'prob' to be a matrix")
    out = np.empty(n_streams * ns * prob_len, dtype=np.float64)
    rc = lib.mb200_rng_multinomial(state.handle, ns, size.ctypes.data_as(C.c_void_p), size.size,
                                   pf.ctypes.data_as(C.c_void_p), prob_len, prob_cols,
                                   out.ctypes.data_as(C.c_void_p))
    check(rc)
    preserve = n_streams > 1 or state.preserve_stream_dimension
    if one:                                  # src/random.cpp:884-905: dim (len, n_streams)
        return out.reshape((n_streams, prob_len)).T if preserve else out
    full = out.reshape((n_streams, ns, prob_len)).transpose(2, 1, 0)   # (len, n_samples, n_streams)
    return full if preserve else full[:, :, 0]


def monty_random_multinomial(size, prob, state):
    """ref: R/random.R (monty_random_multinomial), src/random.cpp:884-905."""
    return _multinomial(state, None, size, prob)


def monty_random_n_multinomial(n_samples, size, prob, state):
    """ref: src/random.cpp:908-934."""
    return _multinomial(state, n_samples, size, prob)


def test_xoshiro_run(obj=None, generator="xoshiro256plus", seed=42, n=10, device=0):
    """ref: src/test_rng.cpp:21-38 and tests/testthat/test-xoshiro.R.

    With a `MontyRngState` runs the protocol's raw part on its first stream is
    not needed: the protocol is run stand-alone on the device for any of the
    12 generators.  Returns (values as decimal strings, final state words as
    hex strings), the two halves of a tests/golden/xoshiro-ref file."""
    g = _lib.GENERATORS.index(generator)
    vals = np.zeros(3 * n, dtype=np.uint64)
    st = np.zeros(8, dtype=np.uint64)
    nw = C.c_int(0)
    check(lib.mb200_xoshiro_run(g, seed, n, device, vals.ctypes.data_as(C.c_void_p),
                                st.ctypes.data_as(C.c_void_p), C.byref(nw)))
    width = 8 if generator.startswith("xoshiro128") else 16
    return [str(int(v)) for v in vals], [format(int(w), "0%dx" % width) for w in st[:nw.value]]
