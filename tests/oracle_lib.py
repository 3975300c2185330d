"""ctypes access to the parity checkers (TEST INFRASTRUCTURE).

`Oracle("port")`  -> oracle/libmonty_oracle.so   (plain-C restatement)
`Oracle("reference")` -> oracle/_ref/libmonty_ref.so (unmodified reference headers)

Both expose the same operations with the same semantics as the reference's
cpp11 layer (src/random.cpp, src/density.cpp); states are numpy uint64 arrays
of shape (n_streams, 4) updated in place.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

GEN_NAMES = [
    "xoshiro256plus", "xoshiro256plusplus", "xoshiro256starstar",
    "xoshiro128plus", "xoshiro128plusplus", "xoshiro128starstar",
    "xoroshiro128plus", "xoroshiro128plusplus", "xoroshiro128starstar",
    "xoshiro512plus", "xoshiro512plusplus", "xoshiro512starstar",
]

DIST = {
    "real": 0, "uniform": 1, "exponential_rate": 2, "exponential_mean": 3,
    "normal_box_muller": 4, "normal_polar": 5, "normal_ziggurat": 6,
    "binomial": 7, "poisson": 8, "gamma_scale": 9, "gamma_rate": 10,
    "beta": 11, "negative_binomial_prob": 12, "negative_binomial_mu": 13,
    "hypergeometric": 14, "truncated_normal": 15, "cauchy": 16, "weibull": 17,
    "log_normal": 18, "beta_binomial_prob": 19, "beta_binomial_ab": 20,
    "zi_poisson": 21, "zi_negative_binomial_prob": 22,
    "zi_negative_binomial_mu": 23,
}
N_PARAMS = [0, 2, 1, 1, 2, 2, 2, 2, 1, 2, 2, 2, 2, 2, 3, 4, 2, 2, 2, 3, 3, 2, 3, 3]

DENS = {
    "binomial": 0, "normal": 1, "negative_binomial_mu": 2,
    "negative_binomial_prob": 3, "beta_binomial_prob": 4,
    "beta_binomial_ab": 5, "poisson": 6, "uniform": 7, "exponential_rate": 8,
    "exponential_mean": 9, "gamma_rate": 10, "gamma_scale": 11, "beta": 12,
    "hypergeometric": 13, "log_normal": 14, "truncated_normal": 15,
    "cauchy": 16, "weibull": 17, "zi_poisson": 18,
    "zi_negative_binomial_mu": 19, "zi_negative_binomial_prob": 20,
}


def build_oracles():
    """(Re)build the checkers; `make ref` is a no-op without /root/reference."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def have_reference():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libmonty_ref.so"))


class OracleError(RuntimeError):
    pass


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            path = os.path.join(ORACLE_DIR, "libmonty_oracle.so")
            self.px = "orc_"
        elif kind == "reference":
            path = os.path.join(ORACLE_DIR, "_ref", "libmonty_ref.so")
            self.px = "ref_"
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            build_oracles()
        self.lib = C.CDLL(path)
        self._fn("last_error").restype = C.c_char_p
        self._fn("max_threads").restype = C.c_int

    def _fn(self, name):
        return getattr(self.lib, self.px + name)

    def _err(self):
        return self._fn("last_error")().decode()

    def max_threads(self):
        return int(self._fn("max_threads")())

    # -- generator ---------------------------------------------------------
    def seed_states(self, seed, n_streams, raw_words=None):
        out = np.zeros((n_streams, 4), dtype=np.uint64)
        if self.kind == "port":
            if raw_words is None:
                words, nw = None, 0
            else:
                raw_words = np.ascontiguousarray(raw_words, dtype=np.uint64).ravel()
                words, nw = raw_words.ctypes.data_as(C.c_void_p), raw_words.size
            rc = self.lib.orc_seed_states(C.c_uint64(seed), words, C.c_size_t(nw),
                                          C.c_size_t(n_streams),
                                          out.ctypes.data_as(C.c_void_p))
        elif raw_words is None:
            rc = self.lib.ref_seed_states(C.c_uint64(seed), C.c_size_t(n_streams),
                                          out.ctypes.data_as(C.c_void_p))
        else:
            raw_words = np.ascontiguousarray(raw_words, dtype=np.uint64).ravel()
            rc = self.lib.ref_seed_states_raw(raw_words.ctypes.data_as(C.c_void_p),
                                              C.c_size_t(raw_words.size),
                                              C.c_size_t(n_streams),
                                              out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise OracleError(self._err())
        return out

    def jump(self, state, n=1, long=False):
        assert state.dtype == np.uint64 and state.flags.c_contiguous
        self._fn("jump")(state.ctypes.data_as(C.c_void_p), C.c_size_t(state.shape[0]),
                         C.c_int(n), C.c_int(1 if long else 0))
        return state

    def raw(self, gen, state, n_samples):
        out = np.zeros((state.shape[0], n_samples), dtype=np.uint64)
        rc = self._fn("raw")(C.c_int(gen), state.ctypes.data_as(C.c_void_p),
                             C.c_size_t(state.shape[0]), C.c_size_t(n_samples),
                             out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise OracleError(self._err())
        return out

    def xoshiro_run(self, gen, seed=42, n=10):
        vals = np.zeros(3 * n, dtype=np.uint64)
        st = np.zeros(8, dtype=np.uint64)
        nw = C.c_int(0)
        rc = self._fn("xoshiro_run")(C.c_int(gen), C.c_uint64(seed), C.c_int(n),
                                     vals.ctypes.data_as(C.c_void_p),
                                     st.ctypes.data_as(C.c_void_p), C.byref(nw))
        if rc != 0:
            raise OracleError(self._err())
        return vals, st[:nw.value]

    # -- samplers ----------------------------------------------------------
    def sample(self, dist, state, n_samples, params=(), real_type="f64",
               deterministic=False, n_threads=1):
        """Returns array (n_streams, n_samples); `state` advanced in place.
        On a parameter error raises OracleError (state stays partially
        advanced, as in the reference)."""
        d = DIST[dist] if isinstance(dist, str) else dist
        n_streams = state.shape[0]
        arrs = [np.ascontiguousarray(np.atleast_1d(p), dtype=np.float64) for p in params]
        assert len(arrs) == N_PARAMS[d], (dist, len(arrs))
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs] + [None] * (4 - len(arrs)))
        lens = (C.c_size_t * 4)(*[a.size for a in arrs] + [0] * (4 - len(arrs)))
        out = np.full((n_streams, n_samples), np.nan, dtype=np.float64)
        rc = self._fn("sample")(C.c_int(d), C.c_int(1 if real_type == "f32" else 0),
                                state.ctypes.data_as(C.c_void_p), C.c_size_t(n_streams),
                                C.c_int(1 if deterministic else 0),
                                C.c_size_t(n_samples), ptrs, lens,
                                out.ctypes.data_as(C.c_void_p), C.c_int(n_threads))
        if rc != 0:
            raise OracleError(self._err())
        return out

    def generator_draw(self, gen, dist, seed, n_streams, n_samples, params=(), real_type="f64",
                       jumps=0, long_jumps=0):
        """The reference's samplers instantiated on another generator (reference
        build only): prng<gen>(n_streams, seed), jumps, n_samples draws per stream.
        Returns (draws, final states as exported)."""
        d = DIST[dist]
        arrs = [np.ascontiguousarray(np.atleast_1d(p), dtype=np.float64) for p in params]
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs] + [None] * (4 - len(arrs)))
        lens = (C.c_size_t * 4)(*[a.size for a in arrs] + [0] * (4 - len(arrs)))
        fam = gen
        for scr in ("plusplus", "starstar", "plus"):
            if gen.endswith(scr):
                fam = gen[:-len(scr)]
                break
        words, wt = {"xoshiro256": (4, np.uint64), "xoshiro128": (4, np.uint32),
                     "xoroshiro128": (2, np.uint64), "xoshiro512": (8, np.uint64)}[fam]
        out = np.full((n_streams, n_samples), np.nan, dtype=np.float64)
        st = np.zeros((n_streams, words), dtype=wt)
        rc = self._fn("generator_draw")(C.c_int(GEN_NAMES.index(gen)), C.c_int(d), C.c_int(1 if real_type == "f32" else 0),
                                        C.c_uint64(seed), C.c_size_t(n_streams), C.c_size_t(n_samples),
                                        C.c_int(jumps), C.c_int(long_jumps), ptrs, lens,
                                        out.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise OracleError(self._err())
        return out, st

    def generator_seed_vector(self, gen, words, n_streams, reimport_and_jump=False):
        """prng<gen>(n_streams, seed vector) [-> export -> import -> jump] -> export (reference build only)."""
        fam = gen
        for scr in ("plusplus", "starstar", "plus"):
            if gen.endswith(scr):
                fam = gen[:-len(scr)]
                break
        n_words, wt = {"xoshiro256": (4, np.uint64), "xoshiro128": (4, np.uint32),
                       "xoroshiro128": (2, np.uint64), "xoshiro512": (8, np.uint64)}[fam]
        w = np.ascontiguousarray(words, dtype=wt)
        st = np.zeros((n_streams, n_words), dtype=wt)
        rc = self._fn("generator_seed_vector")(C.c_int(GEN_NAMES.index(gen)), w.ctypes.data_as(C.c_void_p),
                                               C.c_size_t(w.size), C.c_size_t(n_streams),
                                               C.c_int(1 if reimport_and_jump else 0), st.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise OracleError(self._err())
        return st

    def sir_filter(self, system_state, filter_state, pars, data_time, data_incidence, resample=True):
        """The SIR particle filter of the reference's test-suite over the
        reference's C++ samplers (oracle/ref_harness.cpp, reference kind only).
        States advanced in place; returns (log_likelihood, final [4][n])."""
        assert self.kind == "reference"
        n = system_state.shape[0]
        t = np.ascontiguousarray(data_time, dtype=np.float64)
        y = np.ascontiguousarray(data_incidence, dtype=np.float64)
        ll = C.c_double(0.0)
        final = np.empty((4, n), dtype=np.float64)
        rc = self.lib.ref_sir_filter(
            system_state.ctypes.data_as(C.c_void_p), C.c_size_t(n),
            filter_state.ctypes.data_as(C.c_void_p), C.c_double(pars["N"]), C.c_double(pars["I0"]),
            C.c_double(pars["beta"]), C.c_double(pars["gamma"]), C.c_double(pars["exp_noise"]),
            t.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_size_t(t.size),
            C.c_int(1 if resample else 0), C.byref(ll), final.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise OracleError(self._err())
        return ll.value, final

    def multinomial(self, state, n_samples, size, prob, deterministic=False):
        n_streams = state.shape[0]
        size = np.ascontiguousarray(np.atleast_1d(size), dtype=np.float64)
        prob = np.asarray(prob, dtype=np.float64)
        if prob.ndim == 1:
            prob_len, prob_cols = prob.size, 1
            pf = np.ascontiguousarray(prob)
        else:  # (prob_len, cols) R matrix -> column-major memory
            prob_len, prob_cols = prob.shape
            pf = np.ascontiguousarray(prob.T)
        out = np.zeros((n_streams, n_samples, prob_len), dtype=np.float64)
        if self.kind == "port":
            rc = self.lib.orc_multinomial(state.ctypes.data_as(C.c_void_p),
                                          C.c_size_t(n_streams),
                                          C.c_int(1 if deterministic else 0),
                                          C.c_size_t(n_samples),
                                          size.ctypes.data_as(C.c_void_p), C.c_size_t(size.size),
                                          pf.ctypes.data_as(C.c_void_p), C.c_size_t(prob_len),
                                          C.c_size_t(prob_cols),
                                          out.ctypes.data_as(C.c_void_p))
        else:
            assert not deterministic
            rc = self.lib.ref_multinomial(state.ctypes.data_as(C.c_void_p),
                                          C.c_size_t(n_streams), C.c_size_t(n_samples),
                                          size.ctypes.data_as(C.c_void_p), C.c_size_t(size.size),
                                          pf.ctypes.data_as(C.c_void_p), C.c_size_t(prob_len),
                                          C.c_size_t(prob_cols),
                                          out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise OracleError(self._err())
        return out

    # -- densities ---------------------------------------------------------
    def density(self, dens, x, params, log=True, real_type="f64", n_threads=1):
        d = DENS[dens] if isinstance(dens, str) else dens
        keep = []

        def prep(a):
            a = np.asarray(a)
            if a.dtype.kind in "iu":
                a = np.ascontiguousarray(a, dtype=np.int32)
                keep.append(a)
                return a.ctypes.data, 1
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return a.ctypes.data, 0

        xp, xi = prep(x)
        n = keep[0].size
        pp = [prep(p) for p in params]
        ptrs = (C.c_void_p * 4)(*[p[0] for p in pp] + [None] * (4 - len(pp)))
        isi = (C.c_int * 4)(*[p[1] for p in pp] + [0] * (4 - len(pp)))
        out = np.zeros(n, dtype=np.float64)
        rc = self._fn("density")(C.c_int(d), C.c_int(1 if real_type == "f32" else 0),
                                 C.c_size_t(n), C.c_void_p(xp), C.c_int(xi), ptrs, isi,
                                 C.c_int(1 if log else 0),
                                 out.ctypes.data_as(C.c_void_p), C.c_int(n_threads))
        if rc != 0:
            raise OracleError(self._err())
        return out


def _load_synthetic():
    """monty_b200/synthetic.py (pure numpy) loaded by path, so that the oracle's
    helpers do not import the product package (and its kernel library)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mb200_synthetic", os.path.join(ROOT, "monty_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synthetic = _load_synthetic()
counter_hash = synthetic.counter_hash
