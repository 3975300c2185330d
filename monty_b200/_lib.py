"""ctypes binding of the C-ABI (include/monty_b200.h).

The shared library holds every kernel of the product.  If it is missing the
import fails loudly: there is no Python or CPU fallback for any operation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmonty_b200.so")

MB200_OK = 0
MB200_ERR_CUDA = -1
MB200_ERR_ARG = -2
MB200_ERR_PARAM = -3
MB200_ERR_STATE = -4
MB200_ERR_UNSUPPORTED = -5

REAL_F64 = 0
REAL_F32 = 1

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
DIST_N_PARAMS = [0, 2, 1, 1, 2, 2, 2, 2, 1, 2, 2, 2, 2, 2, 3, 4, 2, 2, 2, 3, 3, 2, 3, 3]

DENS = {
    "binomial": 0, "normal": 1, "negative_binomial_mu": 2,
    "negative_binomial_prob": 3, "beta_binomial_prob": 4,
    "beta_binomial_ab": 5, "poisson": 6, "uniform": 7, "exponential_rate": 8,
    "exponential_mean": 9, "gamma_rate": 10, "gamma_scale": 11, "beta": 12,
    "hypergeometric": 13, "log_normal": 14, "truncated_normal": 15,
    "cauchy": 16, "weibull": 17, "zi_poisson": 18,
    "zi_negative_binomial_mu": 19, "zi_negative_binomial_prob": 20,
}
DENS_N_PARAMS = [2, 2, 2, 2, 3, 3, 1, 2, 1, 1, 2, 2, 2, 3, 2, 4, 2, 2, 2, 3, 3]

GENERATORS = [
    "xoshiro256plus", "xoshiro256plusplus", "xoshiro256starstar",
    "xoshiro128plus", "xoshiro128plusplus", "xoshiro128starstar",
    "xoroshiro128plus", "xoroshiro128plusplus", "xoroshiro128starstar",
    "xoshiro512plus", "xoshiro512plusplus", "xoshiro512starstar",
]


class MontyError(RuntimeError):
    """Raised with the reference's error text (what R would show via cpp11::stop)."""

    def __init__(self, message, code=None):
        super().__init__(message)
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "monty_b200: %s is missing. Build the CUDA library with "
            "`make -C monty_b200/csrc` (or __graft_entry__.build()); there is "
            "no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
    sig = {
        "mb200_version": (C.c_char_p, []),
        "mb200_device_count": (i32, []),
        "mb200_last_error": (C.c_char_p, []),
        "mb200_launch_count": (u64, []),
        "mb200_rng_create": (i32, [u64, vp, sz, sz, i32, u64, u64, i32, C.POINTER(vp)]),
        "mb200_rng_free": (None, [vp]),
        "mb200_rng_n_streams": (sz, [vp]),
        "mb200_rng_deterministic": (i32, [vp]),
        "mb200_rng_device": (i32, [vp]),
        "mb200_rng_last_error": (C.c_char_p, [vp]),
        "mb200_rng_state": (i32, [vp, vp, sz]),
        "mb200_rng_set_state": (i32, [vp, vp, sz]),
        "mb200_rng_jump": (i32, [vp, i32]),
        "mb200_rng_long_jump": (i32, [vp, i32]),
        "mb200_rng_state_dev": (vp, [vp]),
        "mb200_rng_cuda_stream": (vp, [vp]),
        "mb200_rng_sync": (i32, [vp]),
        "mb200_rng_sample": (i32, [vp, i32, i32, sz, vp, vp, vp]),
        "mb200_rng_sample_dev": (i32, [vp, i32, i32, sz, vp, vp, vp, i32]),
        "mb200_rng_multinomial": (i32, [vp, sz, vp, sz, vp, sz, sz, vp]),
        "mb200_rng_raw": (i32, [vp, i32, sz, vp]),
        "mb200_xoshiro_run": (i32, [i32, u64, i32, i32, vp, vp, C.POINTER(i32)]),
        "mb200_density": (i32, [i32, i32, sz, vp, i32, vp, vp, i32, vp, vp, i32]),
        "mb200_density_dev": (i32, [i32, i32, sz, vp, i32, vp, vp, i32, vp, vp, i32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error():
    return lib.mb200_last_error().decode()


def check(rc):
    if rc != MB200_OK:
        raise MontyError(last_error(), rc)
