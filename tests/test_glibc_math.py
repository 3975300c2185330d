"""CPU check of the product's glibc-exact log / cos (include/monty_b200/detail/glibc_math.cuh).

The header compiles as plain host C++ (oracle/libglibc_check.so, g++ -mfma, no
implicit contraction), so the SOURCE that runs on the device is compared here,
bit for bit, with the live libm -- what the reference's std::log / std::cos
call (hdr/normal_box_muller.hpp:27-34, hdr/exponential.hpp:26, hdr/gamma.hpp:
36-50).  1e8 arguments per class in the default run of this file would take
~10 s on 8 cores; 2e7 per class keeps the CPU suite short (the full 1e8 run:
MONTY_B200_GLIBC_N=100000000).  tests/test_gpu_math.py repeats the comparison
with the routines evaluated on the device.
"""
import ctypes
import math
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libglibc_check.so")
N = int(os.environ.get("MONTY_B200_GLIBC_N", "20000000"))

KINDS = {
    9: "lgamma on (0, 10): every piece below 8",
    10: "lgamma of any positive normal double",
    11: "lgamma on (0.5, 300.5): the densities' range",
    12: "lgamma of integers and half-integers",
    13: "exp on [-700, 700]",
    14: "exp of every magnitude below 2",
    15: "pow(uniform, (0.05, 12)): the gamma / Weibull samplers' calls",
    16: "pow(any positive normal, (-20, 20))",
    17: "pow((0, 50), any bit pattern)",
    33: "pow((-100, 100), integers in [-20, 20])",
    34: "squares and cubes of any bit pattern",
    18: "tan(pi u): the cauchy sampler's argument",
    19: "tan on [-30, 30]: both reductions",
    20: "tan of every magnitude below 1.3e8",
    21: "tan next to multiples of pi / 2",
    22: "log1p on (-1, 2)",
    23: "log1p of any bit pattern",
    24: "log1p of any positive normal double",
    25: "log1p of magnitudes 2^-60 .. 8, both signs",
    26: "erfc on [-6, 6]: every polynomial piece",
    27: "erfc on [-30, 30]: out to the underflow",
    28: "erfc of any bit pattern",
    29: "erfc of magnitudes 2^-60 .. 32, both signs",
    30: "exp on [-1100, 1100]: through over- and underflow",
    31: "exp of any bit pattern",
    32: "exp with subnormal results",
    6: "cos (lock-step formulation) of the Box-Muller angle",
    7: "cos (lock-step formulation) on [-1000, 1000]",
    8: "cos (lock-step formulation), magnitudes to 1e8",
    0: "log of random_real() outputs",
    1: "log of random positive normal doubles",
    2: "log on [0.9, 1.1] (the window around 1 and its edges)",
    3: "cos of the Box-Muller angle 2 pi u",
    4: "cos on [-8, 8]",
    5: "log of random bit patterns (negative, NaN, inf, subnormal)",
}


@pytest.fixture(scope="module")
def gm():
    if not os.path.exists(LIB):
        pytest.skip("oracle/libglibc_check.so not built (make -C oracle)")
    lib = ctypes.CDLL(LIB)
    lib.gm_compare.restype = ctypes.c_int64
    lib.gm_compare.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, ctypes.POINTER(ctypes.c_double)]
    for f in (lib.gm_log, lib.gm_cos):
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double]
    lib.gm_compare_float.restype = ctypes.c_int64
    lib.gm_compare_float.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]
    lib.gm_compare_powf.restype = ctypes.c_int64
    lib.gm_compare_powf.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int]
    lib.gm_compare_log_batch.restype = ctypes.c_int64
    lib.gm_compare_log_batch.argtypes = [ctypes.c_uint64, ctypes.c_int64]
    lib.gm_compare_lgamma_batch.restype = ctypes.c_int64
    lib.gm_compare_lgamma_batch.argtypes = [ctypes.c_uint64, ctypes.c_int64]
    return lib


@pytest.mark.parametrize("kind", sorted(KINDS))
def test_same_bits_as_libm(gm, kind):
    bad = ctypes.c_double(0)
    mism = gm.gm_compare(kind, 20240607 + kind, N, ctypes.byref(bad))
    assert mism == 0, "%s: %d of %d differ, first at x = %s" % (KINDS[kind], mism, N, bad.value.hex())


# every 16th float bit pattern by default (the whole set takes ~40 s on 8 cores:
# MONTY_B200_GLIBC_FLOAT_STRIDE=1); the committed result of the exhaustive run is 0 differences for each
FLOAT_STRIDE = int(os.environ.get("MONTY_B200_GLIBC_FLOAT_STRIDE", "16"))


@pytest.mark.parametrize("fn,name,lo,hi", [(0, "logf", 0, 0xFFFFFFFF), (1, "expf", 0, 0xFFFFFFFF),
                                          (2, "cosf", 0, 0xFF800000), (3, "lgammaf", 1, 0x7F800000),
                                          (4, "tanf", 0, 0xFF800000), (5, "log1pf", 0, 0xFF800001)])
def test_float_routines_same_bits_as_libm(gm, fn, name, lo, hi):
    """real_type = float: logf / expf / cosf / lgammaf / tanf / log1pf over the float bit patterns
    (blocks of 2^20 patterns, every FLOAT_STRIDE-th block)."""
    bad = ctypes.c_uint32(0)
    block = 1 << 20
    total = 0
    for k, start in enumerate(range(lo, hi, block)):
        if k % FLOAT_STRIDE:
            continue
        total += gm.gm_compare_float(fn, start, min(hi, start + block), ctypes.byref(bad))
    assert total == 0, "%s: %d differ, first bits 0x%08x" % (name, total, bad.value)


def test_powf_and_the_batched_forms(gm):
    assert gm.gm_compare_powf(5, N, 0) == 0          # powf(uniform float, (0.05, 20))
    assert gm.gm_compare_powf(6, N, 1) == 0          # powf(any positive float, any float)
    assert gm.gm_compare_powf(8, N, 2) == 0          # powf(any float, small integers / any float): negative bases, subnormals
    assert gm.gm_compare_powf(9, N, 3) == 0          # results next to over- and underflow
    assert gm.gm_compare_log_batch(7, N // 4) == 0   # log_positive_normal_batch<4> == four calls
    assert gm.gm_compare_lgamma_batch(3, N // 6) == 0


def test_the_comparison_can_fail(gm):
    """Negative control: the previous round's accurate-but-different log (<= 1 ulp)
    disagrees with glibc on a few percent of arguments; here the same is shown
    with the correctly rounded value from Python's math.fsum-free long double."""
    rs = np.random.default_rng(5)
    x = rs.uniform(0.0, 1.0, 20000)
    ours = np.array([gm.gm_log(float(v)) for v in x])
    libm = np.array([math.log(float(v)) for v in x])
    assert np.array_equal(ours, libm)
    if np.finfo(np.longdouble).nmant >= 63:
        exact = np.log(x.astype(np.longdouble)).astype(np.float64)
        assert (exact != libm).any(), "glibc's log would be correctly rounded -- it is not"


def test_special_values(gm):
    assert gm.gm_log(1.0) == 0.0 and math.copysign(1.0, gm.gm_log(1.0)) == 1.0
    assert gm.gm_log(0.0) == -math.inf and gm.gm_log(-0.0) == -math.inf
    assert gm.gm_log(math.inf) == math.inf
    assert math.isnan(gm.gm_log(-1.0)) and math.isnan(gm.gm_log(math.nan))
    assert gm.gm_log(5e-324) == math.log(5e-324)
    assert gm.gm_cos(0.0) == 1.0 and gm.gm_cos(2.0 ** -30) == 1.0
    two_pi = 2 * math.pi
    for u in (1.0, 2.0 ** -54, 0.25, 0.5, 0.75, 1.0 - 2.0 ** -53):
        assert gm.gm_cos(two_pi * u) == math.cos(two_pi * u)


def test_tables_have_their_mathematical_meaning():
    """glibc_tables.inc is data read from libm.so.6 (tools/extract_glibc_tables.py);
    check what the numbers ARE: logc = log(1 / invc) rounded to 2^-43, the
    sin / cos rows are double-doubles of sin / cos(k / 128) to 2^-100."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    txt = open(os.path.join(ROOT, "include", "monty_b200", "detail", "glibc_tables.inc")).read()

    def macro(name):
        body = txt.split("#define %s " % name)[1]
        body = body.split("\n#define")[0].split("\n/*")[0].replace("\\\n", " ")
        return [float.fromhex(t) for t in body.replace("{", " ").replace("}", " ").replace(",", " ").split()]

    tab = macro("GLIBC_LOG_TAB")
    assert len(tab) == 256
    for k in range(128):
        invc, logc = tab[2 * k], tab[2 * k + 1]
        c = 1 / mp.mpf(invc)
        assert abs(mp.log(c) - logc) < mp.mpf(2) ** -42, k
        assert struct.unpack("<Q", struct.pack("<d", logc))[0] & 0x1FF == 0 or logc == 0.0 or abs(logc) < 2 ** -8
    xfg = macro("GLIBC_TAN_XFG")
    assert len(xfg) == 186 * 4
    for k in range(186):
        xi, fi, gi = xfg[4 * k:4 * k + 3]
        assert abs(xi - (k + 16) / 256) < 2 ** -24, k
        assert abs(mp.tan(mp.mpf(xi)) - fi) < mp.mpf(2) ** -52 * fi and abs(mp.cot(mp.mpf(xi)) - gi) < mp.mpf(2) ** -52 * gi, k
    sc = macro("GLIBC_SINCOS_TAB")
    assert len(sc) == 440
    for k in range(110):
        sn, ssn, cs, ccs = sc[4 * k:4 * k + 4]
        assert abs(mp.sin(mp.mpf(k) / 128) - (mp.mpf(sn) + ssn)) < mp.mpf(2) ** -100, k
        assert abs(mp.cos(mp.mpf(k) / 128) - (mp.mpf(cs) + ccs)) < mp.mpf(2) ** -100, k
