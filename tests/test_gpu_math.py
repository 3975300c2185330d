"""The samplers' own double-precision math routines evaluated ON THE DEVICE
(through the examples library's probe kernel):

  * m_log / m_log_normal and cos_0_2pi (include/monty_b200/detail/glibc_math.cuh) must return the
    BITS of the host's libm -- glibc 2.39, what the reference's std::log /
    std::cos resolve to (hdr/normal_box_muller.hpp:27-34, hdr/exponential.hpp:26,
    hdr/gamma.hpp:36-50): with that Box-Muller normals, exponentials, gamma and
    beta draws are bit-identical to the reference instead of "a few ulp away".
    1e8 arguments for log, 1e8 for cos (MONTY_B200_GLIBC_N to change), expected
    values from the live libm of THIS host (oracle/libglibc_check.so: gm_libm),
    so a host whose libm took another ifunc variant would be caught here;
  * sqrt_normal must be the correctly rounded square root.
"""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(os.environ.get("MONTY_B200_GLIBC_N", "100000000"))
CHUNK = 10_000_000


@pytest.fixture(scope="module")
def probe():
    from monty_b200 import examples
    return examples.math_probe


@pytest.fixture(scope="module")
def libm():
    path = os.path.join(ROOT, "oracle", "libglibc_check.so")
    if not os.path.exists(path):
        pytest.skip("oracle/libglibc_check.so not built")
    lib = ctypes.CDLL(path)
    lib.gm_libm.restype = None
    lib.gm_libm.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]

    def call(fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        lib.gm_libm({"log": 0, "cos": 1, "exp": 2, "sqrt": 3, "lgamma": 4, "pow": 5, "tan": 6, "log1p": 7, "erfc": 8, "tanf": 9}[fn], x.ctypes.data, y.ctypes.data, x.size)
        return y
    return call


def same_bits(a, b):
    return (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))


def random_reals(rs, n):
    k = rs.integers(0, 1 << 53, n, dtype=np.uint64)
    return k.astype(np.float64) * 2.0 ** -53 + 2.0 ** -54          # random_real() outputs


def test_log_of_uniform_reals_is_glibc_bit_for_bit(probe, libm):
    rs = np.random.default_rng(1)
    bad = 0
    for _ in range(max(1, N // CHUNK)):
        x = random_reals(rs, min(N, CHUNK))
        got = probe("log_normal", x)
        ok = same_bits(got, libm("log", x))
        bad += int((~ok).sum())
        assert ok.all(), "first offender x = %s" % x[~ok][0].hex()
    assert bad == 0


def log_corner_arguments(rs):
    wide = np.exp(rs.uniform(-708, 709, 400_000))
    near_one = 1.0 + rs.uniform(-1, 1, 400_000) * 2.0 ** -rs.integers(3, 52, 400_000).astype(np.float64)
    around = rs.uniform(0.6, 1.5, 400_000)
    # both sides of every row boundary of the 128-row table, two binades, and the
    # ends of the window around 1 (1 - 2^-4, 1 + 0x1.09p-4)
    off = np.uint64(0x3FE6000000000000)
    edges = (off + (np.arange(129, dtype=np.uint64) << np.uint64(45)))[:, None] \
        + np.arange(-8, 8).astype(np.int64).astype(np.uint64)[None, :]
    edges = np.concatenate([edges.ravel(), edges.ravel() + (np.uint64(1) << np.uint64(52))]).view(np.float64)
    win = np.array([0x3FEE000000000000, 0x3FF1090000000000], dtype=np.uint64)[:, None] \
        + np.arange(-8, 8).astype(np.int64).astype(np.uint64)[None, :]
    special = np.array([1.0, np.nextafter(1.0, 0), np.nextafter(1.0, 2), 2.0 ** -1022, np.finfo(np.float64).max,
                        2.0 ** -54, 0.5, 2.0, np.e, 10.0])
    return np.concatenate([wide, near_one, around, edges, win.ravel().view(np.float64), special])


def test_log_on_corner_arguments(probe, libm):
    x = log_corner_arguments(np.random.default_rng(4))
    expect = libm("log", x)
    assert same_bits(probe("log", x), expect).all()
    assert same_bits(probe("log_normal", x), expect).all()
    assert probe("log", np.array([1.0]))[0] == 0.0


def test_log_special_arguments(probe, libm):
    x = np.array([0.0, -0.0, -1.0, np.inf, np.nan, 5e-324, 2.0 ** -1040, np.nextafter(2.0 ** -1022, 0)])
    got = probe("log", x)
    assert got[0] == -np.inf and got[1] == -np.inf and np.isnan(got[2]) and got[3] == np.inf and np.isnan(got[4])
    assert same_bits(got[5:], libm("log", x[5:])).all()          # subnormals


def test_cos_of_box_muller_angles_is_glibc_bit_for_bit(probe, libm):
    rs = np.random.default_rng(2)
    two_pi = 2 * np.pi
    for _ in range(max(1, N // CHUNK)):
        x = two_pi * random_reals(rs, min(N, CHUNK))          # rounded like the sampler's two_pi * u2
        ok = same_bits(probe("cos_0_2pi", x), libm("cos", x))
        assert ok.all(), "first offender x = %s" % x[~ok][0].hex()


def test_cos_on_corner_angles(probe, libm):
    two_pi = 2 * np.pi
    # the grid angles nearest every multiple of pi / 4 (where the reduction cancels
    # most), the branch boundaries of s_sin.c (0.855469, 2.426265), both sides of
    # TAYLOR_SIN's 0.126 after reduction, tiny angles
    near = []
    for c in np.arange(0, 9) / 8.0:
        kk = np.int64(c * 2.0 ** 53) + np.arange(-4000, 4000)
        kk = kk[(kk >= 0) & (kk < (1 << 53))]
        near.append(two_pi * (kk.astype(np.float64) * 2.0 ** -53 + 2.0 ** -54))
    cuts = []
    for c in (0.85546875, 2.426265, 0.126, np.pi / 2 - 0.126, np.pi / 2 + 0.126, np.pi - 0.126, np.pi + 0.126,
              3 * np.pi / 2 - 0.126, 3 * np.pi / 2 + 0.126, two_pi - 0.126, 2.0 ** -27):
        b = np.float64(c).view(np.uint64)
        cuts.append((b + np.arange(-2000, 2000).astype(np.int64).astype(np.uint64)).view(np.float64))
    x = np.concatenate(near + cuts + [two_pi * np.array([1.0, 2.0 ** -54, 0.25, 0.5, 0.75]), np.array([0.0, 1e-300, 2.0 ** -30])])
    x = x[(x >= 0) & (x <= two_pi)]
    assert same_bits(probe("cos_0_2pi", x), libm("cos", x)).all()
    assert probe("cos_0_2pi", np.array([0.0]))[0] == 1.0


def test_lgamma_is_glibc_bit_for_bit(probe, libm):
    """m_lgamma (glibc's e_lgamma_r.c restated, detail/glibc_math.cuh) on every
    piece of the function: (0, 2) with its three sub-intervals and the 2^-70
    cut, [2, 8), Stirling above 8, the x (log x - 1) range, integers (the table)
    and half-integers; zero / negative / non-finite arguments take the toolkit's
    routine and only have to agree in value."""
    rs = np.random.default_rng(6)
    n = max(1, N // 10)
    x = np.concatenate([rs.uniform(0, 10, n), rs.uniform(0.5, 300.5, n), np.exp(rs.uniform(-200, 700, n)),
                        np.arange(1, 20000, dtype=np.float64), np.arange(1, 2000) + 0.5,
                        np.array([2.0 ** -71, 2.0 ** -70, 0.23, 0.2316, 0.7316, 0.9, 1.0, 1.2316, 1.4616321449683622,
                                  1.7316, 2.0, 3.0, 7.999999, 8.0, 2.0 ** 58, 1e300])])
    for b in (0.23163998, 0.73163998, 0.9, 1.23164, 1.73163998, 2.0, 8.0):          # both sides of every cut
        v = np.float64(b).view(np.uint64)
        x = np.concatenate([x, (v + np.arange(-3000, 3000).astype(np.int64).astype(np.uint64)).view(np.float64)])
    ok = same_bits(probe("lgamma", x), libm("lgamma", x))
    assert ok.all(), "first offender x = %s" % x[~ok][0].hex()
    odd = np.array([0.0, -0.5, -2.5, np.inf])
    got = probe("lgamma", odd)
    assert got[0] == np.inf and got[3] == np.inf
    assert np.allclose(got[1:3], libm("lgamma", odd[1:3]), rtol=1e-14)


def test_exp_is_glibc_bit_for_bit(probe, libm):
    """m_exp (glibc's e_exp.c restated, the scaled evaluation next to over- and
    underflow included): the same bits as the host's libm for every argument."""
    rs = np.random.default_rng(7)
    n = max(1, N // 10)
    x = np.concatenate([rs.uniform(-512, 512, n), rs.uniform(-40, 3, n), rs.uniform(-1, 1, n) * 2.0 ** -rs.integers(0, 70, n).astype(np.float64),
                        np.array([0.0, -0.0, 1.0, -1.0, 2.0 ** -54, 2.0 ** -55, -2.0 ** -60, 511.999, -511.999, 709.0, -708.0])])
    far = np.array([512.0, 600.0, 709.7, 709.79, -600.0, -708.3, -740.0, -745.2, -750.0, 720.0, 1024.0, -1024.0, 1e300, -1e300,
                    np.inf, -np.inf, np.nan])
    x = np.concatenate([x, far, rs.uniform(-1030, 1030, n), -rs.uniform(700, 750, n)])
    got, want = probe("exp", x), libm("exp", x)
    ok = same_bits(got, want)
    assert ok.all(), "first offender x = %s" % x[~ok][0].hex()


def test_pow_is_glibc_bit_for_bit(probe, libm):
    """m_pow (glibc's e_pow.c restated): pow(u, 1 / shape) as the gamma and Weibull
    samplers call it, general positive bases, and special operands (which take the
    toolkit's pow and must agree in value)."""
    rs = np.random.default_rng(8)
    n = max(2, (N // 20) * 2)
    x = np.empty(n)
    x[0::2] = random_reals(rs, n // 2)                   # pow(x[i], x[i ^ 1]): even = base for odd's exponent and vice versa
    x[1::2] = rs.uniform(0.05, 12, n // 2)
    got, want = probe("pow", x), libm("pow", x)
    assert same_bits(got, want).all(), "first offender %s" % x[~same_bits(got, want)][:2]
    y = np.empty(n)
    y[0::2] = np.exp(rs.uniform(-300, 300, n // 2))
    y[1::2] = rs.uniform(0.1, 2.2, n // 2)
    got, want = probe("pow", y), libm("pow", y)
    # results through over- and underflow (specialcase() of e_pow.c) included
    assert same_bits(got, want).all(), "first offender %s" % y[~same_bits(got, want)][:2]
    # negative bases with integer exponents (the cauchy density squares (x - location) / scale
    # through pow): the library's |x| path with the sign restored
    z = np.empty(n)
    z[0::2] = rs.uniform(-50, 50, n // 2)
    z[1::2] = rs.integers(1, 4, n // 2).astype(np.float64)
    got, want = probe("pow", z), libm("pow", z)
    ok = same_bits(got, want) | (np.arange(n) % 2 == 1)          # odd slots: pow(integer, a real), not of interest
    assert ok.all(), "first offender %s" % z[~ok][:2]
    odd = np.array([0.0, 2.5, 2.0, 0.0, -2.0, 3.0, 1.0, 1e300, np.inf, 0.5, 4.0, 0.5])
    g, w = probe("pow", odd), libm("pow", odd)
    with np.errstate(invalid="ignore"):
        assert (same_bits(g, w) | (np.abs(g - w) <= np.abs(np.spacing(w)))).all()


def test_tan_is_glibc_bit_for_bit(probe, libm):
    """m_tan (glibc's s_tan.c restated): tan(pi u) as the cauchy sampler and the Poisson
    Cauchy-envelope regime call it, both reductions (|x| <= 25 and <= 1e8), next to the
    poles; beyond 1e8 the toolkit's tan answers.  The float overload (s_tanf.c /
    k_tanf.c) on the float cauchy argument and below 120."""
    rs = np.random.default_rng(9)
    n = max(1, N // 10)
    k = rs.integers(1, 64, n).astype(np.float64)
    x = np.concatenate([np.pi * random_reals(rs, n), rs.uniform(-30, 30, n), rs.uniform(-1, 1, n) * 10.0 ** rs.uniform(-9, 8, n),
                        k * (np.pi / 2) * (1 + rs.uniform(-1e-7, 1e-7, n)), np.array([0.0, -0.0, 1e-9, 0.0608, 0.787, 25.0, 1e8, np.pi / 2, np.pi])])
    got, want = probe("tan", x), libm("tan", x)
    ok = same_bits(got, want)
    assert ok.all(), "first offender x = %s" % x[~ok][0].hex()
    far = np.array([1.5e8, 1e12, -3e15, 1e300])
    g, w = probe("tan", far), libm("tan", far)
    assert (np.abs(g - w) <= 2 * np.abs(np.spacing(w))).all()
    xf = np.concatenate([(np.float32(np.pi) * random_reals(rs, n).astype(np.float32)).astype(np.float64),
                         rs.uniform(-119, 119, n).astype(np.float32).astype(np.float64),
                         (rs.uniform(-1, 1, n) * 2.0 ** -rs.integers(0, 30, n)).astype(np.float32).astype(np.float64)])
    got, want = probe("tanf", xf), libm("tanf", xf)
    ok = same_bits(got, want)
    assert ok.all(), "first offender x = %s" % xf[~ok][0].hex()


def test_log1p_and_erfc_are_glibc_bit_for_bit(probe, libm):
    """m_log1p (s_log1p.c) and m_erfc (s_erf.c): the cauchy and truncated-normal densities'
    last two libm calls; erfc's two exponentials are e_exp.c (m_exp)."""
    rs = np.random.default_rng(10)
    n = max(1, N // 10)
    x = np.concatenate([rs.uniform(-1, 2, n), np.exp(rs.uniform(-60, 60, n)), -np.exp(rs.uniform(-60, 0, n)) * 0.999,
                        np.array([0.0, -0.0, 2.0 ** -30, -2.0 ** -30, 2.0 ** -55, 0.41421, -0.2929, 1e300])])
    got, want = probe("log1p", x), libm("log1p", x)
    ok = same_bits(got, want)
    assert ok.all(), "first offender x = %s" % x[~ok][0].hex()
    x = np.concatenate([rs.uniform(-6.5, 28.5, n), rs.uniform(-1.3, 1.3, n), rs.uniform(-1, 1, n) * 2.0 ** -rs.integers(0, 60, n).astype(np.float64),
                        np.array([0.0, -0.0, 0.84375, 1.25, 1 / 0.35, -6.0, 6.0, 22.0])])
    got, want = probe("erfc", x), libm("erfc", x)
    ok = same_bits(got, want)
    assert ok.all(), "first offender x = %s" % x[~ok][0].hex()
    far = np.array([22.7, 25.0, 26.5, 27.9, 28.0, 40.0, -7.0, -30.0, np.inf, -np.inf, np.nan])
    assert same_bits(probe("erfc", far), libm("erfc", far)).all()


def test_sqrt_is_correctly_rounded(probe):
    rs = np.random.default_rng(3)
    a = np.concatenate([np.exp(rs.uniform(-650, 650, 300_000)),
                        -2 * np.log(rs.uniform(2.3e-16, 1.0, 300_000)),        # Box-Muller radicands
                        np.array([1.1102230246251565e-16, 73.0, 1.0, 2.0, 4.0, np.nextafter(4.0, 0), 0.25])])
    a = a[a > 0]
    got = probe("sqrt_normal", a)
    assert np.array_equal(got, np.sqrt(a))
    assert np.array_equal(got, probe("sqrt", a))
