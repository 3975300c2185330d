"""Accuracy of the samplers' own double-precision math routines
(csrc/samplers.cuh: fast_log / m_log_normal, cos_0_2pi, sqrt_normal), evaluated
on the device through the examples library and compared with 80-bit long
double results of the host libm (64-bit significand: the comparison resolves
0.001 ulp of a double).

The reference calls std::log / std::cos / std::sqrt (glibc: < 1 ulp, sqrt
exact); the continuous samplers are held to 4 ulp of the reference's draws
(test_gpu_samplers.py), which these bounds leave room for."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LD = np.longdouble


@pytest.fixture(scope="module")
def probe():
    if np.finfo(LD).nmant < 63:
        pytest.skip("long double is not wider than double on this host")
    from monty_b200 import examples
    return examples.math_probe


def ulp_error(got, exact_ld):
    exact = exact_ld.astype(np.float64)
    ulp = np.abs(np.spacing(exact)).astype(LD)
    return np.abs(got.astype(LD) - exact_ld) / ulp


def log_arguments(rs):
    k = rs.integers(0, 1 << 53, 200_000, dtype=np.uint64)
    uniform = k.astype(np.float64) * 2.0 ** -53 + 2.0 ** -54          # random_real() outputs
    wide = np.exp(rs.uniform(-708, 709, 100_000))
    near_one = 1.0 + rs.uniform(-1, 1, 100_000) * 2.0 ** -rs.integers(5, 52, 100_000).astype(np.float64)
    around = rs.uniform(0.6, 1.5, 100_000)
    # both sides of every row boundary of the table, in the binades next to 1
    off = np.uint64(0x3FE5C00000000000)
    edges = (off + (np.arange(33, dtype=np.uint64) << np.uint64(47)))[:, None] + np.arange(-8, 8).astype(np.int64).astype(np.uint64)[None, :]
    edges = np.concatenate([edges.ravel(), edges.ravel() + (np.uint64(1) << np.uint64(52))]).view(np.float64)
    special = np.array([1.0, np.nextafter(1.0, 0), np.nextafter(1.0, 2), 2.0 ** -1022, np.finfo(np.float64).max, 2.0 ** -54,
                        0.5, 2.0, np.e, 10.0])
    return np.concatenate([uniform, wide, near_one, around, edges, special])


def test_log_within_one_ulp(probe):
    x = log_arguments(np.random.default_rng(1))
    got = probe("fast_log", x)
    err = ulp_error(got, np.log(x.astype(LD)))
    print("fast_log: max error %.4f ulp at x = %r" % (err.max(), x[err.argmax()]))
    assert err.max() <= 1.01
    assert probe("fast_log", np.array([1.0]))[0] == 0.0
    # the unchecked variant is the same arithmetic
    assert np.array_equal(probe("log_normal", x), got)


def test_log_special_arguments_take_the_toolkit_path(probe):
    x = np.array([0.0, -0.0, -1.0, np.inf, np.nan, 5e-324, 2.0 ** -1040, np.nextafter(2.0 ** -1022, 0)])
    got = probe("fast_log", x)
    with np.errstate(divide="ignore", invalid="ignore"):
        expect = np.log(x)
    assert got[0] == -np.inf and got[1] == -np.inf and np.isnan(got[2]) and got[3] == np.inf and np.isnan(got[4])
    assert (np.abs(got[5:] - expect[5:]) <= 2 * np.abs(np.spacing(expect[5:]))).all()


def test_cos_on_box_muller_angles(probe):
    rs = np.random.default_rng(2)
    two_pi = 2 * np.pi
    k = rs.integers(0, 1 << 53, 400_000, dtype=np.uint64)
    u = k.astype(np.float64) * 2.0 ** -53 + 2.0 ** -54
    # the grid angles nearest every multiple of pi / 4, where the reduction cancels most
    near = []
    for c in np.arange(0, 9) / 8.0:
        kk = np.int64(c * 2.0 ** 53) + np.arange(-400, 400)
        kk = kk[(kk >= 0) & (kk < (1 << 53))]
        near.append(kk.astype(np.float64) * 2.0 ** -53 + 2.0 ** -54)
    u = np.concatenate([u] + near + [np.array([1.0, 2.0 ** -54, 0.25, 0.5, 0.75])])
    x = two_pi * u                                  # rounded like the sampler's two_pi * u2
    assert x.min() > 0 and x.max() <= two_pi
    got = probe("cos_0_2pi", x)
    err = ulp_error(got, np.cos(x.astype(LD)))
    print("cos_0_2pi: max error %.4f ulp at x = %r" % (err.max(), x[err.argmax()]))
    assert err.max() <= 1.5
    assert probe("cos_0_2pi", np.array([0.0]))[0] == 1.0


def test_sqrt_is_correctly_rounded(probe):
    rs = np.random.default_rng(3)
    a = np.concatenate([np.exp(rs.uniform(-650, 650, 300_000)),
                        -2 * np.log(rs.uniform(2.3e-16, 1.0, 300_000)),        # Box-Muller radicands
                        np.array([1.1102230246251565e-16, 73.0, 1.0, 2.0, 4.0, np.nextafter(4.0, 0), 0.25])])
    a = a[a > 0]
    got = probe("sqrt_normal", a)
    assert np.array_equal(got, np.sqrt(a))
    assert np.array_equal(got, probe("sqrt", a))
