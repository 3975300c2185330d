"""GPU parity, sampler level, against the reference's CPU implementation on
identical seeds (the compiled reference when oracle/_ref is present, else the
C restatement pinned by test_oracle.py).

Bars (BASELINE.json north_star):
  * uniform reals, final generator states: bit-exact;
  * integer-valued draws (double): bit-exact up to a mismatch fraction <= 1e-6
    attributable to libm ulp differences -- we assert 0 mismatches here;
  * continuous draws (double): within 4 ulp.  The device's log, exp, pow, cos and
    tan return glibc's bits (include/monty_b200/detail/glibc_math.cuh), so all 12 continuous
    samplers -- exponential, the three normals, truncated normal, log_normal,
    gamma, beta, weibull, cauchy -- are asserted BIT-IDENTICAL, shifted and
    scaled parameters included;
  * float real_type: bit-identical too (the north-star asks for 1e-6 relative):
    glibc's logf / expf / powf / cosf / lgammaf are restated bit for bit, which
    also makes the integer-valued float draws that subtract large lgammaf terms
    (Poisson lambda >~ 1e3, hypergeometric H2PE, negative binomial) exact --
    they depend on the last bit of the platform's float libm in the REFERENCE
    ITSELF; tanf (s_tanf.c / k_tanf.c) makes the float cauchy draws exact too.
"""
import numpy as np
import pytest

from cases import INTEGER_VALUED, sampler_cases, standard_cases
from gpu_util import gpu_sample, states_of, ulp_distance

pytestmark = pytest.mark.gpu

ALL = sorted(sampler_cases(4))
CONTINUOUS = [n for n in ALL if n not in INTEGER_VALUED and n not in ("real", "uniform")]
# log, exp, pow, cos, sqrt and IEEE arithmetic only: the same bits as the reference
BIT_EXACT = {"exponential_rate", "exponential_mean", "normal_box_muller", "normal_polar", "normal_ziggurat",
             "truncated_normal", "log_normal", "gamma_scale", "gamma_rate", "beta", "weibull", "cauchy"}
EPS = np.finfo(np.float64).eps


def assert_continuous_parity(name, y, exp, pars, n):
    """The north-star bar for continuous double draws (module docstring)."""
    if name in BIT_EXACT:
        bad = y != exp
        assert not bad.any(), (name, int(bad.sum()), y[bad][:3], exp[bad][:3])
        return
    d = ulp_distance(y, exp)
    assert d.max() <= 4, (name, d.max())


@pytest.fixture(scope="module")
def mb():
    import monty_b200
    return monty_b200


@pytest.mark.parametrize("name", ALL)
def test_double_vs_reference_fixture(mb, golden, name):
    """The committed vectors of the unmodified reference (64 streams x 5)."""
    pars = sampler_cases(64)[name]
    y, st, _ = gpu_sample(mb, name, 64, 5, pars, seed=20240607)
    exp = golden["sample/f64/%s/sto" % name]
    assert np.array_equal(st, golden["sample/f64/%s/sto/state" % name])
    if name in INTEGER_VALUED or name in ("real", "uniform"):
        assert np.array_equal(y, exp)
    else:
        assert_continuous_parity(name, y, exp, pars, 64)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("shape", [(4096, 37), (1000, 1), (333, 64), (97, 33)])
def test_double_parity(mb, oracle, name, shape):
    """Mixed regimes, ragged shapes: n_streams not a multiple of 32, odd
    n_samples (unaligned rows -> scalar store path), n_samples = 1."""
    n, ns = shape
    pars = sampler_cases(n)[name]
    y, st, _ = gpu_sample(mb, name, n, ns, pars)
    ost = oracle.seed_states(7, n)
    exp = oracle.sample(name, ost, ns, pars)
    assert np.array_equal(st, ost), "final generator states differ"
    if name in INTEGER_VALUED or name in ("real", "uniform"):
        mism = np.mean(y != exp)
        assert mism == 0.0, "integer-valued mismatch fraction %g" % mism
    else:
        assert_continuous_parity(name, y, exp, pars, n)


@pytest.mark.parametrize("name", ["gamma_scale", "beta", "truncated_normal", "hypergeometric", "normal_box_muller",
                                  "exponential_rate"])
def test_cheap_decisions_never_change_a_stream(mb, oracle, name):
    """Several accept tests are decided in single precision outside a band
    (truncated normal, Marsaglia-Tsang log test), quotients and libm calls have
    range-specific forms (hypergeometric squeeze, Box-Muller, exponential).  A
    decision that differed from the reference's would change how many words the
    stream consumes: 1.6e7 draws per sampler, every final state equal."""
    n, ns = 250_000, 64
    pars = sampler_cases(n)[name]
    y, st, _ = gpu_sample(mb, name, n, ns, pars, seed=99)
    ost = oracle.seed_states(99, n)
    exp = oracle.sample(name, ost, ns, pars)
    assert np.array_equal(st, ost), "%d streams ended in a different state" % int((st != ost).any(axis=1).sum())
    if name in INTEGER_VALUED:
        assert np.array_equal(y, exp)
    else:
        assert_continuous_parity(name, y, exp, pars, n)


SORTED = ["binomial", "poisson", "gamma_scale", "gamma_rate", "beta", "negative_binomial_prob",
          "negative_binomial_mu", "hypergeometric", "truncated_normal"]


@pytest.mark.parametrize("name", SORTED)
@pytest.mark.parametrize("shape", [(5000, 1), (4096, 16), (70001, 7), (2048, 2)])
def test_regime_sorted_kernel_parity(mb, oracle, name, shape):
    """Per-stream parameters, n_samples <= 16, >= 2048 streams: the regime-
    sorting kernel (csrc/sorted_kernel.cuh).  Same bars as test_double_parity;
    shapes cover a ragged last tile, n_samples = 1 and odd n_samples."""
    n, ns = shape
    pars = sampler_cases(n)[name]
    y, st, _ = gpu_sample(mb, name, n, ns, pars)
    ost = oracle.seed_states(7, n)
    exp = oracle.sample(name, ost, ns, pars)
    assert np.array_equal(st, ost), "final generator states differ"
    if name in INTEGER_VALUED:
        assert np.mean(y != exp) == 0.0
    else:
        assert_continuous_parity(name, y, exp, pars, n)


@pytest.mark.parametrize("name", sorted(standard_cases(4)))
def test_double_continuous_within_4_ulp(mb, oracle, name):
    n, ns = 2048, 64
    pars = standard_cases(n)[name]
    y, st, _ = gpu_sample(mb, name, n, ns, pars)
    ost = oracle.seed_states(7, n)
    exp = oracle.sample(name, ost, ns, pars)
    assert np.array_equal(st, ost)
    d = ulp_distance(y, exp)
    assert d.max() <= 4, (name, d.max())           # the north-star's bar, every sampler
    if name in BIT_EXACT or name in ("gamma_scale", "beta"):
        assert d.max() == 0, (name, d.max())       # shapes >= 2 here: no pow


@pytest.mark.parametrize("name", ALL)
def test_float_real_type_is_bit_identical(mb, oracle, name):
    """real_type = float (the reference's `<float>` instantiations, what dust2 /
    odin2 models compiled for single precision call): logf, expf, powf, cosf and
    lgammaf return glibc's bits on the device (detail/glibc_math.cuh: every float
    checked against the host libm, exhaustively, in tests/test_glibc_math.py), so
    every sampler is bit-identical -- draws AND final states, the integer-valued
    ones whose accept tests subtract large lgammaf terms included (round 1: 0.5-3%
    of those draws differed)."""
    n, ns = 8192, 32
    pars = sampler_cases(n)[name]
    y, st, _ = gpu_sample(mb, name, n, ns, pars, rt="f32")
    ost = oracle.seed_states(7, n)
    exp = oracle.sample(name, ost, ns, pars, "f32")
    assert np.array_equal(st, ost), "%d streams ended in a different state" % int((st != ost).any(axis=1).sum())
    same = (y == exp) | (np.isnan(y) & np.isnan(exp))
    assert same.all(), (name, float(1 - same.mean()), y[~same][:3], exp[~same][:3])


def test_empty_requests(mb):
    """n_samples = 0 returns an empty result and leaves every stream alone; a
    density of no observations is empty (and its device-side sum is 0)."""
    rng = mb.monty_rng_create(5, 3)
    before = states_of(mb, rng)
    y = np.asarray(mb.monty_random_n_real(0, rng))
    assert y.size == 0
    z = np.asarray(mb.monty_random_n_binomial(0, 10.0, 0.5, rng))
    assert z.size == 0
    assert np.array_equal(states_of(mb, rng), before)
    d = np.asarray(mb.density.density_poisson(np.zeros(0, dtype=np.int32), np.zeros(0), True))
    assert d.size == 0


def test_scalar_and_vector_parameters_agree(mb):
    """ref: tests/testthat/test-random.R:77-92: a parameter vector behaves like
    independent single-stream generators; a scalar is broadcast."""
    n = 65
    a = mb.monty_rng_create(n, 3)
    b = mb.monty_rng_create(n, 3)
    x = mb.monty_random_n_binomial(9, 100.0, 0.3, a)
    y = mb.monty_random_n_binomial(9, np.full(n, 100.0), np.full(n, 0.3), b)
    assert np.array_equal(x, y)
    # stream k of an n-stream generator == a 1-stream generator jumped k times
    k = 40
    one = mb.monty_rng_create(1, 3)
    mb.monty_rng_jump(one, k)
    assert np.array_equal(mb.monty_random_n_binomial(9, 100.0, 0.3, one), x[:, k])


def test_exponential_rate_division_is_ieee(mb):
    """exponential_rate divides by a per-stream reciprocal refined once per
    launch (samplers.cuh: ConstDivisor) -- the quotient must be the correctly
    rounded e / rate for every rate, awkward mantissas and extreme exponents
    (which take the plain division) included.  exponential_mean(1) returns the
    same e (e * 1 is exact), so numpy's IEEE division is the checker."""
    rs = np.random.default_rng(5)
    n = 4096
    rate = np.exp(rs.uniform(-40, 40, n))
    rate[:8] = [1.0, 3.0, 1 / 3, 2.0 ** 52 - 1, np.nextafter(2.0, 1.0), np.nextafter(1.0, 2.0), 0.1, 7.0]
    rate[8:16] = [1e-300, 1e300, 2.0 ** -901, 2.0 ** 901, 2.0 ** -900, 2.0 ** 900, 5e-324, 1e308]
    mant = rs.integers(0, 1 << 52, 256, dtype=np.uint64) | np.uint64(0x3FF0000000000000)
    rate[16:16 + 256] = mant.view(np.float64)
    rate[272:272 + 64] = np.nextafter(2.0, 1.0) * 2.0 ** rs.integers(-60, 60, 64)
    a = mb.monty_rng_create(n, 17)
    b = mb.monty_rng_create(n, 17)
    e = np.asarray(mb.monty_random_n_exponential_mean(33, 1.0, a))
    with np.errstate(over="ignore", under="ignore"):
        expect = e / rate
    got = np.asarray(mb.monty_random_n_exponential_rate(33, rate, b))
    assert np.array_equal(got, expect)
    assert np.array_equal(states_of(mb, a), states_of(mb, b))


def test_full_size_config_spot_check(mb, oracle):
    """BASELINE config C2 at full size (2^20 streams x 1000 draws, double,
    ziggurat), device resident: bit-exact on 64 streams spread over the
    generator, final states included, plus moments of the whole 2^30 sample."""
    import torch
    S, D = 1 << 20, 1000
    rng = mb.monty_rng_create(S, 42)
    first = states_of(mb, rng)
    dev = torch.device("cuda", 0)
    mean = torch.zeros(1, dtype=torch.float64, device=dev)
    sd = torch.ones(1, dtype=torch.float64, device=dev)
    out = rng.sample_dev("normal_ziggurat", D, [mean, sd])
    rng.sync()
    last = states_of(mb, rng)
    pick = np.linspace(0, S - 1, 64).astype(np.int64)
    st = first[pick].copy()
    exp = oracle.sample("normal_ziggurat", st, D, [np.array([0.0]), np.array([1.0])])
    got = out[torch.from_numpy(pick).to(dev)].cpu().numpy()
    assert np.array_equal(got, exp)
    assert np.array_equal(last[pick], st)
    m = out.mean().item()
    v = out.var().item()
    assert abs(m) < 2e-4 and abs(v - 1) < 2e-4


@pytest.mark.parametrize("name,draws", [("binomial", 1), ("poisson", 1), ("binomial", 16), ("hypergeometric", 16),
                                        ("gamma_scale", 16), ("truncated_normal", 16)])
def test_full_size_sir_step_spot_check(mb, oracle, name, draws):
    """BASELINE configs C3 / C4 at full size (2^22 streams, per-stream parameters
    spanning every regime; the regime-sorting kernel), device resident: 4096
    streams spread over the generator are checked against the reference --
    integer-valued draws and final states bit for bit."""
    import torch
    S = 1 << 22 if draws == 1 else 1 << 20
    rng = mb.monty_rng_create(S, 42)
    first = states_of(mb, rng)
    dev = torch.device("cuda", 0)
    pars = sampler_cases(S)[name]
    pd = [torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in pars]
    out = rng.sample_dev(name, draws, pd)
    rng.sync()
    last = states_of(mb, rng)
    pick = np.linspace(0, S - 1, 4096).astype(np.int64)
    st = first[pick].copy()
    exp = oracle.sample(name, st, draws, [p[pick] if p.size == S else p for p in pars])
    got = out[torch.from_numpy(pick).to(dev)].cpu().numpy()
    assert np.array_equal(last[pick], st), "final generator states differ"
    if name in INTEGER_VALUED:
        assert np.array_equal(got, exp)
    else:
        assert_continuous_parity(name, got, exp, [p[pick] if p.size == S else p for p in pars], pick.size)


def test_device_resident_matches_host_path(mb):
    import torch
    n, ns = 1000, 50
    dev = torch.device("cuda", 0)
    pars = sampler_cases(n)["gamma_scale"]
    a = mb.monty_rng_create(n, 5)
    b = mb.monty_rng_create(n, 5)
    host = mb.monty_random_n_gamma_scale(ns, pars[0], pars[1], a)
    out = b.sample_dev("gamma_scale", ns, [torch.from_numpy(p).to(dev) for p in pars])
    b.sync()
    assert np.array_equal(out.cpu().numpy(), host.T)
    assert mb.monty_rng_state(a) == mb.monty_rng_state(b)


def test_float_output_mode(mb):
    import torch
    n, ns = 512, 40
    dev = torch.device("cuda", 0)
    a = mb.monty_rng_create(n, 5, real_type="f32")
    b = mb.monty_rng_create(n, 5, real_type="f32")
    p = [torch.zeros(1, dtype=torch.float64, device=dev), torch.ones(1, dtype=torch.float64, device=dev)]
    o64 = a.sample_dev("normal_ziggurat", ns, p)
    o32 = torch.empty((n, ns), dtype=torch.float32, device=dev)
    b.sample_dev("normal_ziggurat", ns, p, out=o32)
    a.sync(); b.sync()
    assert torch.equal(o64.float(), o32) and torch.equal(o32.double(), o64)


# ---- errors ---------------------------------------------------------------------
from test_oracle import ERROR_CASES  # noqa: E402


@pytest.mark.parametrize("name,pars,message,offender", ERROR_CASES)
def test_error_text_and_partial_advance(mb, name, pars, message, offender):
    """Same text as the reference, streams before the offender advanced, the
    offender and later streams untouched (ref: src/random.cpp:211-228)."""
    rng = mb.monty_rng_create(4, 5)
    before = states_of(mb, rng)
    with pytest.raises(mb.MontyError) as e:
        mb.random._sample(rng, name, 3, pars)
    assert str(e.value) == message
    after = states_of(mb, rng)
    changed = (after != before).any(axis=1)
    assert changed[:offender].all() and not changed[offender:].any()
    # ... and the advanced streams drew exactly what the reference draws
    if offender:
        from oracle_lib import Oracle, OracleError, have_reference
        orc = Oracle("reference" if have_reference() else "port")
        ost = orc.seed_states(5, 4)
        with pytest.raises(OracleError):
            orc.sample(name, ost, 3, pars)
        assert np.array_equal(after, ost)


@pytest.mark.parametrize("ns", [1, 16])
def test_errors_in_the_regime_sorted_kernels(mb, oracle, ns):
    """An invalid parameter in the middle of 5000 per-stream parameter sets (the regime-sorting
    kernel, its persistent-lane variant at one draw per stream): the reference's message,
    streams before the offender advanced exactly like the reference, the offender and later
    streams untouched (ref: src/random.cpp:211-228)."""
    n, bad = 5000, 3333
    pars = [p.copy() for p in sampler_cases(n)["binomial"]]
    pars[1][bad] = 1.5
    rng = mb.monty_rng_create(n, 7)
    before = states_of(mb, rng)
    with pytest.raises(mb.MontyError, match=r"Invalid call to binomial with n = .*, p = 1\.5"):
        mb.random._sample(rng, "binomial", ns, pars)
    after = states_of(mb, rng)
    ost = oracle.seed_states(7, n)
    good = [p[:bad] for p in pars]
    oracle.sample("binomial", ost[:bad], ns, good)          # a view: advanced in place
    assert np.array_equal(after[:bad], ost[:bad])
    assert np.array_equal(after[bad:], before[bad:])


def test_parameter_length_errors(mb):
    """ref: tests/testthat/test-random.R:105-120, src/random.cpp:21-42."""
    rng = mb.monty_rng_create(5, 1)
    with pytest.raises(mb.MontyError, match="Expected 'rate' to have length 5 or 1, not 3"):
        mb.monty_random_exponential_rate(np.ones(3), rng)
    one = mb.monty_rng_create(1, 1)
    with pytest.raises(mb.MontyError, match="Expected 'prob' to have length 1, not 2"):
        mb.monty_random_binomial(10.0, np.array([0.1, 0.2]), one)


def test_device_api_reports_errors_on_sync(mb):
    import torch
    dev = torch.device("cuda", 0)
    rng = mb.monty_rng_create(64, 1)
    lam = torch.ones(64, dtype=torch.float64, device=dev)
    lam[17] = -2.0
    out = rng.sample_dev("poisson", 3, [lam])
    with pytest.raises(mb.MontyError, match="Invalid call to Poisson with lambda = -2"):
        rng.sync()
    assert torch.isnan(out[17]).all() and not torch.isnan(out[16]).any()


def test_multinomial(mb, oracle, golden):
    """ref: tests/testthat/test-rng.R:530-663, src/random.cpp:884-934."""
    rng = mb.monty_rng_create(9, 20240607)
    prob = np.array([0.1, 0.0, 0.4, 0.5, 0.2])
    y = mb.monty_random_n_multinomial(4, 100.0, prob, rng)
    assert y.shape == (5, 4, 9)
    assert np.array_equal(y.transpose(2, 1, 0), golden["multinomial/shared"])
    assert np.array_equal(states_of(mb, rng), golden["multinomial/shared/state"])
    # per-stream probability matrix and sizes
    pm = np.abs(np.random.RandomState(1).randn(5, 9))
    size = np.arange(9) * 50.0
    y = mb.monty_random_multinomial(size, pm, rng)
    ost = golden["multinomial/shared/state"].copy()
    exp = oracle.multinomial(ost, 1, size, pm)
    assert np.array_equal(y.T, exp[:, 0, :])
    with pytest.raises(mb.MontyError, match="Negative prob passed to multinomial"):
        mb.monty_random_multinomial(10.0, np.array([0.5, -0.1]), rng)
    with pytest.raises(mb.MontyError, match="No positive prob in call to multinomial"):
        mb.monty_random_multinomial(10.0, np.array([0.0, 0.0]), rng)


# ---- deterministic mode (ref: tests/testthat/test-rng.R:675-764, 952-967, ...) ---
@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("name", [n for n in ALL if n != "cauchy"])
def test_deterministic_mode(mb, golden, oracle, name, rt):
    """Every sampler returns its expectation and leaves the stream alone
    (except random_real and the unbounded truncated normal, which draw in the
    reference too)."""
    pars = sampler_cases(64)[name]
    rng = mb.monty_rng_create(64, 20240607, deterministic=True, real_type=rt)
    before = states_of(mb, rng)
    y = np.ascontiguousarray(mb.random._sample(rng, name, 5, pars).T)
    exp = golden["sample/%s/%s/det" % (rt, name)]
    st = states_of(mb, rng)
    assert np.array_equal(st, golden["sample/%s/%s/det/state" % (rt, name)])
    if name not in ("real", "truncated_normal"):
        assert np.array_equal(st, before)
    tol = 1e-6 if rt == "f32" else 1e-13
    if name == "truncated_normal":
        tol *= 1e3      # (d_min - d_max) / (z_max - z_min): erf differences cancel
    assert np.allclose(y, exp, rtol=tol, atol=tol, equal_nan=True), name


def test_deterministic_errors(mb):
    rng = mb.monty_rng_create(2, 1, deterministic=True)
    with pytest.raises(mb.MontyError, match="Can't use Cauchy distribution deterministically; it has no mean"):
        mb.monty_random_cauchy(0.0, 1.0, rng)
    with pytest.raises(mb.MontyError, match=r"Invalid call to binomial with n = -1\.000000"):
        mb.monty_random_binomial(-1.0, 0.5, rng)
    with pytest.raises(mb.MontyError, match="Invalid call to binomial with n = 10, p = 1.5, q = -0.5"):
        mb.monty_random_binomial(10.0, 1.5, rng)
    assert mb.monty_random_binomial(-1e-10, 0.5, rng)[0] == 0     # round-off in n is forgiven
    assert np.array_equal(mb.monty_random_n_poisson(3, 2.5, rng), np.full((3, 2), 2.5))
    y = mb.monty_random_multinomial(10.0, np.array([0.2, 0.3, 0.5]), rng)
    assert np.allclose(y[:, 0], [2.0, 3.0, 5.0])


def test_sample_dev_is_ordered_with_torchs_stream(mb):
    """Parameters written by torch kernels that are still pending when
    sample_dev is called, and a consumer on torch's stream right after it: the
    handle's private stream must wait for the former and be waited for by the
    latter (the launch used to race with both)."""
    import torch
    dev = torch.device("cuda", 0)
    n, ns = 1 << 18, 8
    a = mb.monty_rng_create(n, 11)
    b = mb.monty_rng_create(n, 11)
    lam_host = np.full(n, 7.5)
    ref = np.asarray(mb.monty_random_n_poisson(ns, lam_host, a))          # host path: fully synchronous
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        lam = torch.zeros(n, dtype=torch.float64, device=dev)
        big = torch.empty(1 << 27, dtype=torch.float64, device=dev)
        for _ in range(20):
            big.normal_()                      # keeps the stream busy ahead of the parameter write
        lam.add_(7.5)                          # the parameter is only valid after this kernel
        out = b.sample_dev("poisson", ns, [lam])
        total = out.sum()                      # consumer on torch's stream, no explicit sync
        got_sum = float(total.item())          # (.item() waits for the stream it is issued on)
    b.sync()
    assert np.array_equal(out.cpu().numpy(), ref.T)
    assert got_sum == float(ref.sum())
