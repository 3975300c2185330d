"""GPU parity of the dedicated ziggurat kernel (csrc/ziggurat_kernel.cuh)
against the reference's CPU implementation on identical seeds.

The kernel re-orders nothing inside a stream, decides wedges in single
precision outside an ambiguity band and in the reference's own double
expression inside it, and runs the tail out of line -- so the bar is
  * final generator states: bit-exact (every stream consumed exactly the
    reference's number of words);
  * draws: bit-exact in double, tail draws (|z| > x[1] = 3.65..., 3e-4 of
    draws) included -- their log() returns glibc's bits (include/monty_b200/detail/glibc_math.cuh);
    float real_type: bit-exact as well (logf returns glibc's bits).
Shapes cover ragged first/last windows (n_samples not a multiple of 32, odd
n_samples -> scalar store path), n_streams not a multiple of 32, one warp-tile
problems and enough draws (3.3e7) to hit the tail ~1e4 and the ambiguity band
~1e2 times.
"""
import numpy as np
import pytest

from gpu_util import gpu_sample, states_of, ulp_distance

pytestmark = pytest.mark.gpu

X1 = 3.6541528853610088    # x[1] of the reference's table: |z| beyond it comes from the tail


@pytest.fixture(scope="module")
def mb():
    import monty_b200
    return monty_b200


def _params(kind, n):
    if kind == "standard":
        return [np.array([0.0]), np.array([1.0])]
    if kind == "scalar":
        return [np.array([1.5]), np.array([0.25])]
    i = np.arange(n, dtype=np.float64)
    return [np.sin(i) * 5.0, 0.5 + (i % 7) / 4.0]


def _check(y, exp, pars, rt):
    same = y == exp
    if same.all():
        return 0.0
    assert False, "draws must be bit-identical (%s): %r vs %r" % (rt, y[~same][:3], exp[~same][:3])
    n = y.shape[0]
    mean = np.broadcast_to(pars[0], (n,))[:, None]
    sd = np.broadcast_to(pars[1], (n,))[:, None]
    z = np.abs((exp - mean) / sd)
    bad = ~same
    # every differing draw is a tail draw ...
    assert (z[bad] > X1 * (1 - 1e-6)).all(), "a non-tail draw differs: %r vs %r" % (y[bad][:3], exp[bad][:3])
    # ... and differs by libm rounding only
    if rt == "f64":
        tol = 8 * np.finfo(np.float64).eps * (np.abs(exp) + np.abs(mean))
    else:
        tol = 4 * np.finfo(np.float32).eps * (np.abs(exp) + np.abs(mean))
    assert (np.abs(y - exp) <= tol)[bad].all()
    return bad.mean()


SHAPES = [(32768, 1000), (1000, 2), (77, 31), (4099, 33), (513, 1001), (2048, 64), (100000, 10), (31, 4097)]


@pytest.mark.parametrize("rt", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["standard", "scalar", "vector"])
@pytest.mark.parametrize("shape", SHAPES)
def test_ziggurat_kernel_parity(mb, oracle, shape, kind, rt):
    n, ns = shape
    if kind != "standard" and n * ns > 5_000_000 and rt == "f32":
        pytest.skip("covered by the f64 case")
    pars = _params(kind, n)
    y, st, _ = gpu_sample(mb, "normal_ziggurat", n, ns, pars, seed=11, rt=rt)
    ost = oracle.seed_states(11, n)
    exp = oracle.sample("normal_ziggurat", ost, ns, pars, real_type=rt)
    assert np.array_equal(st, ost), "final generator states differ"
    frac = _check(y, exp, pars, rt)
    assert frac < 2e-3


def test_ziggurat_device_resident_float_output(mb, oracle):
    """float real_type with float32 device output, odd and even row lengths."""
    import torch
    dev = torch.device("cuda", 0)
    for n, ns in [(4096, 1000), (300, 37), (2048, 36)]:
        rng = mb.monty_rng_create(n, 3, real_type="f32")
        p = [torch.zeros(1, dtype=torch.float64, device=dev), torch.ones(1, dtype=torch.float64, device=dev)]
        out = torch.empty((n, ns), dtype=torch.float32, device=dev)
        rng.sample_dev("normal_ziggurat", ns, p, out=out)
        rng.sync()
        ost = oracle.seed_states(3, n)
        exp = oracle.sample("normal_ziggurat", ost, ns, [np.array([0.0]), np.array([1.0])], real_type="f32")
        assert np.array_equal(states_of(mb, rng), ost)
        _check(out.cpu().numpy().astype(np.float64), exp, [np.array([0.0]), np.array([1.0])], "f32")


def test_ziggurat_continues_a_stream(mb, oracle):
    """Two consecutive calls continue the streams exactly like one long call
    split at the same place (the parked-candidate state never leaks across
    calls)."""
    n = 257
    pars = _params("standard", n)
    rng = mb.monty_rng_create(n, 5)
    a = np.asarray(mb.random._sample(rng, "normal_ziggurat", 100, pars)).reshape(100, n).T
    b = np.asarray(mb.random._sample(rng, "normal_ziggurat", 77, pars)).reshape(77, n).T
    ost = oracle.seed_states(5, n)
    ea = oracle.sample("normal_ziggurat", ost, 100, pars)
    eb = oracle.sample("normal_ziggurat", ost, 77, pars)
    _check(a, ea, pars, "f64")
    _check(b, eb, pars, "f64")
    assert np.array_equal(states_of(mb, rng), ost)


def test_ziggurat_matches_lockstep_kernel(mb):
    """n_samples = 1 takes the generic lock-step kernel; a batch of single
    draws must equal one batched call."""
    n = 1024
    pars = _params("vector", n)
    one = mb.monty_rng_create(n, 9)
    many = mb.monty_rng_create(n, 9)
    cols = [np.asarray(mb.random._sample(one, "normal_ziggurat", None, pars)) for _ in range(40)]
    batch = np.asarray(mb.random._sample(many, "normal_ziggurat", 40, pars)).reshape(40, n)
    assert np.array_equal(states_of(mb, one), states_of(mb, many))
    d = ulp_distance(np.stack(cols), batch)
    assert (d == 0).all()


@pytest.mark.parametrize("n_samples", [(1 << 24), (1 << 24) + 2])
def test_longest_rows(mb, oracle, n_samples):
    """The batched kernel addresses its write-out with 32-bit byte offsets from
    the first row of a warp-tile, which bounds it at n_samples <= 2^24; one
    more and the generic kernel (64-bit addressing) takes over.  34 streams
    (a full warp-tile and a ragged one) x 1.7e7 draws, device resident: rows
    0, 30, 31, 32 and 33 against the reference, final states included."""
    import torch
    n = 34
    dev = torch.device("cuda", 0)
    rng = mb.monty_rng_create(n, 13)
    first = states_of(mb, rng)
    p = [torch.zeros(1, dtype=torch.float64, device=dev), torch.ones(1, dtype=torch.float64, device=dev)]
    out = rng.sample_dev("normal_ziggurat", n_samples, p)
    rng.sync()
    last = states_of(mb, rng)
    pick = np.array([0, 30, 31, 32, 33])
    st = first[pick].copy()
    exp = oracle.sample("normal_ziggurat", st, n_samples, [np.array([0.0]), np.array([1.0])])
    got = out[torch.from_numpy(pick).to(dev)].cpu().numpy()
    del out
    assert np.array_equal(last[pick], st)
    _check(got, exp, [np.array([0.0]), np.array([1.0])], "f64")
