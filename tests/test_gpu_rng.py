"""GPU parity, generator level: golden vectors, seeding by the doubling
ladder, jump / long_jump, state export / import, sharding.  Everything here is
integer work and must be BIT-EXACT against the reference."""
import os

import numpy as np
import pytest

from gpu_util import states_of
from oracle_lib import GEN_NAMES

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mb():
    import monty_b200
    return monty_b200


@pytest.mark.parametrize("name", GEN_NAMES)
def test_golden_protocol_on_device(mb, name):
    """ref: tests/testthat/test-xoshiro.R, src/test_rng.cpp:21-38 -- for all 12
    generator variants, not only the xoshiro256+ file the R test reads."""
    vals, st = mb.test_xoshiro_run(generator=name)
    with open(os.path.join(HERE, "golden", "xoshiro-ref", name)) as fh:
        assert vals + st == fh.read().split()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 1000, 4097])
def test_seeding_ladder_matches_sequential_jumps(mb, oracle, n):
    """prng(n, seed): stream i is i sequential jumps from the seed state
    (ref: prng.hpp:38-53); here produced in ceil(log2 n) launches."""
    rng = mb.monty_rng_create(n, 42)
    assert np.array_equal(states_of(mb, rng), oracle.seed_states(42, n))


def test_seeding_matches_reference_fixture(mb, golden):
    rng = mb.monty_rng_create(100, 42)
    assert np.array_equal(states_of(mb, rng), golden["seed_states_42_100"])


def test_large_generator_spot_check(mb, oracle):
    """2^20 streams; the oracle jumps sequentially, so compare the first 3000
    streams and a shard created at an offset."""
    n = 1 << 20
    rng = mb.monty_rng_create(n, 42)
    st = states_of(mb, rng)
    assert np.array_equal(st[:3000], oracle.seed_states(42, 3000))
    # a shard [off, off + 257) of the same generator, as another GPU would hold it
    off = 777777
    shard = mb.monty_rng_create(257, 42, stream_offset=off)
    assert np.array_equal(states_of(mb, shard), st[off:off + 257])
    assert len(np.unique(st, axis=0)) == n


def test_raw_seed_then_jump_from_last(mb, oracle, golden):
    """ref: tests/testthat/test-rng.R:435-509, prng.hpp:45-51."""
    words = golden["raw_seed_words"]
    rng = mb.monty_rng_create(10, seed=words.tobytes())
    assert np.array_equal(states_of(mb, rng), golden["raw_seed_states_10"])
    # fewer streams than supplied states: just the first ones
    rng = mb.monty_rng_create(2, seed=words.tobytes())
    assert np.array_equal(states_of(mb, rng), words.reshape(3, 4)[:2])
    # offsets straddling the supplied states
    for off in (1, 2, 3, 5):
        rng = mb.monty_rng_create(4, seed=words.tobytes(), stream_offset=off)
        assert np.array_equal(states_of(mb, rng), golden["raw_seed_states_10"][off:off + 4])


def test_jump_and_long_jump(mb, oracle, golden):
    rng = mb.monty_rng_create(33, 7)
    mb.monty_rng_jump(rng, 3)
    assert np.array_equal(states_of(mb, rng), golden["jump3_seed7_33"])
    mb.monty_rng_long_jump(rng, 2)
    assert np.array_equal(states_of(mb, rng), golden["longjump2_after"])
    # many jumps at the cost of one: n = 1000
    rng = mb.monty_rng_create(5, 3)
    mb.monty_rng_jump(rng, 1000)
    exp = oracle.seed_states(3, 5)
    oracle.jump(exp, 1000)
    assert np.array_equal(states_of(mb, rng), exp)


def test_jump_on_raw_state(mb, oracle):
    """monty_rng_jump(raw, n) round-trips through a temporary generator
    (ref: R/random.R:131-141)."""
    st = oracle.seed_states(9, 3)
    out = mb.monty_rng_jump(st.tobytes(), 2)
    oracle.jump(st, 2)
    assert out == st.tobytes()


def test_state_export_import_roundtrip(mb, oracle):
    """ref: tests/testthat/test-random.R:32-45, test-rng.R:415-476."""
    a = mb.monty_rng_create(7, 1)
    raw = mb.monty_rng_state(a)
    assert isinstance(raw, bytes) and len(raw) == 32 * 7
    x = mb.monty_random_n_real(5, a)
    b = mb.monty_rng_create(7, 999)
    mb.monty_rng_set_state(raw, b)
    assert np.array_equal(mb.monty_random_n_real(5, b), x)
    with pytest.raises(mb.MontyError, match=r"'value' must be a raw vector of length 224 \(but was 32\)"):
        mb.monty_rng_set_state(raw[:32], b)


def test_distributed_state_is_long_jump_ladder(mb, oracle):
    """ref: tests/testthat/test-random.R:66-74, R/random.R:161-171."""
    states = mb.monty_rng_distributed_state(seed=42, n_nodes=4)
    exp = oracle.seed_states(42, 1)
    for s in states:
        assert s == exp.tobytes()
        oracle.jump(exp, 1, long=True)
    # the sharded constructor: GPU g = seed long-jumped g times, then a jump ladder
    g = mb.monty_rng_create(5, 42, long_jumps=3)
    base = oracle.seed_states(42, 1)
    oracle.jump(base, 3, long=True)
    assert np.array_equal(states_of(mb, g), oracle.seed_states(0, 5, base.ravel()))


def test_raw_output_all_scramblers(mb, golden):
    for gen in range(3):
        rng = mb.monty_rng_create(5, 20240607)
        out = np.zeros((5, 11), dtype=np.uint64)
        import ctypes as C
        mb._lib.check(mb._lib.lib.mb200_rng_raw(rng.handle, gen, 11, out.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(out, golden["raw_gen%d" % gen])
        assert np.array_equal(states_of(mb, rng), golden["raw_gen%d_state" % gen])


def test_successive_calls_continue_the_stream(mb):
    """ref: tests/testthat/test-rng.R:368-377."""
    a = mb.monty_rng_create(3, 5)
    b = mb.monty_rng_create(3, 5)
    x = mb.monty_random_n_real(10, a)
    y = np.vstack([mb.monty_random_n_real(4, b), mb.monty_random_n_real(6, b)])
    assert np.array_equal(x, y)


def test_result_shapes(mb):
    """ref: src/random.cpp:84-86, 197-201: dim = c(n_samples, n_streams) iff
    n_streams > 1 or preserve_stream_dimension."""
    one = mb.monty_rng_create(1, 1)
    assert mb.monty_random_real(one).shape == (1,)
    assert mb.monty_random_n_real(5, one).shape == (5,)
    keep = mb.monty_rng_create(1, 1, preserve_stream_dimension=True)
    assert mb.monty_random_n_real(5, keep).shape == (5, 1)
    many = mb.monty_rng_create(3, 1)
    y = mb.monty_random_n_real(5, many)
    assert y.shape == (5, 3) and y.flags.f_contiguous
    assert mb.monty_random_real(many).shape == (3,)
