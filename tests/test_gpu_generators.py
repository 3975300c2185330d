"""The device-side monty::random API (include/monty_b200/random.cuh) on the reference's
OTHER generators.  A dust2 / odin2 model compiled for real_type = float draws from
monty::random::generator<float>, which is xoshiro128+ (4 x u32 state; the ziggurat takes
its layer from a second word, inst/include/monty/random/normal_ziggurat.hpp:36-65);
any of the twelve state types can be named explicitly.  The samplers of
detail/samplers.cuh are templates on the generator, the states are seeded and jumped by
monty_b200::device_prng<G> (include/monty_b200/device_prng.cuh).

Checked against the reference's own templates instantiated on the same generator
(oracle/ref_harness.cpp: ref_generator_draw): every sampler, draws AND final states
bit-identical -- against the committed vectors (tools/make_golden_generators.py) and,
where oracle/_ref is present, against the live reference on more streams."""
import os

import numpy as np
import pytest

from cases import sampler_cases

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
COMBOS = [("xoshiro128plus", "f32"), ("xoshiro128plus", "f64"), ("xoshiro256plus", "f32"), ("xoshiro256plus", "f64"),
          ("xoshiro128starstar", "f32"), ("xoshiro256starstar", "f64"), ("xoroshiro128plusplus", "f64"),
          ("xoshiro512starstar", "f64")]
ALL = sorted(sampler_cases(4))


@pytest.fixture(scope="module")
def ex():
    from monty_b200 import examples
    return examples


@pytest.fixture(scope="module")
def vectors():
    return np.load(os.path.join(HERE, "golden", "generator_vectors.npz"))


def same(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


@pytest.mark.parametrize("gen,rt", COMBOS)
def test_every_sampler_matches_the_committed_reference_vectors(ex, vectors, gen, rt):
    n, ns, seed = 48, 6, 20240607
    for name in ALL:
        y, st = ex.generator_draw(gen, name, seed, n, ns, sampler_cases(n)[name], rt, jumps=1, long_jumps=1)
        exp, est = vectors["%s/%s/%s" % (gen, rt, name)], vectors["%s/%s/%s/state" % (gen, rt, name)]
        assert np.array_equal(st, est), (gen, rt, name, "final states differ")
        ok = same(y, exp)
        assert ok.all(), (gen, rt, name, float(1 - ok.mean()), y[~ok][:3], exp[~ok][:3])


@pytest.mark.parametrize("gen,rt", COMBOS[:2] + COMBOS[4:])
def test_every_sampler_matches_the_live_reference(ex, reference, gen, rt):
    """2000 streams x 24 draws (all regimes of every sampler), other seed, no jumps."""
    n, ns, seed = 2000, 24, 99
    for name in ALL:
        pars = sampler_cases(n)[name]
        y, st = ex.generator_draw(gen, name, seed, n, ns, pars, rt)
        exp, est = reference.generator_draw(gen, name, seed, n, ns, pars, rt)
        assert np.array_equal(st, est), (gen, rt, name, "final states differ")
        ok = same(y, exp)
        assert ok.all(), (gen, rt, name, float(1 - ok.mean()), y[~ok][:3], exp[~ok][:3])


@pytest.mark.parametrize("gen,rt", COMBOS)
def test_seeding_and_jumps(ex, reference, gen, rt):
    """device_prng<G>(n, seed): stream i is i jumps from the seed (hdr/prng.hpp:38-53);
    jump() / long_jump() of all streams (hdr/generator.hpp:59-106, each generator's own
    polynomial) -- no draws, states only."""
    for jumps, long_jumps in ((0, 0), (2, 0), (0, 1), (3, 2)):
        _, st = ex.generator_draw(gen, "real", 42, 257, 0, [], rt, jumps=jumps, long_jumps=long_jumps)
        _, est = reference.generator_draw(gen, "real", 42, 257, 0, [], rt, jumps=jumps, long_jumps=long_jumps)
        assert np.array_equal(st, est), (gen, jumps, long_jumps)


@pytest.mark.parametrize("gen", ["xoshiro128plus", "xoshiro256plus", "xoroshiro128plusplus", "xoshiro512starstar"])
def test_seed_vector_export_import(ex, reference, gen):
    """device_prng<G>(n, seed words): full states from the words while they last, then jump()
    from the last one (hdr/prng.hpp:38-53); export_state -> import_state into a second
    container -> jump -> export_state (hdr/prng.hpp:92-122)."""
    n_words, wt = ex.generator_layout(gen)
    rs = np.random.default_rng(5)
    for n_states, n_streams in ((1, 9), (3, 9), (9, 9), (12, 9)):
        words = rs.integers(1, np.iinfo(wt).max, n_states * n_words, dtype=wt)
        for rj in (False, True):
            got = ex.generator_seed_vector(gen, words, n_streams, rj)
            exp = reference.generator_seed_vector(gen, words, n_streams, rj)
            assert np.array_equal(got, exp), (gen, n_states, rj)


def test_default_generators(ex, vectors):
    """generator<float> is xoshiro128+ (16-byte states), generator<double> xoshiro256+."""
    _, st = ex.generator_draw("xoshiro128plus", "real", 1, 3, 1, [], "f32")
    assert st.dtype == np.uint32 and st.shape == (3, 4)
    _, st = ex.generator_draw("xoshiro256plus", "real", 1, 3, 1, [], "f64")
    assert st.dtype == np.uint64 and st.shape == (3, 4)
