#!/usr/bin/env python
"""Generate tests/golden/generator_vectors.npz from the UNMODIFIED reference: every
sampler instantiated on the reference's other generators -- generator<float> =
xoshiro128+ (inst/include/monty/random/random.hpp:31-47) first of all -- through
oracle/_ref/libmonty_ref.so (ref_generator_draw in oracle/ref_harness.cpp).  The vectors
pin the device-side API on those generators (tests/test_gpu_generators.py) where
/root/reference is not available.  Re-run in the build container:

    make -C oracle ref && python tools/make_golden_generators.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Oracle  # noqa: E402
from cases import sampler_cases  # noqa: E402

# (generator, real type): the two defaults, the defaults crossed, one of each other family
COMBOS = [("xoshiro128plus", "f32"), ("xoshiro128plus", "f64"), ("xoshiro256plus", "f32"), ("xoshiro256plus", "f64"),
          ("xoshiro128starstar", "f32"), ("xoshiro256starstar", "f64"), ("xoroshiro128plusplus", "f64"),
          ("xoshiro512starstar", "f64")]
N_STREAMS, N_SAMPLES, SEED = 48, 6, 20240607


def main():
    ref = Oracle("reference")
    out = {}
    for gen, rt in COMBOS:
        for name, pars in sampler_cases(N_STREAMS).items():
            y, st = ref.generator_draw(gen, name, SEED, N_STREAMS, N_SAMPLES, pars, rt, jumps=1, long_jumps=1)
            out["%s/%s/%s" % (gen, rt, name)] = y
            out["%s/%s/%s/state" % (gen, rt, name)] = st
    path = os.path.join(ROOT, "tests", "golden", "generator_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
