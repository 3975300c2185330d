#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz from the UNMODIFIED reference.

Runs the reference's own headers (compiled as oracle/_ref/libmonty_ref.so, see
oracle/Makefile) on fixed seeds and parameter sets and stores inputs, outputs
and final generator states.  The fixtures pin the C oracle (tests/
test_oracle.py) and the CUDA path (tests/test_gpu_*.py) on machines where
/root/reference is not available.  Re-run in the build container:

    make -C oracle ref && python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Oracle, counter_hash  # noqa: E402
from cases import sampler_cases, density_cases  # noqa: E402

N_STREAMS, N_SAMPLES, SEED = 64, 5, 20240607


def main():
    ref = Oracle("reference")
    out = {}
    out["seed_states_42_100"] = ref.seed_states(42, 100)
    st = ref.seed_states(7, 33)
    ref.jump(st, 3)
    out["jump3_seed7_33"] = st.copy()
    ref.jump(st, 2, long=True)
    out["longjump2_after"] = st.copy()
    raw = (np.arange(12, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(7))
    out["raw_seed_words"] = raw
    out["raw_seed_states_10"] = ref.seed_states(0, 10, raw)
    for gen in range(3):
        s = ref.seed_states(SEED, 5)
        out["raw_gen%d" % gen] = ref.raw(gen, s, 11)
        out["raw_gen%d_state" % gen] = s
    for rt in ("f64", "f32"):
        for name, pars in sampler_cases(N_STREAMS).items():
            for det in (False, True):
                if det and name == "cauchy":
                    continue
                s = ref.seed_states(SEED, N_STREAMS)
                y = ref.sample(name, s, N_SAMPLES, pars, rt, det)
                key = "sample/%s/%s/%s" % (rt, name, "det" if det else "sto")
                out[key] = y
                out[key + "/state"] = s
    for rt in ("f64", "f32"):
        for name, (x, pars) in density_cases(129).items():
            for lg in (True, False):
                out["density/%s/%s/%d" % (rt, name, lg)] = ref.density(name, x, pars, lg, rt)
    s = ref.seed_states(SEED, 9)
    prob = np.array([0.1, 0.0, 0.4, 0.5, 0.2])
    out["multinomial/shared"] = ref.multinomial(s, 4, np.array([100.0]), prob)
    out["multinomial/shared/state"] = s
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
