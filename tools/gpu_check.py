#!/usr/bin/env python
"""Quick GPU-vs-oracle parity sweep (development aid; the formal checks are
tests/test_gpu_*.py).  Prints per sampler: exact-match fraction, max ulp
distance, state equality."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import monty_b200 as mb
from oracle_lib import Oracle, counter_hash, have_reference, GEN_NAMES


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    ia = a.view(np.int64).copy(); ib = b.view(np.int64).copy()
    ia = np.where(ia < 0, np.int64(-2**63) - ia, ia); ib = np.where(ib < 0, np.int64(-2**63) - ib, ib)
    d = np.abs(ia - ib).astype(np.float64)
    d[np.isnan(a) & np.isnan(b)] = 0
    d[(a == b)] = 0
    return d


def cases(n):
    h = lambda k: counter_hash(np.arange(n), k)
    n1 = np.floor(2000 * h(0)); n2 = np.floor(2000 * h(1))
    return {
        "real": [], "uniform": [h(0) * 10 - 5, h(1) * 3 + 6],
        "exponential_rate": [h(0) * 5 + 0.1], "exponential_mean": [h(0) * 5 + 0.1],
        "normal_box_muller": [h(0) * 10 - 5, 0.5 + h(1)], "normal_polar": [h(0) * 10 - 5, 0.5 + h(1)],
        "normal_ziggurat": [h(0) * 10 - 5, 0.5 + h(1)],
        "binomial": [np.floor(10 ** (6 * h(0))), h(1)],
        "poisson": [np.where(np.arange(n) % 128 == 0, 1e8 * (1 + h(0)), 10 ** (-1 + 5 * h(2)))],
        "gamma_scale": [10 ** (-1 + 2.5 * h(0)), 0.5 + h(1)], "gamma_rate": [10 ** (-1 + 2.5 * h(0)), 0.5 + h(1)],
        "beta": [0.2 + 20 * h(0), 0.2 + 20 * h(1)],
        "negative_binomial_prob": [0.5 + 100 * h(0), 0.05 + 0.9 * h(1)],
        "negative_binomial_mu": [0.5 + 100 * h(0), 100 * h(1)],
        "hypergeometric": [n1, n2, np.floor((n1 + n2) * h(2))],
        "truncated_normal": [np.zeros(n), np.ones(n),
                             np.where(np.arange(n) % 16 == 0, -np.inf, -3 + 5 * h(0)),
                             np.where(np.arange(n) % 8 == 1, np.inf, -3 + 5 * h(0) + 0.1 + 4 * h(1))],
        "cauchy": [h(0), h(1) + 0.1], "weibull": [h(0) * 3 + 0.2, h(1) + 0.1], "log_normal": [h(0), h(1) + 0.1],
        "beta_binomial_prob": [np.floor(100 * h(0)), 0.05 + 0.9 * h(1), 0.01 + 0.49 * h(2)],
        "beta_binomial_ab": [np.floor(100 * h(0)), 0.2 + 5 * h(1), 0.2 + 5 * h(2)],
        "zi_poisson": [h(0), 10 ** (-1 + 3 * h(1))],
        "zi_negative_binomial_prob": [h(0), 0.5 + 50 * h(1), 0.05 + 0.9 * h(2)],
        "zi_negative_binomial_mu": [h(0), 0.5 + 50 * h(1), 100 * h(2)],
    }


def main():
    O = Oracle("reference" if have_reference() else "port")
    print("oracle:", O.kind)
    # golden protocol, all 12 generators
    for g, name in enumerate(GEN_NAMES):
        vals, st = mb.test_xoshiro_run(generator=name)
        ov, os_ = O.xoshiro_run(g)
        assert vals == [str(int(v)) for v in ov], name
        assert [int(s, 16) for s in st] == [int(s) for s in os_], name
    print("golden protocol: 12/12 generators match")
    # seeding
    for n in (1, 2, 3, 31, 32, 33, 1000, 4097):
        r = mb.monty_rng_create(n, 42)
        got = np.frombuffer(mb.monty_rng_state(r), dtype=np.uint64).reshape(n, 4)
        assert (got == O.seed_states(42, n)).all(), n
    r = mb.monty_rng_create(1000, 42); mb.monty_rng_jump(r, 3); mb.monty_rng_long_jump(r, 2)
    exp = O.seed_states(42, 1000); O.jump(exp, 3); O.jump(exp, 2, long=True)
    assert (np.frombuffer(mb.monty_rng_state(r), dtype=np.uint64).reshape(-1, 4) == exp).all()
    print("seeding / jump / long_jump: match")
    n, ns = 4096, 37
    for rt in ("f64", "f32"):
        for name, pars in cases(n).items():
            r = mb.monty_rng_create(n, 7, real_type=rt)
            st = O.seed_states(7, n)
            t0 = time.time()
            y = mb.random._sample(r, name, ns, pars)          # (ns, n) F-order
            y = np.ascontiguousarray(y.T)
            exp = O.sample(name, st, ns, pars, rt)
            got_state = np.frombuffer(mb.monty_rng_state(r), dtype=np.uint64).reshape(n, 4)
            same = (y == exp) | (np.isnan(y) & np.isnan(exp))
            u = ulp_diff(y, exp)
            rel = np.nanmax(np.abs(y - exp) / np.maximum(np.abs(exp), 1e-300)) if not same.all() else 0.0
            state_ok = (got_state == st).all(axis=1)
            print("%-4s %-26s exact %.6f  max_ulp %-8g max_rel %.2e  streams_state_equal %.6f"
                  % (rt, name, same.mean(), u.max(), rel, state_ok.mean()))


if __name__ == "__main__":
    main()
