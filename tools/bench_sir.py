#!/usr/bin/env python
"""Throughput of the fused SIR particle filter (examples/sir_filter.cu) next to
the same model over the reference's C++ samplers on one host core
(oracle/ref_harness.cpp: ref_sir_filter).  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import monty_b200 as mb  # noqa: E402
import monty_b200.examples as ex  # noqa: E402
from oracle_lib import Oracle, have_reference  # noqa: E402

SIR = dict(N=1000.0, I0=10.0, beta=0.2, gamma=0.1, exp_noise=1e6)
TIMES = np.arange(4.0, 81.0, 4.0)
INC = np.array([1, 3, 5, 7, 11, 14, 19, 22, 25, 24, 21, 18, 14, 10, 8, 5, 4, 2, 1, 1], dtype=np.float64)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(TIMES[-1])                       # model steps per particle
system = mb.monty_rng_create(n, 42)
filt = mb.monty_rng_create(1, 43)
ex.sir_filter(system, filt, SIR, TIMES, INC)                 # warm-up
t = []
for _ in range(5):
    t0 = time.perf_counter()
    ll = ex.sir_filter(system, filt, SIR, TIMES, INC)
    t.append(time.perf_counter() - t0)
gpu = min(t)
line = {"workload": "SIR bootstrap particle filter, %d particles x %d steps, %d observations" % (n, steps, len(TIMES)),
        "gpu_s": gpu, "gpu_particle_steps_per_s": n * steps / gpu, "log_likelihood": ll}
if have_reference():
    orc = Oracle("reference")
    n_cpu = min(n, 1 << 15)
    st = orc.seed_states(42, n_cpu)
    fs = orc.seed_states(43, 1)
    t0 = time.perf_counter()
    orc.sir_filter(st, fs, SIR, TIMES, INC)
    cpu = time.perf_counter() - t0
    line.update({"cpu_particles": n_cpu, "cpu_s": cpu, "cpu_particle_steps_per_s": n_cpu * steps / cpu,
                 "cpu": "reference samplers, 1 core", "ratio": (n * steps / gpu) / (n_cpu * steps / cpu)})
print(json.dumps(line))
