#!/usr/bin/env python
"""Evaluates one fused log-likelihood a few times on device-resident config-C5 inputs (for profiling).
usage: density_run.py [name] [n_obs] [repeats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from monty_b200.distributed import density_sum_dev
from monty_b200.synthetic import density_inputs_torch
name = sys.argv[1] if len(sys.argv) > 1 else "poisson"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 250_000_000
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
if name in ("poisson", "negative_binomial_mu", "beta_binomial_prob"):
    x, pars = density_inputs_torch(name, n, 0, dev)
else:                                   # other densities: simple uniform inputs (development aid)
    g = torch.Generator(device=dev); g.manual_seed(1)
    u = lambda lo, hi: lo + (hi - lo) * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    xi = lambda hi: torch.randint(0, hi, (n,), dtype=torch.int32, device=dev, generator=g)
    x, pars = {"gamma_rate": lambda: (u(0.05, 20), [u(0.3, 30), u(0.5, 2)]),
               "gamma_scale": lambda: (u(0.05, 20), [u(0.3, 30), u(0.5, 2)]),
               "beta": lambda: (u(0.01, 0.99), [u(0.2, 20), u(0.2, 20)]),
               "zi_negative_binomial_mu": lambda: (xi(60), [u(0, 1), u(0.5, 50), u(0.1, 100)]),
               "negative_binomial_prob": lambda: (xi(60), [u(0.5, 50), u(0.05, 0.95)])}[name]()
for _ in range(rep):
    t = density_sum_dev(name, x, pars)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); t = density_sum_dev(name, x, pars); b.record(); torch.cuda.synchronize()
print(name, n, "%.3f ms  %.1f Gobs/s" % (a.elapsed_time(b), n / a.elapsed_time(b) / 1e6), t.item())
