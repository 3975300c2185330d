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
x, pars = density_inputs_torch(name, n, 0, dev)
for _ in range(rep):
    t = density_sum_dev(name, x, pars)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); t = density_sum_dev(name, x, pars); b.record(); torch.cuda.synchronize()
print(name, n, "%.3f ms  %.1f Gobs/s" % (a.elapsed_time(b), n / a.elapsed_time(b) / 1e6), t.item())
