#!/usr/bin/env python
"""Device-resident timing of a list of samplers in one process (development aid).
usage: quick_bench.py [dist[:streams[:draws]] ...]   default shapes: 2^20 x 1000 (simple), C3/C4 shapes otherwise"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import monty_b200 as mb
import bench

dev = torch.device("cuda", 0)
specs = sys.argv[1:] or ["real", "exponential_rate", "normal_box_muller", "normal_polar", "normal_ziggurat"]
for spec in specs:
    parts = spec.split(":")
    name = parts[0]
    simple = name in ("real", "uniform", "exponential_rate", "normal_box_muller", "normal_polar", "normal_ziggurat")
    S = int(parts[1]) if len(parts) > 1 else (1 << 20 if simple else 1 << 21)
    D = int(parts[2]) if len(parts) > 2 else (1000 if simple else 16)
    rng = mb.monty_rng_create(S, 42)
    p = bench.synthetic_params(name, S)
    pd = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in p]
    out = torch.empty((S, D), dtype=torch.float64, device=dev)
    tm, km, _ = bench.time_device(mb, torch, rng, name, "f64", D, pd, out, 10, 3, lambda: None)
    k = float(np.mean(km))
    print("%-26s %8d x %-5d %8.1f Gdraws/s  %.3f ms" % (name, S, D, S * D / (k * 1e-3) / 1e9, k), flush=True)
    rng.close(); del out
