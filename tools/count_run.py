#!/usr/bin/env python
"""Launches every sampler of the bench's tables exactly twice (warm-up + counted), device resident,
and writes the launch manifest; run under `ncu --metrics ...` by tools/ncu_counts.sh."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import monty_b200 as mb
from monty_b200.synthetic import sampler_params

dev = torch.device("cuda", 0)
C2 = (1 << 20, 1000)
ROWS = [("normal_ziggurat", (1 << 24, 128), False), ("normal_ziggurat", C2, False), ("normal_ziggurat", C2, True),
        ("real", C2, False), ("uniform", C2, False), ("exponential_rate", C2, False),
        ("normal_box_muller", C2, False), ("normal_polar", C2, False),
        ("binomial", (1 << 22, 1), False), ("poisson", (1 << 22, 1), False),
        ("binomial", (1 << 22, 16), False), ("poisson", (1 << 22, 16), False),
        ("gamma_scale", (1 << 21, 16), False), ("beta", (1 << 21, 16), False),
        ("negative_binomial_prob", (1 << 21, 16), False), ("hypergeometric", (1 << 21, 16), False),
        ("truncated_normal", (1 << 21, 16), False)]
only = sys.argv[1:]
manifest = []
for name, (S, D), per_stream in ROWS:
    if only and name not in only:
        continue
    rng = mb.monty_rng_create(S, 42)
    p = sampler_params(name, S, per_stream_normal=per_stream)
    pd = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in p]
    out = torch.empty((S, D), dtype=torch.float64, device=dev)
    for _ in range(2):
        rng.sample_dev(name, D, pd, out, "f64")
        rng.sync()
    manifest.append({"dist": name, "streams": S, "draws": D, "per_stream": per_stream, "launches": 2})
    rng.close(); del out
with open(os.path.join(ROOT, "gpurun_out", "count_manifest.json"), "w") as fh:
    json.dump(manifest, fh)
print("launched", len(manifest), "rows")
