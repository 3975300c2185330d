#!/usr/bin/env python
"""Reads a bench.py JSON line on stdin and prints the headline and the detail table (development aid)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("%s: %.1f %s  %.3f ms/step  roofline %.3f" % (d["metric"], d["value"], d["unit"], d["ms_per_step"],
                                                  d["roofline"]["frac"]))
for k, v in (d.get("detail") or {}).items():
    print("   %-40s %8.1f Gdraws/s %8.3f ms" % (k, v["gdraws_per_s"], v["kernel_ms"]))
