#!/usr/bin/env python
"""Generates monty_b200/csrc/log_table.cuh: the 128-bin table of the samplers' double log
(csrc/samplers.cuh: fast_log).  Needs mpmath."""
import struct
import mpmath as mp

mp.mp.prec = 200
OFF = 0x3fe5f00000000000          # 1.0 sits in the middle of bin 80


def asd(i):
    return struct.unpack("<d", struct.pack("<Q", i))[0]


rows = []
for i in range(128):
    zl, zh = asd(OFF + (i << 45)), asd(OFF + ((i + 1) << 45))
    invc = 1.0 if i == 80 else float(1 / ((mp.mpf(zl) + mp.mpf(zh)) / 2))
    lc = -mp.log(mp.mpf(invc))
    m = struct.unpack("<Q", struct.pack("<d", float(lc)))[0] & ~((1 << 11) - 1)
    lch = struct.unpack("<d", struct.pack("<Q", m))[0]      # 42 bits: k * Ln2hi + lch is exact
    rows.append((invc, lch, float(lc - mp.mpf(lch))))
# {1/c, log c high} as 16-byte rows, the low parts as a float array (24 bits of a
# 2^-44 quantity are plenty): a warp's 32 look-ups touch 16 + 4 cache lines
print("static __device__ const double2 g_log_tab[128] = {")
for invc, lch, lcl in rows:
    print("  {%s, %s}," % (invc.hex(), lch.hex()))
print("};")
print("static __device__ const float g_log_lo[128] = {")
for i in range(0, 128, 4):
    print("  " + " ".join("%sf," % float(struct.unpack("<f", struct.pack("<f", r[2]))[0]).hex() for r in rows[i:i + 4]))
print("};")
