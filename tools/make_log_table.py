#!/usr/bin/env python
"""Generates monty_b200/csrc/log_table.cuh: the 32-row table and the polynomial of the
samplers' double log (csrc/samplers.cuh: fast_log).  Needs mpmath.

x = 2^k z with z in [OFF, 2 OFF); the top 5 mantissa bits of (x - OFF) pick a row
{1/c (a float: only its closeness to the bin centre matters), log c low (float),
log c high (42 bits, so that k * ln2_high + it is exact)}; 1.0 is the centre of
row 20, whose c is exactly 1.  log1p(r) - r = r^2 P(r), P of degree 7 interpolated
at Chebyshev nodes of [-2^-6, 2^-6] (the range of r = z / c - 1)."""
import struct
import mpmath as mp

mp.mp.prec = 300
B, N0 = 5, 20
SH = 52 - B
OFF = 0x3ff0000000000000 - ((2 * N0 + 1) << (SH - 1))


def asd(i):
    return struct.unpack("<d", struct.pack("<Q", i))[0]


def f32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


rows = []
for i in range(1 << B):
    zl, zh = asd(OFF + (i << SH)), asd(OFF + ((i + 1) << SH))
    invc = 1.0 if i == N0 else f32(float(1 / ((mp.mpf(zl) + mp.mpf(zh)) / 2)))
    lc = -mp.log(mp.mpf(invc))
    lch = asd(struct.unpack("<Q", struct.pack("<d", float(lc)))[0] & ~((1 << 11) - 1))
    rows.append((invc, f32(float(lc - mp.mpf(lch))), lch))

D = 7
R = mp.mpf(2) ** -6 * mp.mpf("1.01")
xs = [R * mp.cos(mp.pi * (2 * j + 1) / (2 * (D + 1))) for j in range(D + 1)]
A, b = mp.matrix(D + 1, D + 1), mp.matrix(D + 1, 1)
for a, x in enumerate(xs):
    for j in range(D + 1):
        A[a, j] = x ** j
    b[a] = (mp.log1p(x) - x) / (x * x)
co = [float(c) for c in mp.lu_solve(A, b)]
assert abs(co[0] + 0.5) < 1e-15
print("// offset 0x%016x; P(r) = -1/2 + r c1 + ... + r^7 c7" % OFF)
print("constexpr unsigned long long kLogOff = 0x%016xull;" % OFF)
print("__constant__ double c_log_poly[7] = {%s};" % ", ".join(c.hex() for c in co[1:]))
print("struct __align__(16) LogRow { float invc, lo; double hi; };")
print("static __device__ const LogRow g_log_tab[%d] = {" % (1 << B))
for invc, lcl, lch in rows:
    print("  {%sf, %sf, %s}," % (invc.hex(), lcl.hex(), lch.hex()))
print("};")
