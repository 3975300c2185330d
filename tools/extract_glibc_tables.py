#!/usr/bin/env python
"""Extract the data tables of glibc's double-precision log() and cos().

Why.  The reference's samplers call std::log / std::cos (hdr/normal_box_muller.hpp:
27-34, hdr/exponential.hpp:26, hdr/gamma.hpp:36-50 ...), i.e. glibc's libm on the
oracle's host.  glibc's results are NOT correctly rounded (log <= 0.52 ulp, cos
<= 0.55 ulp), so "the same bits as the reference" means "the same algorithm and
the same tables as glibc", not "an accurate log".  The algorithms are published
(log: sysdeps/ieee754/dbl-64/e_log.c, Szabolcs Nagy's table-driven log from ARM
optimized-routines; cos: sysdeps/ieee754/dbl-64/s_sin.c, the IBM Accurate
Mathematical Library's do_sin / do_cos over __sincostab) and are restated in
include/monty_b200/detail/glibc_math.cuh (device) and oracle/glibc_math_port.c (CPU check);
the TABLES are data, read here from the image's own libm.so.6 (glibc 2.39,
Ubuntu 2.39-0ubuntu8.5, the library the oracle links) and written as hex-float
literals to

    include/monty_b200/detail/glibc_tables.inc

Tables:
    __log_data      ln2hi, ln2lo, poly[5], poly1[11], tab[128] {invc, logc}
    __sincostab     440 doubles: {sin hi, sin lo, cos hi, cos lo} of x_k ~ k/128
    scalar constants of s_sin.c / branred (big, hp0, hp1, mp1, mp2, pp3, pp4,
    hpinv, toint, sn3, sn5, cs2, cs4, cs6, s1..s5) -- written from their
    published values and CHECKED to be present in the library's .rodata.
    __logf_data, __exp2f_data, __sincosf_table, e_lgammaf_r.c's pool: the float
                    routines behind real_type = float
    __pow_log_data  poly[7], tab[128] {invc, logc, logctail} of pow's own log
    __exp_data      invln2N, shift, -ln2/N (hi, lo), C2..C5, tab[128] {tail, scale}
    e_lgamma_r.c    the 65 polynomial coefficients of fdlibm's lgamma (a, t, u, v,
    s, r, w families, tc, tf, tt), read from the function's constant pool.

The tables are located by signature (ln2hi followed by ln2lo; {0, 0, 1, 0,
sin(1/128)}), not by address, so another build of glibc 2.39 works too.
tests/test_glibc_math.py checks the committed file against mathematical
properties of the tables and the CPU restatement against the live libm on
1e7..1e8 arguments.
"""
import math
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBM = "/lib/x86_64-linux-gnu/libm.so.6"

LN2HI = float.fromhex("0x1.62e42fefa3800p-1")
LN2LO = float.fromhex("0x1.ef35793c76730p-45")

# s_sin.c / usncs.h / branred constants (published values)
SCALARS = {
    "big": "0x1.8p45", "toint": "0x1.8p52",
    "hp0": "0x1.921fb54442d18p+0", "hp1": "0x1.1a62633145c07p-54",
    "hpinv": "0x1.45f306dc9c883p-1",
    "mp1": "0x1.921fb58000000p+0", "mp2": "-0x1.dde973c000000p-27",
    "pp3": "-0x1.cb3b398000000p-55", "pp4": "-0x1.d747f23e32ed7p-83",
    "sn3": "-0x1.5555555555515p-3", "sn5": "0x1.11110e829872fp-7",
    "cs2": "0x1p-1", "cs4": "-0x1.5555555555535p-5", "cs6": "0x1.6c16bedd9e239p-10",
    "sin_taylor_limit": "0x1.020c49ba5e354p-3",   # 0.126
    "s1": "-0x1.5555555555555p-3", "s2": "0x1.1111111110ecep-7", "s3": "-0x1.a01a019db08b8p-13",
    "s4": "0x1.71de27b9a7ed9p-19", "s5": "-0x1.addffc2fcdf59p-26",
}


def rodata(path):
    with open(path, "rb") as f:
        blob = f.read()
    # ELF64 section headers
    shoff = struct.unpack_from("<Q", blob, 0x28)[0]
    shentsize, shnum, shstrndx = struct.unpack_from("<HHH", blob, 0x3A)
    secs = []
    for i in range(shnum):
        name, typ, flags, addr, off, size = struct.unpack_from("<IIQQQQ", blob, shoff + i * shentsize)
        secs.append((name, addr, off, size))
    stroff = secs[shstrndx][2]
    for name, addr, off, size in secs:
        end = blob.index(b"\0", stroff + name)
        if blob[stroff + name:end] == b".rodata":
            return blob, addr, off, size
    raise SystemExit("no .rodata in " + path)


def dbl(blob, off, n):
    return list(struct.unpack_from("<%dd" % n, blob, off))


def find_doubles(blob, off, size, seq):
    pat = struct.pack("<%dd" % len(seq), *seq)
    i = blob.find(pat, off, off + size)
    while i >= 0 and i % 8:
        i = blob.find(pat, i + 1, off + size)
    return i


def main(out=None):
    blob, addr, off, size = rodata(LIBM)
    # ---- __log_data: {ln2hi, ln2lo, poly[5], poly1[11], tab[128]{invc, logc}}
    # (poly first appears at +0x10; the FMA build has no tab2 in use)
    i = find_doubles(blob, off, size, [LN2HI, LN2LO])
    if i < 0:
        raise SystemExit("log table signature not found")
    log_off = i
    logd = dbl(blob, i, 2 + 5 + 11 + 256)
    ln2hi, ln2lo = logd[0], logd[1]
    A = logd[2:7]
    B = logd[7:18]
    T = logd[18:18 + 256]
    assert abs(A[0] + 0.5) < 1e-15 and B[0] == -0.5, (A[0], B[0])
    # table sanity: invc * c ~ 1 for c in [0.6875 .. 1.375], logc = log(c) to 2^-43
    for k in range(128):
        invc, logc = T[2 * k], T[2 * k + 1]
        assert 0.72 < invc < 1.46, invc
        assert abs(logc - math.log(1 / invc)) < 2 ** -40, (k, logc)
    # ---- __sincostab
    s1 = math.sin(1 / 128)
    j = off
    while True:
        j = find_doubles(blob, j, off + size - j, [0.0, 0.0, 1.0, 0.0])
        if j < 0:
            raise SystemExit("sincostab signature not found")
        if abs(dbl(blob, j + 32, 1)[0] - s1) < 1e-9:
            break
        j += 8
    sct = dbl(blob, j, 440)
    for k in range(110):
        sn, ssn, cs, ccs = sct[4 * k:4 * k + 4]
        assert abs(sn - math.sin(k / 128)) < 1e-8 and abs(cs - math.cos(k / 128)) < 1e-8, k
        assert abs(ssn) < 2 ** -50 and abs(ccs) < 2 ** -50
    # ---- scalar constants: confirm each published value occurs in .rodata
    consts = dict(SCALARS)
    missing = []
    for name, hx in consts.items():
        v = float.fromhex(hx)
        if find_doubles(blob, off, size, [v]) < 0:
            missing.append(name)
    # ---- e_lgamma_r.c (fdlibm) constants: the compiler laid them out in one pool, in
    # order of first use; anchored on a0 = Euler's gamma - 1/2 and read by offset.
    # A constant the code SUBTRACTS is stored as its magnitude (x + (-c) == x - c bit
    # for bit), so the signs of the published fdlibm values are restored here.
    a0 = float.fromhex("0x1.3c467e37db0c8p-4")
    k = find_doubles(blob, off, size, [float.fromhex("0x1.13e001a5562a7p-4"), a0])     # a2, a0
    if k < 0:
        raise SystemExit("lgamma constant pool not found")
    base = k + 8                                    # file offset of a0
    def at(rel):
        return dbl(blob, base + rel, 1)[0]
    LG = {}
    for name, rel in (("a0", 0), ("a1", 0x30), ("a2", -0x8), ("a3", 0x28), ("a4", -0x10), ("a5", 0x20), ("a6", -0x18),
                      ("a7", 0x18), ("a8", -0x20), ("a9", 0x10), ("a10", -0x28), ("a11", 0x8),
                      ("tc", -0x30), ("tc_m1", -0x38), ("tt", 0xb0),
                      ("t0", 0x58), ("t3", -0x50), ("t6", 0x48), ("t9", -0x40), ("t12", 0x38),
                      ("t2", 0x80), ("t5", -0x78), ("t8", 0x70), ("t11", -0x68), ("t14", 0x60),
                      ("t1", -0xa8), ("t4", 0xa0), ("t7", -0x98), ("t10", 0x90), ("t13", 0x88),
                      ("u1", 0xd8), ("u2", 0xd0), ("u3", 0xc8), ("u4", 0xc0), ("u5", 0xb8),
                      ("v1", 0x100), ("v2", 0xf8), ("v3", 0xf0), ("v4", 0xe8), ("v5", 0xe0),
                      ("s1", 0x138), ("s2", 0x130), ("s3", 0x128), ("s4", 0x120), ("s5", 0x118), ("s6", 0x110),
                      ("r1", 0x168), ("r2", 0x160), ("r3", 0x158), ("r4", 0x150), ("r5", 0x148), ("r6", 0x140),
                      ("w0", 0x1b8), ("w1", 0x1b0), ("w2", -0x1a8), ("w3", 0x1a0), ("w4", -0x198), ("w5", 0x190),
                      ("w6", 0x188)):
        neg = rel < 0 and name[0] in "tw" and name not in ("tc", "tc_m1")
        v = at(-rel if neg else rel)
        LG[name] = -v if neg else v
    LG["u0"] = -LG["a0"]
    LG["s0"] = -LG["a0"]
    LG["tf"] = float.fromhex("-0x1.f19b9bcc38a42p-4")
    if find_doubles(blob, off, size, [-LG["tf"]]) < 0:
        raise SystemExit("lgamma: tf not found")
    # the published values (s_lgamma / e_lgamma_r.c): spot checks of the layout
    assert abs(LG["tc"] - 1.46163214496836224576) < 1e-18 and abs(LG["tc_m1"] - 0.46163214496836225) < 1e-17
    assert abs(LG["w0"] - 0.41893853320467274178) < 1e-17 and abs(LG["w1"] - 1 / 12) < 1e-14 and LG["w2"] < 0
    assert abs(LG["t0"] - 0.483836122723810047042) < 1e-16 and LG["t1"] < 0 and LG["t3"] < 0 and LG["t13"] < 0
    assert abs(LG["a1"] - 0.322467033424113591611) < 1e-16 and abs(LG["u1"] - 0.632827064025093366517) < 1e-16
    assert abs(LG["s1"] - 0.214982415960608852501) < 1e-16 and abs(LG["r1"] - 1.39200533467621045958) < 1e-15
    # ---- __exp_data (e_exp.c, Szabolcs Nagy's exp, N = 128): invln2N, shift, -ln2hi/N,
    # -ln2lo/N, C2..C5, [pad], tab[2 N] = {tail, scale bits} -- anchored on InvLn2N
    invln2n = float.fromhex("0x1.71547652b82fep7")
    e = find_doubles(blob, off, size, [invln2n, float.fromhex("0x1.8p52")])
    if e < 0:
        raise SystemExit("exp data signature not found")
    EX = dict(zip(("invln2n", "shift", "negln2hin", "negln2lon", "c2", "c3", "c4", "c5"), dbl(blob, e, 8)))
    assert abs(EX["c2"] - 0.5) < 1e-9 and abs(EX["c3"] - 1 / 6) < 1e-9 and abs(EX["negln2hin"] + math.log(2) / 128) < 1e-12
    exptab = list(struct.unpack_from("<256Q", blob, e + 0xb0))
    assert exptab[0] == 0 and exptab[1] == 0x3ff0000000000000
    for i in range(128):          # scale bits + (i << 45) is 2^(i/128); tail is its rounding error
        sc = struct.unpack("<d", struct.pack("<Q", exptab[2 * i + 1] + (i << 45)))[0]
        assert abs(sc - 2.0 ** (i / 128)) < 1e-15, i
    # ---- __pow_log_data (e_pow.c): ln2hi, ln2lo (the same values as log's), poly[7],
    # tab[128] = {invc, pad, logc, logctail}: the SECOND occurrence of {ln2hi, ln2lo}
    pw = find_doubles(blob, log_off + 16, off + size - (log_off + 16), [LN2HI, LN2LO])
    if pw < 0:
        raise SystemExit("pow log data signature not found")
    PW = dbl(blob, pw + 16, 7)
    assert PW[0] == -0.5 and abs(PW[1] + 2 / 3.0) < 1e-12       # A[0] = -1/2, A[1] = (1/3) * -2 (scaled by powers of -2)
    powtab = dbl(blob, pw + 0x48, 128 * 4)
    for k in range(128):
        invc, pad, logc, tail = powtab[4 * k:4 * k + 4]
        assert pad == 0.0 and 0.7 < invc < 1.45 and abs(logc + tail - math.log(1 / invc)) < 1e-15, k
    # ---- float routines (real_type = float): e_logf.c, e_expf.c, s_sincosf.h (Szabolcs
    # Nagy; double arithmetic inside), e_lgammaf_r.c (fdlibm, float arithmetic)
    lf = find_doubles(blob, off, size, [float.fromhex("0x1.62e42fefa39efp-1"), float.fromhex("-0x1.00ea348b88334p-2")])
    if lf < 0:
        raise SystemExit("logf data signature not found")
    LOGF = dict(ln2=dbl(blob, lf, 1)[0], poly=dbl(blob, lf + 8, 3), tab=dbl(blob, lf - 256, 32))
    for k in range(16):
        assert abs(LOGF["tab"][2 * k + 1] - math.log(1 / LOGF["tab"][2 * k])) < 1e-15, k
    ef = find_doubles(blob, off, size, [float.fromhex("0x1.8p52"), float.fromhex("0x1.71547652b82fep+5")])
    if ef < 0:
        raise SystemExit("expf data signature not found")
    EXPF = dict(shift=dbl(blob, ef, 1)[0], invln2n=dbl(blob, ef + 8, 1)[0], poly=dbl(blob, ef + 16, 3),
                tab=list(struct.unpack_from("<32Q", blob, ef - 0x120)))
    for k in range(32):
        sc = struct.unpack("<d", struct.pack("<Q", EXPF["tab"][k] + (k << 47)))[0]
        assert abs(sc - 2.0 ** (k / 32)) < 1e-15, k
    sf = find_doubles(blob, off, size, [float.fromhex("0x1.45f306dc9c883p+23"), float.fromhex("0x1.921fb54442d18p+0")])
    if sf < 0:
        raise SystemExit("sincosf table signature not found")
    SINCOSF = dbl(blob, sf - 0x20, 28)          # two entries: {sign[4], hpi_inv, hpi, c0, c1, s1, c2, s2, c3, s3, c4}
    assert SINCOSF[:4] == [1.0, -1.0, -1.0, 1.0] and SINCOSF[6] == 1.0 and SINCOSF[14 + 6] == -1.0
    # __powf_log2_data {tab[16] {invc, logc}, poly[5]} anchored on the published polynomial;
    # powf's exp2 half uses __exp2f_data.tab with shift_scaled and poly[3] (the words before `shift`)
    pf = find_doubles(blob, off, size, [float.fromhex("0x1.27616c9496e0bp-2"), float.fromhex("-0x1.71969a075c67ap-2")])
    if pf < 0:
        raise SystemExit("powf log2 data signature not found")
    POWF = dict(poly=dbl(blob, pf, 5), tab=dbl(blob, pf - 256, 32), shift_scaled=dbl(blob, ef - 0x20, 1)[0],
                exp2_poly=dbl(blob, ef - 0x18, 3))
    for k in range(16):
        assert abs(POWF["tab"][2 * k + 1] - math.log2(1 / POWF["tab"][2 * k])) < 1e-15, k
    assert POWF["shift_scaled"] == float.fromhex("0x1.8p47") and abs(POWF["exp2_poly"][2] - math.log(2)) < 1e-9
    a2f, a0f = struct.unpack("<f", struct.pack("<f", 0.0673523023724556))[0], struct.unpack("<f", struct.pack("<f", 0.07721566408872604))[0]
    pat = struct.pack("<2f", a2f, a0f)
    gf = blob.find(pat, off, off + size)
    if gf < 0:
        raise SystemExit("lgammaf constant pool not found")
    gf += 4                                         # file offset of a0 (float)
    def atf(idx):
        return struct.unpack_from("<f", blob, gf + 4 * idx)[0]
    LGF = {}
    for name, idx, neg in (("a0", 0, 0), ("a1", 6, 0), ("a2", -1, 0), ("a3", 5, 0), ("a4", -2, 0), ("a5", 4, 0), ("a6", -3, 0),
                           ("a7", 3, 0), ("a8", -4, 0), ("a9", 2, 0), ("a10", -5, 0), ("a11", 1, 0), ("tc", -6, 0), ("tc_m1", -7, 0),
                           ("t12", 7, 0), ("t9", 8, 1), ("t6", 9, 0), ("t3", 10, 1), ("t0", 11, 0),
                           ("t14", 12, 0), ("t11", 13, 1), ("t8", 14, 0), ("t5", 15, 1), ("t2", 16, 0),
                           ("t13", 17, 0), ("t10", 18, 0), ("t7", 19, 1), ("t4", 20, 0), ("t1", 21, 1), ("tt", 22, 0), ("tf", 23, 1),
                           ("u5", 24, 0), ("u4", 25, 0), ("u3", 26, 0), ("u2", 27, 0), ("u1", 28, 0),
                           ("v5", 29, 0), ("v4", 30, 0), ("v3", 31, 0), ("v2", 32, 0), ("v1", 33, 0),
                           ("s6", 34, 0), ("s5", 35, 0), ("s4", 36, 0), ("s3", 37, 0), ("s2", 38, 0), ("s1", 39, 0),
                           ("r6", 40, 0), ("r5", 41, 0), ("r4", 42, 0), ("r3", 43, 0), ("r2", 44, 0), ("r1", 45, 0),
                           ("w6", 46, 0), ("w5", 47, 0), ("w4", 48, 1), ("w3", 49, 0), ("w2", 50, 1), ("w1", 51, 0), ("w0", 52, 0)):
        LGF[name] = -atf(idx) if neg else atf(idx)
    LGF["u0"] = -LGF["a0"]
    LGF["s0"] = -LGF["a0"]
    assert abs(LGF["tc"] - 1.46163214) < 1e-6 and abs(LGF["w0"] - 0.41893853) < 1e-6 and abs(LGF["w1"] - 1 / 12) < 1e-6
    assert LGF["w6"] < 0 and LGF["t13"] < 0 and abs(LGF["r1"] - 1.3920053) < 1e-6 and abs(LGF["u1"] - 0.63282706) < 1e-6
    # ---- s_tan.c (IBM Accurate Mathematical Library, slow paths removed in 2.28+): the
    # 186-row table xfg {x_i, tan x_i, cot x_i, (unused)} anchored on its published first row,
    # and the scalars of utan.h, each confirmed present in .rodata
    tn = find_doubles(blob, off, size, [float.fromhex("0x1.000001e519d60p-4"), float.fromhex("0x1.0055796c4e240p-4")])
    if tn < 0:
        raise SystemExit("tan table (xfg) signature not found")
    xfg = dbl(blob, tn, 186 * 4)
    for k in range(186):
        xi, fi, gi = xfg[4 * k:4 * k + 3]
        assert abs(xi - (k + 16) / 256) < 2 ** -25 and abs(fi - math.tan(xi)) < 1e-15 and abs(gi * fi - 1) < 1e-15, k
    TAN = {
        "g1": "0x1.b096cp-27", "g2": "0x1.f212dp-5", "g3": "0x1.92f1ap-1", "g4": "0x1.9p+4", "g5": "0x1.7d784p+26",
        "d3": "0x1.5555555555555p-2", "d5": "0x1.11111111107c6p-3", "d7": "0x1.ba1ba1cdb8745p-5",
        "d9": "0x1.664ed49cfc666p-6", "d11": "0x1.2385a3cf2e4eap-7",
        "e0": "0x1.5555555554dbdp-2", "e1": "0x1.11112e0a6b45fp-3", "mp3": "-0x1.cb3b399d747f2p-55",
    }
    for name, hx in TAN.items():
        if find_doubles(blob, off, size, [float.fromhex(hx)]) < 0:
            missing.append("tan_" + name)
    # ---- k_tanf.c (fdlibm, float arithmetic): pio4, pio4lo, 2^-13 and T[0..12] in the pool's order
    # T12 T10 T8 T6 T4 T2 T11 T9 T7 T5 T3 T1 T0, anchored on {pio4, pio4lo}
    tf_ = blob.find(struct.pack("<2I", 0x3f490fda, 0x33222168), off, off + size)
    if tf_ < 0:
        raise SystemExit("k_tanf constant pool not found")
    tw = struct.unpack_from("<16f", blob, tf_)
    order = (12, 10, 8, 6, 4, 2, 11, 9, 7, 5, 3, 1, 0)
    TANF_T = [0.0] * 13
    for pos, idx in enumerate(order):
        TANF_T[idx] = tw[3 + pos]
    assert tw[2] == 2.0 ** -13 and abs(TANF_T[0] - 1 / 3) < 1e-7 and abs(TANF_T[1] - 2 / 15) < 1e-7 and abs(TANF_T[2] - 17 / 315) < 1e-7
    TANF_PIO4, TANF_PIO4LO = tw[0], tw[1]
    # ---- s_log1p.c (fdlibm; FMA ifunc variant): published constants, confirmed present
    LOG1P = {
        "ln2_hi": "0x1.62e42fee00000p-1", "ln2_lo": "0x1.a39ef35793c76p-33", "two_thirds": "0x1.5555555555555p-1",
        "lp1": "0x1.5555555555593p-1", "lp2": "0x1.999999997fa04p-2", "lp3": "0x1.2492494229359p-2",
        "lp4": "0x1.c71c51d8e78afp-3", "lp5": "0x1.7466496cb03dep-3", "lp6": "0x1.39a09d078c69fp-3",
        "lp7": "0x1.2f112df3e5244p-3",
    }
    for name, hx in LOG1P.items():
        if find_doubles(blob, off, size, [float.fromhex(hx)]) < 0:
            missing.append("log1p_" + name)
    # ---- s_erf.c, erfc half (fdlibm with glibc's split polynomials; not an ifunc, no FMA).
    # The compiler pooled the 58 constants of __erfc in order of first use and stored
    # the magnitude of each coefficient the code subtracts; glibc_math.cuh's erfc indexes
    # the pool the way the routine does.  Anchored on erx (index 23) and 1 - erx (57).
    erx = float.fromhex("0x1.b0ac160000000p-1")
    ex = find_doubles(blob, off, size, [erx])
    while ex >= 0 and dbl(blob, ex + 8 * 34, 1)[0] != 1.0 - erx:
        ex = find_doubles(blob, ex + 8, off + size - ex - 8, [erx])
    if ex < 0:
        raise SystemExit("erfc constant pool not found")
    ERFC = dbl(blob, ex - 8 * 23, 58)
    assert ERFC[56] == 0.5625 and ERFC[25] == 1e-300 and abs(ERFC[3] - 0.1283791670955126) < 1e-15, ERFC
    if out is None:
        return dict(ln2hi=ln2hi, lgamma=LG, ln2lo=ln2lo, A=A, B=B, T=T, sincos=sct, missing=missing,
                    log_addr=addr + (log_off - off), sincos_addr=addr + (j - off))
    if missing:
        raise SystemExit("constants not found in libm .rodata: %s" % missing)
    fh = lambda v: float(v).hex()  # noqa: E731
    L = []
    L.append("/* GENERATED by tools/extract_glibc_tables.py -- do not edit.")
    L.append(" * Numeric data only (hex-float literals, bit-exact): the tables of glibc 2.39's")
    L.append(" * double-precision log (__log_data) and sin/cos (__sincostab), read from the")
    L.append(" * image's libm.so.6.  See the tool's docstring for why parity needs them. */")
    L.append("#define GLIBC_LOG_LN2HI %s" % fh(ln2hi))
    L.append("#define GLIBC_LOG_LN2LO %s" % fh(ln2lo))
    L.append("#define GLIBC_LOG_POLY { %s }" % ", ".join(fh(v) for v in A))
    L.append("#define GLIBC_LOG_POLY1 { %s }" % ", ".join(fh(v) for v in B))
    L.append("/* tab[128]: {invc, logc} */")
    L.append("#define GLIBC_LOG_TAB { \\")
    for k in range(0, 128, 2):
        L.append("  " + ", ".join("{%s, %s}" % (fh(T[2 * m]), fh(T[2 * m + 1])) for m in (k, k + 1)) + ", \\")
    L[-1] = L[-1][:-3] + " }"
    L.append("/* __sincostab[440]: {sin hi, sin lo, cos hi, cos lo} per k = round(128 |x|) */")
    L.append("#define GLIBC_SINCOS_TAB { \\")
    for k in range(110):
        L.append("  " + ", ".join(fh(v) for v in sct[4 * k:4 * k + 4]) + ", \\")
    L[-1] = L[-1][:-3] + " }"
    # the same rows arranged for the lock-step formulation of do_sin / do_cos
    # (glibc_math.cuh: sincos_core): row [kind][k] = {P, p, Q, q} with
    # sin (kind 1): {sn, ssn, cs, ccs};  cos (kind 0): {cs, ccs, -sn, -ssn}
    L.append("/* [2][110] rows {P, p, Q, q}: kind 0 (cos) = {cs, ccs, -sn, -ssn}, kind 1 (sin) = {sn, ssn, cs, ccs} */")
    L.append("#define GLIBC_SINCOS_ROWS { \\")
    for kind in (0, 1):
        for k in range(110):
            sn, ssn, cs, ccs = sct[4 * k:4 * k + 4]
            row = (sn, ssn, cs, ccs) if kind else (cs, ccs, -sn, -ssn)
            L.append("  {" + ", ".join(fh(v) for v in row) + "}, \\")
    L[-1] = L[-1][:-3] + " }"
    for name, hx in consts.items():
        L.append("#define GLIBC_SC_%s %s" % (name.upper(), fh(float.fromhex(hx))))
    L.append("/* __exp_data: scalars, then tab[128] = {tail bits, scale bits - (i << 45)} */")
    for name, v in EX.items():
        L.append("#define GLIBC_EXP_%s %s" % (name.upper(), fh(v)))
    L.append("#define GLIBC_EXP_TAB { \\")
    for i in range(0, 256, 4):
        L.append("  " + ", ".join("0x%016xull" % v for v in exptab[i:i + 4]) + ", \\")
    L[-1] = L[-1][:-3] + " }"
    L.append("/* __pow_log_data: poly[7], tab[128] = {invc, logc, logctail} (the pad word dropped) */")
    L.append("#define GLIBC_POW_POLY { %s }" % ", ".join(fh(v) for v in PW))
    L.append("#define GLIBC_POW_TAB { \\")
    for k in range(128):
        L.append("  {%s, %s, %s}, \\" % (fh(powtab[4 * k]), fh(powtab[4 * k + 2]), fh(powtab[4 * k + 3])))
    L[-1] = L[-1][:-3] + " }"
    fhf = lambda v: float(v).hex() + "f"  # noqa: E731  (exact: a float widened to double prints exactly)
    L.append("/* __logf_data: tab[16] = {invc, logc}, ln2, poly[3] (double arithmetic) */")
    L.append("#define GLIBC_LOGF_LN2 %s" % fh(LOGF["ln2"]))
    L.append("#define GLIBC_LOGF_POLY { %s }" % ", ".join(fh(v) for v in LOGF["poly"]))
    L.append("#define GLIBC_LOGF_TAB { %s }" % ", ".join("{%s, %s}" % (fh(LOGF["tab"][2 * k]), fh(LOGF["tab"][2 * k + 1])) for k in range(16)))
    L.append("/* __exp2f_data as expf uses it: tab[32] (scale bits - (i << 47)), shift, invln2 * 32, poly_scaled[3] */")
    L.append("#define GLIBC_EXPF_SHIFT %s" % fh(EXPF["shift"]))
    L.append("#define GLIBC_EXPF_INVLN2N %s" % fh(EXPF["invln2n"]))
    L.append("#define GLIBC_EXPF_POLY { %s }" % ", ".join(fh(v) for v in EXPF["poly"]))
    L.append("#define GLIBC_EXPF_TAB { %s }" % ", ".join("0x%016xull" % v for v in EXPF["tab"]))
    L.append("/* powf: __powf_log2_data {tab[16] {invc, log2 c}, poly[5]}, exp2f's shift_scaled and poly[3] */")
    L.append("#define GLIBC_POWF_POLY { %s }" % ", ".join(fh(v) for v in POWF["poly"]))
    L.append("#define GLIBC_POWF_TAB { %s }" % ", ".join("{%s, %s}" % (fh(POWF["tab"][2 * k]), fh(POWF["tab"][2 * k + 1])) for k in range(16)))
    L.append("#define GLIBC_EXP2F_SHIFT_SCALED %s" % fh(POWF["shift_scaled"]))
    L.append("#define GLIBC_EXP2F_POLY { %s }" % ", ".join(fh(v) for v in POWF["exp2_poly"]))
    L.append("/* __sincosf_table[2]: {sign[4], hpi_inv, hpi, c0, c1, s1, c2, s2, c3, s3, c4} */")
    L.append("#define GLIBC_SINCOSF_TAB { %s }" % ", ".join(fh(v) for v in SINCOSF))
    L.append("/* s_tan.c: utan.h scalars and xfg[186] = {x_i, tan x_i, cot x_i}, x_i ~ (i + 16) / 256 */")
    for name, hx in TAN.items():
        L.append("#define GLIBC_TAN_%s %s" % (name.upper(), fh(float.fromhex(hx))))
    L.append("#define GLIBC_TAN_XFG { \\")
    for k in range(186):
        L.append("  {%s, %s, %s, 0.0}, \\" % tuple(fh(v) for v in xfg[4 * k:4 * k + 3]))
    L[-1] = L[-1][:-3] + " }"
    L.append("/* k_tanf.c (fdlibm, float): T[13], pio4, pio4lo */")
    L.append("#define GLIBC_TANF_T { %s }" % ", ".join(fhf(v) for v in TANF_T))
    L.append("#define GLIBC_TANF_PIO4 %s" % fhf(TANF_PIO4))
    L.append("#define GLIBC_TANF_PIO4LO %s" % fhf(TANF_PIO4LO))
    L.append("/* s_log1p.c (fdlibm) */")
    for name, hx in LOG1P.items():
        L.append("#define GLIBC_LOG1P_%s %s" % (name.upper(), fh(float.fromhex(hx))))
    L.append("/* s_erf.c: __erfc's constant pool in the library's order (see the tool) */")
    L.append("#define GLIBC_ERFC_POOL { \\")
    for k in range(0, 58, 4):
        L.append("  " + ", ".join(fh(v) for v in ERFC[k:k + 4]) + ", \\")
    L[-1] = L[-1][:-3] + " }"
    L.append("/* e_lgammaf_r.c (fdlibm, float) constants, signs as published */")
    for name in sorted(LGF, key=lambda n: (n.rstrip("0123456789_m"), int("".join(c for c in n if c.isdigit()) or 0))):
        L.append("#define GLIBC_LGF_%s %s" % (name.upper(), fhf(LGF[name])))
    L.append("/* e_lgamma_r.c (fdlibm) constants, signs as published */")
    for name in sorted(LG, key=lambda n: (n.rstrip("0123456789_m"), int("".join(c for c in n if c.isdigit()) or 0))):
        L.append("#define GLIBC_LG_%s %s" % (name.upper(), fh(LG[name])))
    with open(out, "w") as f:
        f.write("\n".join(L) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--probe":
        r = main(None)
        print("log_data at 0x%x, sincostab at 0x%x" % (r["log_addr"], r["sincos_addr"]))
        print("missing scalar constants:", r["missing"])
    else:
        main(os.path.join(ROOT, "include", "monty_b200", "detail", "glibc_tables.inc"))
