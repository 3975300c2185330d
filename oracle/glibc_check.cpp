// TEST INFRASTRUCTURE ONLY (see oracle/Makefile).
// Host build of the PRODUCT's glibc-exact log / cos (include/monty_b200/detail/glibc_math.cuh
// compiled as plain C++ with g++ -mfma, every operation explicit) next to the
// LIVE libm of the host it runs on -- the library the reference's std::log /
// std::cos resolve to (hdr/normal_box_muller.hpp:27-34).  tests/test_glibc_math.py
// compares the two on 1e8 arguments per class; tests/test_gpu_math.py uses
// gm_libm() as the expected value of the device routines.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../include/monty_b200/detail/glibc_math.cuh"


static inline uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static inline double from_bits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
static inline uint64_t bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }

// kind 0: log of uniform (0, 1] reals as int_to_real makes them
//      1: log of random positive normal doubles (any exponent)
//      2: log of values in [0.9, 1.1] (the window around 1 and its edges)
//      3: cos of 2 pi u, u in (0, 1]  (the Box-Muller angle)
//      4: cos of values in [-8, 8]
//      5: log of random bit patterns (special cases: negative, NaN, inf, subnormal)
//      26-29: erfc on [-6, 6], [-30, 30], any bit pattern, magnitudes 2^-60 .. 32
//      22-25: log1p on (-1, 2), any bit pattern, any positive normal, magnitudes 2^-60 .. 8
//      18-21: tan of pi u, of [-30, 30], of every magnitude below 1.3e8, next to multiples of pi/2
//      33, 34: pow(negative and positive bases, integer exponents), squares and cubes of any bit pattern
//      15-17: pow(uniform, (0.05, 12)), pow(any positive normal, (-20, 20)), pow((0, 50), any bit pattern)
//      13, 14: exp on [-700, 700] and on every magnitude below 2; 30-32: [-1100, 1100], any bit pattern, subnormal results
//      9-12: lgamma of positive arguments: (0, 10), any positive normal, (0.5, 300.5), integers and halves
//      6, 7, 8: the lock-step formulation of cos on the Box-Muller angle, on
//         [-1000, 1000], on log-uniform magnitudes below 1.0e8
// Returns the number of arguments whose result differs in any bit; the first
// offender is stored in *bad.
extern "C" int64_t gm_compare(int kind, uint64_t seed, int64_t n, double* bad) {
  int64_t mism = 0;
  double first = 0;
#pragma omp parallel for reduction(+ : mism) schedule(static)
  for (int64_t c = 0; c < (n + 65535) / 65536; ++c) {
    uint64_t s = seed + 0x1234567ull * static_cast<uint64_t>(c);
    const int64_t m = (c + 1) * 65536 <= n ? 65536 : n - c * 65536;
    for (int64_t j = 0; j < m; ++j) {
      const uint64_t w = splitmix(s);
      double x, got, want;
      int sg = 0;
      const double u = static_cast<double>(w >> 11) * 0x1p-53 + 0x1p-54;
      switch (kind) {
        case 0: x = u; got = mb200::glibc::log_positive_normal(x); want = std::log(x); break;
        case 1: x = from_bits((w >> 1) % 0x7fe0000000000000ull + 0x0010000000000000ull);
                got = mb200::glibc::log_positive_normal(x); want = std::log(x); break;
        case 2: x = 0.9 + 0.2 * u; got = mb200::glibc::log(x); want = std::log(x); break;
        case 3: x = 2 * M_PI * u; got = mb200::glibc::cos_medium(x); want = std::cos(x); break;
        case 4: x = 16 * u - 8; got = mb200::glibc::cos_medium(x); want = std::cos(x); break;
        case 6: x = 2 * M_PI * u; got = mb200::glibc::cos_medium_lockstep(x); want = std::cos(x); break;
        case 7: x = 2000 * u - 1000; got = mb200::glibc::cos_medium_lockstep(x); want = std::cos(x); break;
        case 8: x = from_bits((w >> 1) % 0x4199000000000000ull);        // any magnitude below 1.0e8, log-uniform
                got = mb200::glibc::cos_medium_lockstep(x); want = std::cos(x); break;
        case 9: x = 10 * u; got = mb200::glibc::lgamma_positive(x); want = lgamma_r(x, &sg); break;        // every piece below 8
        case 10: x = from_bits((w >> 1) % 0x7fe0000000000000ull + 0x0010000000000000ull);                   // any positive normal
                 got = mb200::glibc::lgamma_positive(x); want = lgamma_r(x, &sg); break;
        case 11: x = 0.5 + 300 * u; got = mb200::glibc::lgamma_positive(x); want = lgamma_r(x, &sg); break;  // the densities' range
        case 12: x = static_cast<double>(1 + (w >> 40) % 4096) + ((w & 1) ? 0.5 : 0.0);                      // integers and halves
                 got = mb200::glibc::lgamma_positive(x); want = lgamma_r(x, &sg); break;
        case 13: x = 1400 * u - 700; got = mb200::glibc::exp(x); want = std::exp(x); break;
        case 14: x = from_bits(w & 0xbfffffffffffffffull);                                   // any exponent below 2, both signs
                 got = mb200::glibc::exp(x); want = std::exp(x); break;
        case 15: x = u; { const double yy = 0.05 + 12 * (static_cast<double>(splitmix(s) >> 11) * 0x1p-53);     // pow(u, 1 / shape)
                  got = mb200::glibc::pow(x, yy); want = std::pow(x, yy); } break;
        case 16: x = from_bits((w >> 1) % 0x7fe0000000000000ull + 0x0010000000000000ull);                      // any positive normal base
                 { const double yy = 40 * (static_cast<double>(splitmix(s) >> 11) * 0x1p-53) - 20;
                   got = mb200::glibc::pow(x, yy); want = std::pow(x, yy); } break;
        case 17: x = 50 * u; { const double yy = from_bits(splitmix(s));                                        // any exponent
                  got = mb200::glibc::pow(x, yy); want = std::pow(x, yy); } break;
        case 18: x = M_PI * u; got = mb200::glibc::tan(x); want = std::tan(x); break;                            // the cauchy sampler's argument
        case 19: x = 60 * u - 30; got = mb200::glibc::tan(x); want = std::tan(x); break;                         // both reductions below / above 25
        case 20: x = from_bits(w & 0xc19fffffffffffffull);                                                       // any magnitude below 1.3e8, both signs
                 got = mb200::glibc::tan(x); want = std::tan(x); break;
        case 21: x = (static_cast<double>(1 + (w >> 40) % 64) + 0x1p-20 * (u - 0.5)) * M_PI_2;                   // next to the poles and zeros
                 got = mb200::glibc::tan(x); want = std::tan(x); break;
        case 22: x = 3 * u - 1; got = mb200::glibc::log1p(x); want = std::log1p(x); break;                       // (-1, 2): every branch on k
        case 23: x = from_bits(w); got = mb200::glibc::log1p(x); want = std::log1p(x); break;                    // any bit pattern
        case 24: x = from_bits((w >> 1) % 0x7fe0000000000000ull + 0x0010000000000000ull);                        // any positive normal
                 got = mb200::glibc::log1p(x); want = std::log1p(x); break;
        case 25: x = from_bits((w & 0x800fffffffffffffull) | (static_cast<uint64_t>(0x3ff - 60 + (w >> 52) % 64) << 52));  // small magnitudes, both signs
                 got = mb200::glibc::log1p(x); want = std::log1p(x); break;
        case 26: x = 12 * u - 6; got = mb200::glibc::erfc(x); want = std::erfc(x); break;                        // every polynomial piece
        case 27: x = 60 * u - 30; got = mb200::glibc::erfc(x); want = std::erfc(x); break;                       // out to the underflow
        case 28: x = from_bits(w); got = mb200::glibc::erfc(x); want = std::erfc(x); break;                      // any bit pattern
        case 29: x = from_bits((w & 0x800fffffffffffffull) | (static_cast<uint64_t>(0x3ff - 60 + (w >> 52) % 65) << 52));  // magnitudes 2^-60 .. 32
                 got = mb200::glibc::erfc(x); want = std::erfc(x); break;
        case 30: x = 2200 * u - 1100; got = mb200::glibc::exp(x); want = std::exp(x); break;                     // through over- and underflow
        case 31: x = from_bits(w); got = mb200::glibc::exp(x); want = std::exp(x); break;                        // any bit pattern
        case 32: x = -(708 + 40 * u); got = mb200::glibc::exp(x); want = std::exp(x); break;                     // the subnormal results
        case 33: x = 200 * u - 100; { const double yy = static_cast<double>(static_cast<int>(splitmix(s) % 41) - 20);   // negative bases, integer exponents
                  got = mb200::glibc::pow(x, yy); want = std::pow(x, yy); } break;
        case 34: x = from_bits(w); { const double yy = (splitmix(s) & 1) ? 2.0 : 3.0;                           // squares and cubes of anything
                  got = mb200::glibc::pow(x, yy); want = std::pow(x, yy); } break;
        default: x = from_bits(w); got = mb200::glibc::log(x); want = std::log(x); break;
      }
      const bool same = bits(got) == bits(want) || (got != got && want != want);
      if (!same) {
        ++mism;
#pragma omp critical
        if (first == 0) first = x;
      }
    }
  }
  if (bad) *bad = first;
  return mism;
}
// log_positive_normal_batch<4> against four single calls of libm, n groups
extern "C" int64_t gm_compare_log_batch(uint64_t seed, int64_t n) {
  int64_t mism = 0;
#pragma omp parallel for reduction(+ : mism) schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    uint64_t s = seed + 0x9e3779b9ull * static_cast<uint64_t>(c);
    double x[4], lg[4];
    for (int i = 0; i < 4; ++i) {
      const uint64_t w = splitmix(s);
      x[i] = (w & 3) == 0 ? 0.93 + 0.14 * (static_cast<double>(w >> 11) * 0x1p-53)   // many in the window
                          : static_cast<double>(w >> 11) * 0x1p-53 + 0x1p-54;
    }
    mb200::glibc::log_positive_normal_batch<4>(x, lg);
    for (int i = 0; i < 4; ++i) mism += bits(lg[i]) != bits(std::log(x[i]));
  }
  return mism;
}
// lgamma_positive_batch<6> against six single calls of libm's lgamma, n groups
extern "C" int64_t gm_compare_lgamma_batch(uint64_t seed, int64_t n) {
  int64_t mism = 0;
#pragma omp parallel for reduction(+ : mism) schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    uint64_t s = seed + 0x9e3779b9ull * static_cast<uint64_t>(c);
    double x[6], lg[6];
    for (int i = 0; i < 6; ++i) {
      const uint64_t w = splitmix(s);
      const double u = static_cast<double>(w >> 11) * 0x1p-53 + 0x1p-54;
      switch (w & 7) {                      // every range, exact 1 and 2, huge
        case 0: x[i] = 2 * u; break;
        case 1: x[i] = 2 + 6 * u; break;
        case 2: x[i] = (w & 8) ? 1.0 : 2.0; break;
        case 3: x[i] = 0x1p58 * (1 + u); break;
        default: x[i] = 8 + 300 * u; break;
      }
    }
    mb200::glibc::lgamma_positive_batch<6>(x, lg);
    for (int i = 0; i < 6; ++i) { int sg; mism += bits(lg[i]) != bits(lgamma_r(x[i], &sg)); }
  }
  return mism;
}
// The float routines, EXHAUSTIVELY: every float bit pattern in [lo_bits, hi_bits) through
// fn 0 = logf, 1 = expf, 2 = cosf, 4 = tanf, 5 = log1pf, 3 = lgammaf (positive finite arguments only) against the
// live libm; returns the number of differing results, *bad = the first offender's bits.
extern "C" int64_t gm_compare_float(int fn, uint32_t lo_bits, uint32_t hi_bits, uint32_t* bad) {
  int64_t mism = 0;
  uint32_t first = 0;
#pragma omp parallel for reduction(+ : mism) schedule(static)
  for (int64_t u = lo_bits; u < static_cast<int64_t>(hi_bits); ++u) {
    float x, got, want;
    const uint32_t b = static_cast<uint32_t>(u);
    std::memcpy(&x, &b, 4);
    int sg = 0;
    switch (fn) {
      case 0: got = mb200::glibc::logf(x); want = ::logf(x); break;
      case 1: got = mb200::glibc::expf(x); want = ::expf(x); break;
      case 2: got = mb200::glibc::cosf(x); want = ::cosf(x); break;
      case 4: got = mb200::glibc::tanf(x); want = ::tanf(x); break;
      case 5: got = mb200::glibc::log1pf(x); want = ::log1pf(x); break;
      default: got = mb200::glibc::lgammaf_positive(x); want = ::lgammaf_r(x, &sg); break;
    }
    uint32_t gb, wb;
    std::memcpy(&gb, &got, 4); std::memcpy(&wb, &want, 4);
    if (gb != wb && !(got != got && want != want)) {
      ++mism;
#pragma omp critical
      if (first == 0) first = b;
    }
  }
  if (bad) *bad = first;
  return mism;
}
// powf(x, y) for n pseudo-random pairs: x a uniform float in (0, 1] or any positive float,
// y in (0.05, 20) or any float
extern "C" int64_t gm_compare_powf(uint64_t seed, int64_t n, int mode) {
  int64_t mism = 0;
#pragma omp parallel for reduction(+ : mism) schedule(static)
  for (int64_t c = 0; c < (n + 65535) / 65536; ++c) {
    uint64_t s = seed + 0x1234567ull * static_cast<uint64_t>(c);
    for (int j = 0; j < 65536; ++j) {
      const uint64_t w = splitmix(s);
      float x, y;
      uint32_t xb = static_cast<uint32_t>(w), yb = static_cast<uint32_t>(w >> 32);
      if (mode == 0) {
        x = static_cast<float>(xb >> 8) * 0x1p-24f + 0x1p-25f;
        y = 0.05f + 20.0f * (static_cast<float>(yb >> 8) * 0x1p-24f);
      } else if (mode == 1) {
        xb &= 0x7fffffffu;
        std::memcpy(&x, &xb, 4);
        std::memcpy(&y, &yb, 4);
      } else if (mode == 2) {                     // any x, small integer or any y: negative bases, subnormals
        std::memcpy(&x, &xb, 4);
        if (yb & 1) y = static_cast<float>(static_cast<int>((yb >> 8) % 41) - 20); else std::memcpy(&y, &yb, 4);
      } else {                                    // results next to over- and underflow: x^y with |y log2 x| in (120, 155)
        x = 0.5f + static_cast<float>(xb >> 8) * 0x1p-24f * 1.5f;
        const float t = 120.0f + 35.0f * (static_cast<float>(yb >> 8) * 0x1p-24f);
        y = ((yb & 1) ? t : -t) / std::log2(x);
      }
      const float got = mb200::glibc::powf(x, y), want = ::powf(x, y);
      uint32_t gb, wb;
      std::memcpy(&gb, &got, 4); std::memcpy(&wb, &want, 4);
      mism += gb != wb && !(got != got && want != want);
    }
  }
  return mism;
}
extern "C" double gm_pow(double x, double y) { return mb200::glibc::pow(x, y); }
extern "C" double gm_exp(double x) { return mb200::glibc::exp(x); }
extern "C" double gm_lgamma(double x) { return mb200::glibc::lgamma_positive(x); }
extern "C" double gm_log(double x) { return mb200::glibc::log(x); }
extern "C" double gm_cos(double x) { return mb200::glibc::cos_medium(x); }
extern "C" double gm_cos_lockstep(double x) { return mb200::glibc::cos_medium_lockstep(x); }

// fn 0 = log, 1 = cos, 2 = exp, 3 = sqrt, 4 = lgamma, 5 = pow(x[i], x[i ^ 1]), 6 = tan, 7 = log1p, 8 = erfc, 9 = tanf of the live libm, elementwise
extern "C" void gm_libm(int fn, const double* x, double* y, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    switch (fn) {
      case 0: y[i] = std::log(x[i]); break;
      case 1: y[i] = std::cos(x[i]); break;
      case 2: y[i] = std::exp(x[i]); break;
      case 4: { int sg; y[i] = lgamma_r(x[i], &sg); break; }
      case 5: y[i] = std::pow(x[i], x[i ^ 1]); break;       // exponent: the neighbouring element (n even)
      case 6: y[i] = std::tan(x[i]); break;
      case 7: y[i] = std::log1p(x[i]); break;
      case 8: y[i] = std::erfc(x[i]); break;
      case 9: y[i] = static_cast<double>(::tanf(static_cast<float>(x[i]))); break;
      default: y[i] = std::sqrt(x[i]); break;
    }
  }
}
// the product's routines on the host, elementwise: fn 0 = log, 1 = cos_medium
extern "C" void gm_port(int fn, const double* x, double* y, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] = fn == 0 ? mb200::glibc::log(x[i]) : mb200::glibc::cos_medium(x[i]);
}
