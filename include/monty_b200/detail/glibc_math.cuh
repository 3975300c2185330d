// glibc_math.cuh -- double-precision log() and cos() that return glibc 2.39's BITS.
//
// The reference's samplers and densities call std::log / std::cos
// (hdr/normal_box_muller.hpp:27-34, hdr/exponential.hpp:26, hdr/gamma.hpp:36-50,
// hdr/density.hpp), which on the oracle's host is glibc 2.39's libm.  Those
// routines are not correctly rounded (log <= 0.52 ulp, cos <= 0.55 ulp), so an
// accurate device log / cos differs from the reference in the last bit of ~3% of
// the results -- and Marsaglia-Tsang's (1 + c x)^3 turns a 1-ulp normal into a
// 16-ulp gamma.  Parity to the north-star's letter needs the same ALGORITHM on
// the same TABLES, operation for operation:
//
//   log : sysdeps/ieee754/dbl-64/e_log.c (Szabolcs Nagy's table-driven log):
//         128-entry {1/c, log c} table, r = z/c - 1 by one FMA, degree-5
//         polynomial; a separate degree-11 polynomial with a split r for
//         x in [1 - 2^-4, 1 + 0x1.09p-4).
//   cos : sysdeps/ieee754/dbl-64/s_sin.c (IBM Accurate Mathematical Library):
//         |x| < 0.855 do_cos; < 2.426 do_sin(pi/2 - |x|); < 1.05e8 reduce_sincos
//         (x - n pi/2 in four pieces) + do_cos / do_sin by quadrant; do_sin /
//         do_cos = 440-entry {sin, cos}(k/128) double-double table + short
//         polynomials, TAYLOR_SIN below 0.126.
//
// x86-64 glibc selects, by ifunc, the build of these files compiled with
// -mfma -mavx2 on every CPU that has FMA (all B200 hosts; the build container
// too), and GCC contracts a*b+c freely in that build.  WHICH products are fused
// is therefore part of the function's definition; the sequences below follow
// the instruction stream of that build (__log_fma / __cos_fma of Ubuntu glibc
// 2.39-0ubuntu8.5), each fused operation written as an explicit FMA.  Tables:
// glibc_tables.inc (tools/extract_glibc_tables.py).
//
// The file also compiles as plain host C++ (g++ -mfma): tests/test_glibc_math.py
// builds it that way and compares every routine with the live libm on 1e8
// arguments, so the source that is checked on the CPU is the source that runs on
// the device (tests/test_gpu_math.py repeats the comparison on the device).
#pragma once
#include <cstdint>
#include <cstring>

#include "glibc_tables.inc"

#if defined(__CUDACC__)
#define GM_FN __host__ __device__ __forceinline__
// cold paths (results next to over- / underflow): out of line, so the hot callers do not grow
#define GM_COLD static __host__ __device__ __noinline__
#else
#define GM_FN static inline
#define GM_COLD static
#endif

namespace mb200 {
namespace glibc {

struct alignas(16) LogRow { double invc, logc; };
struct SinCosRow { double sn, ssn, cs, ccs; };
struct alignas(32) SinCosRow2 { double P, p, Q, q; };    // lock-step rows, see sincos_core
struct PowLogRow { double invc, logc, logctail; };
struct alignas(32) TanRow { double x, f, g, pad; };            // xfg of s_tan.c: x_i, tan x_i, cot x_i

#if defined(__CUDACC__)
static __device__ const LogRow g_log_tab[128] = GLIBC_LOG_TAB;
static __device__ __align__(32) const double g_sincos_tab[440] = GLIBC_SINCOS_TAB;
static __device__ const SinCosRow2 g_sincos_rows[220] = GLIBC_SINCOS_ROWS;
static __device__ __align__(16) const unsigned long long g_exp_tab[256] = GLIBC_EXP_TAB;
static __device__ const PowLogRow g_pow_tab[128] = GLIBC_POW_TAB;
static __device__ const TanRow g_tan_tab[186] = GLIBC_TAN_XFG;
static __device__ const LogRow g_logf_tab[16] = GLIBC_LOGF_TAB;
static __device__ __align__(16) const unsigned long long g_expf_tab[32] = GLIBC_EXPF_TAB;
static __constant__ double c_sincosf_tab[28] = GLIBC_SINCOSF_TAB;
static __constant__ float c_tanf_t[13] = GLIBC_TANF_T;
static __device__ const LogRow g_powf_tab[16] = GLIBC_POWF_TAB;
static __constant__ double c_powf_poly[5] = GLIBC_POWF_POLY;
static __constant__ double c_exp2f_poly[3] = GLIBC_EXP2F_POLY;
static __constant__ double c_logf_poly[3] = GLIBC_LOGF_POLY;
static __constant__ double c_expf_poly[3] = GLIBC_EXPF_POLY;
static __constant__ double c_pow_poly[7] = GLIBC_POW_POLY;
static __constant__ double c_erfc_pool[58] = GLIBC_ERFC_POOL;
static __constant__ double c_log_poly[5] = GLIBC_LOG_POLY;
static __constant__ double c_log_poly1[11] = GLIBC_LOG_POLY1;
#endif
#if !defined(__CUDA_ARCH__)
static const LogRow h_log_tab[128] = GLIBC_LOG_TAB;
static const double h_sincos_tab[440] = GLIBC_SINCOS_TAB;
static const SinCosRow2 h_sincos_rows[220] = GLIBC_SINCOS_ROWS;
static const unsigned long long h_exp_tab[256] = GLIBC_EXP_TAB;
static const PowLogRow h_pow_tab[128] = GLIBC_POW_TAB;
static const TanRow h_tan_tab[186] = GLIBC_TAN_XFG;
static const LogRow h_logf_tab[16] = GLIBC_LOGF_TAB;
static const unsigned long long h_expf_tab[32] = GLIBC_EXPF_TAB;
static const double h_sincosf_tab[28] = GLIBC_SINCOSF_TAB;
static const float h_tanf_t[13] = GLIBC_TANF_T;
static const LogRow h_powf_tab[16] = GLIBC_POWF_TAB;
static const double h_powf_poly[5] = GLIBC_POWF_POLY;
static const double h_exp2f_poly[3] = GLIBC_EXP2F_POLY;
static const double h_logf_poly[3] = GLIBC_LOGF_POLY;
static const double h_expf_poly[3] = GLIBC_EXPF_POLY;
static const double h_pow_poly[7] = GLIBC_POW_POLY;
static const double h_erfc_pool[58] = GLIBC_ERFC_POOL;
static const double h_log_poly[5] = GLIBC_LOG_POLY;
static const double h_log_poly1[11] = GLIBC_LOG_POLY1;
#endif

// ---- primitives: one IEEE operation each, never contracted -------------------
#if defined(__CUDA_ARCH__)
GM_FN double f_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
GM_FN double f_add(double a, double b) { return __dadd_rn(a, b); }
GM_FN double f_sub(double a, double b) { return __dadd_rn(a, -b); }
GM_FN double f_mul(double a, double b) { return __dmul_rn(a, b); }
GM_FN double f_div(double a, double b) { return __ddiv_rn(a, b); }
GM_FN uint64_t f_bits(double x) { return static_cast<uint64_t>(__double_as_longlong(x)); }
GM_FN double f_from_bits(uint64_t u) { return __longlong_as_double(static_cast<long long>(u)); }
GM_FN double f_i2d(int k) { return __int2double_rn(k); }
GM_FN LogRow log_row(int i) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(&g_log_tab[i]));
  return LogRow{v.x, v.y};
}
GM_FN SinCosRow sincos_row(int k) {
  const double2* p = reinterpret_cast<const double2*>(&g_sincos_tab[4 * k]);
  const double2 a = __ldg(p), b = __ldg(p + 1);
  return SinCosRow{a.x, a.y, b.x, b.y};
}
GM_FN SinCosRow2 sincos_row2(int i) {
  const double2* p = reinterpret_cast<const double2*>(&g_sincos_rows[i]);
  const double2 a = __ldg(p), b = __ldg(p + 1);
  return SinCosRow2{a.x, a.y, b.x, b.y};
}
GM_FN double f_sel(bool c, double a, double b) { return c ? a : b; }
GM_FN void exp_row(int i, uint64_t& tail, uint64_t& sbits) {
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(&g_exp_tab[2 * i]));
  tail = v.x; sbits = v.y;
}
GM_FN PowLogRow pow_row(int i) {
  return PowLogRow{__ldg(&g_pow_tab[i].invc), __ldg(&g_pow_tab[i].logc), __ldg(&g_pow_tab[i].logctail)};
}
GM_FN double pow_special(double x, double y) { return ::pow(x, y); }
GM_FN TanRow tan_row(int i) {
  const double2* p = reinterpret_cast<const double2*>(&g_tan_tab[i]);
  const double2 a = __ldg(p);
  return TanRow{a.x, a.y, __ldg(&g_tan_tab[i].g), 0.0};
}
GM_FN double tan_huge(double x) { return ::tan(x); }
// single precision: one IEEE float operation each, never contracted
GM_FN float ff_add(float a, float b) { return __fadd_rn(a, b); }
GM_FN float ff_sub(float a, float b) { return __fadd_rn(a, -b); }
GM_FN float ff_mul(float a, float b) { return __fmul_rn(a, b); }
GM_FN float ff_div(float a, float b) { return __fdiv_rn(a, b); }
GM_FN uint32_t ff_bits(float x) { return static_cast<uint32_t>(__float_as_int(x)); }
GM_FN float ff_from_bits(uint32_t u) { return __int_as_float(static_cast<int>(u)); }
GM_FN double f_f2d(float x) { return static_cast<double>(x); }
GM_FN float f_d2f(double x) { return __double2float_rn(x); }
GM_FN LogRow logf_row(int i) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(&g_logf_tab[i]));
  return LogRow{v.x, v.y};
}
GM_FN uint64_t expf_row(int i) { return __ldg(&g_expf_tab[i]); }
GM_FN float cosf_large(float x) { return ::cosf(x); }
GM_FN float tanf_large(float x) { return ::tanf(x); }
GM_FN float ff_i2f(int k) { return __int2float_rn(k); }
GM_FN LogRow powf_row(int i) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(&g_powf_tab[i]));
  return LogRow{v.x, v.y};
}
GM_FN float powf_special(float x, float y) { return ::powf(x, y); }
#define GM_POWF_POLY c_powf_poly
#define GM_EXP2F_POLY c_exp2f_poly
#define GM_SINCOSF c_sincosf_tab
#define GM_TANF_T c_tanf_t
#define GM_LOGF_POLY c_logf_poly
#define GM_EXPF_POLY c_expf_poly
#define GM_POW_POLY c_pow_poly
#define GM_ERFC_POOL c_erfc_pool
#define GM_LOG_POLY c_log_poly
#define GM_LOG_POLY1 c_log_poly1
#else
GM_FN double f_fma(double a, double b, double c) { return __builtin_fma(a, b, c); }
GM_FN double f_add(double a, double b) { volatile double r = a + b; return r; }
GM_FN double f_sub(double a, double b) { volatile double r = a - b; return r; }
GM_FN double f_mul(double a, double b) { volatile double r = a * b; return r; }
GM_FN double f_div(double a, double b) { volatile double r = a / b; return r; }
GM_FN uint64_t f_bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
GM_FN double f_from_bits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
GM_FN double f_i2d(int k) { return static_cast<double>(k); }
GM_FN LogRow log_row(int i) { return h_log_tab[i]; }
GM_FN SinCosRow sincos_row(int k) {
  return SinCosRow{h_sincos_tab[4 * k], h_sincos_tab[4 * k + 1], h_sincos_tab[4 * k + 2], h_sincos_tab[4 * k + 3]};
}
GM_FN SinCosRow2 sincos_row2(int i) { return h_sincos_rows[i]; }
GM_FN double f_sel(bool c, double a, double b) { return c ? a : b; }
GM_FN void exp_row(int i, uint64_t& tail, uint64_t& sbits) { tail = h_exp_tab[2 * i]; sbits = h_exp_tab[2 * i + 1]; }
GM_FN PowLogRow pow_row(int i) { return h_pow_tab[i]; }
GM_FN double pow_special(double x, double y) { return __builtin_pow(x, y); }
GM_FN TanRow tan_row(int i) { return h_tan_tab[i]; }
GM_FN double tan_huge(double x) { return __builtin_tan(x); }
GM_FN float ff_add(float a, float b) { volatile float r = a + b; return r; }
GM_FN float ff_sub(float a, float b) { volatile float r = a - b; return r; }
GM_FN float ff_mul(float a, float b) { volatile float r = a * b; return r; }
GM_FN float ff_div(float a, float b) { volatile float r = a / b; return r; }
GM_FN uint32_t ff_bits(float x) { uint32_t u; std::memcpy(&u, &x, 4); return u; }
GM_FN float ff_from_bits(uint32_t u) { float x; std::memcpy(&x, &u, 4); return x; }
GM_FN double f_f2d(float x) { return static_cast<double>(x); }
GM_FN float f_d2f(double x) { volatile float r = static_cast<float>(x); return r; }
GM_FN LogRow logf_row(int i) { return h_logf_tab[i]; }
GM_FN uint64_t expf_row(int i) { return h_expf_tab[i]; }
GM_FN float cosf_large(float x) { return __builtin_cosf(x); }
GM_FN float tanf_large(float x) { return __builtin_tanf(x); }
GM_FN float ff_i2f(int k) { return static_cast<float>(k); }
GM_FN LogRow powf_row(int i) { return h_powf_tab[i]; }
GM_FN float powf_special(float x, float y) { return __builtin_powf(x, y); }
#define GM_POWF_POLY h_powf_poly
#define GM_EXP2F_POLY h_exp2f_poly
#define GM_SINCOSF h_sincosf_tab
#define GM_TANF_T h_tanf_t
#define GM_LOGF_POLY h_logf_poly
#define GM_EXPF_POLY h_expf_poly
#define GM_POW_POLY h_pow_poly
#define GM_ERFC_POOL h_erfc_pool
#define GM_LOG_POLY h_log_poly
#define GM_LOG_POLY1 h_log_poly1
#endif
// the two 32-bit words of a double: the exponent arithmetic of log needs the high one only
// (the constants it subtracts have a zero low word), which halves its integer instructions
#if defined(__CUDA_ARCH__)
GM_FN uint32_t f_hi(double x) { return static_cast<uint32_t>(__double2hiint(x)); }
GM_FN uint32_t f_lo(double x) { return static_cast<uint32_t>(__double2loint(x)); }
GM_FN double f_from_hilo(uint32_t hi, uint32_t lo) { return __hiloint2double(static_cast<int>(hi), static_cast<int>(lo)); }
#else
GM_FN uint32_t f_hi(double x) { return static_cast<uint32_t>(f_bits(x) >> 32); }
GM_FN uint32_t f_lo(double x) { return static_cast<uint32_t>(f_bits(x)); }
GM_FN double f_from_hilo(uint32_t hi, uint32_t lo) { return f_from_bits((static_cast<uint64_t>(hi) << 32) | lo); }
#endif
GM_FN double f_abs(double x) { return f_from_bits(f_bits(x) & 0x7fffffffffffffffull); }
GM_FN double f_neg(double x) { return f_from_bits(f_bits(x) ^ 0x8000000000000000ull); }

// ============================================================================
// log
// ============================================================================
constexpr uint64_t kLogOff = 0x3fe6000000000000ull;       // e_log.c: OFF
constexpr uint64_t kLogNearLo = 0x3fee000000000000ull;    // 1 - 2^-4
constexpr uint64_t kLogNearSpan = 0x0003090000000000ull;  // (1 + 0x1.09p-4) - (1 - 2^-4), in bits

// x in [1 - 2^-4, 1 + 0x1.09p-4): log1p(r) by a degree-11 polynomial; r^2 / 2 is
// taken from r split at 2^-27 so that the leading terms are exact.
GM_FN double log_near_one(double x) {
  if (f_bits(x) == 0x3ff0000000000000ull) return 0.0;
  const double* B = GM_LOG_POLY1;
  const double r = f_sub(x, 1.0);
  const double r2 = f_mul(r, r);
  const double r3 = f_mul(r, r2);
  // y = r3 (B1 + r B2 + r2 B3 + r3 (B4 + r B5 + r2 B6 + r3 (B7 + r B8 + r2 B9 + r3 B10)))
  const double p12 = f_fma(r, B[2], B[1]);            // B1 + r B2
  const double p45 = f_fma(r, B[5], B[4]);            // B4 + r B5
  const double p78 = f_fma(r, B[8], B[7]);            // B7 + r B8
  const double p123 = f_fma(r2, B[3], p12);
  const double p456 = f_fma(r2, B[6], p45);
  double q = f_fma(r2, B[9], p78);                    // B7 + r B8 + r2 B9
  q = f_fma(r3, B[10], q);
  q = f_fma(q, r3, p456);
  q = f_fma(q, r3, p123);                             // y / r3
  // rhi = r + w - w, w = r 2^27 (both fused with the multiplication)
  const double c27 = 0x1p27;
  const double t = f_fma(r, c27, r);                  // r + w
  const double rhi = f_fma(-c27, r, t);               // (r + w) - w
  const double rhi2 = f_mul(rhi, rhi);
  const double rlo = f_sub(r, rhi);
  const double hi = f_fma(rhi2, B[0], r);             // r + rhi^2 B0
  double lo = f_sub(r, hi);
  const double rsum = f_add(r, rhi);
  lo = f_fma(rhi2, B[0], lo);                         // r - hi + rhi^2 B0
  const double brlo = f_mul(B[0], rlo);
  lo = f_fma(brlo, rsum, lo);                         // += B0 rlo (rhi + r)
  const double y = f_fma(q, r3, lo);                  // y r3-part + lo
  return f_add(hi, y);
}

// Positive, finite, normal x outside the window around 1 (the caller checks).
// (xh, xl) = the words of x.  e_log.c: tmp = ix - OFF; i = (tmp >> 45) % 128; k = tmp >> 52;
// iz = ix - (tmp & 0xfff << 52) -- OFF has a zero low word, so all of it happens in the high word
GM_FN double log_main(uint32_t xh, uint32_t xl) {
  const double* A = GM_LOG_POLY;
  const uint32_t th = xh - static_cast<uint32_t>(kLogOff >> 32);
  const int i = static_cast<int>((th >> 13) & 127u);
  const int k = static_cast<int32_t>(th) >> 20;
  const double z = f_from_hilo(xh - (th & 0xfff00000u), xl);
  const LogRow row = log_row(i);
  const double kd = f_i2d(k);
  const double w = f_fma(kd, GLIBC_LOG_LN2HI, row.logc);
  const double r = f_fma(z, row.invc, -1.0);
  const double p12 = f_fma(r, A[2], A[1]);
  const double hi = f_add(r, w);
  const double r2 = f_mul(r, r);
  double lo = f_add(f_sub(w, hi), r);
  lo = f_fma(kd, GLIBC_LOG_LN2LO, lo);
  const double r3 = f_mul(r, r2);
  const double p34 = f_fma(r, A[4], A[3]);
  lo = f_fma(r2, A[0], lo);
  const double p = f_fma(p34, r2, p12);
  return f_add(f_fma(r3, p, lo), hi);
}

// log(x) for every double (special cases as e_log.c, without errno).
// the window [1 - 2^-4, 1 + 0x1.09p-4): both of its ends have a zero low word
GM_FN bool log_is_near_one(uint32_t hi) {
  return hi - static_cast<uint32_t>(kLogNearLo >> 32) < static_cast<uint32_t>(kLogNearSpan >> 32);
}

GM_FN double log(double x) {
  uint64_t ix = f_bits(x);
  if (log_is_near_one(static_cast<uint32_t>(ix >> 32))) return log_near_one(x);
  const uint32_t top = static_cast<uint32_t>(ix >> 48);
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
    if (ix * 2 == 0) return f_from_bits(0xfff0000000000000ull);                  // log(+-0) = -inf
    if (ix == 0x7ff0000000000000ull) return x;                                   // log(inf) = inf
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) {                         // negative or NaN
      return (ix & 0x7fffffffffffffffull) > 0x7ff0000000000000ull ? f_add(x, x) : f_from_bits(0x7ff8000000000000ull | 0x8000000000000000ull);
    }
    ix = f_bits(f_mul(x, 0x1p52));                                               // subnormal
    ix -= 52ull << 52;
  }
  return log_main(static_cast<uint32_t>(ix >> 32), static_cast<uint32_t>(ix));
}

// log of a positive, finite, normal number (a random_real() output, a ratio of
// positive normals ...): no special-case test.
GM_FN double log_positive_normal(double x) {
  const uint32_t hi = f_hi(x);
  if (log_is_near_one(hi)) return log_near_one(x);
  return log_main(hi, f_lo(x));
}

// N logs of positive, finite, normal numbers at once.  The window around 1
// (1/16 of uniform arguments) costs 27 FP64 operations against the main path's
// 15, and with 32 independent lanes some lane is in it for 87% of the calls, so
// in lock-step everyone pays for both, every time.  Here the main path runs for
// all N arguments and the window's arguments are then worked off in a loop that
// takes ONE pending argument per lane per trip: the warp makes max-over-lanes
// trips, 1.5 per 4 uniform arguments instead of 3.5.  Same operations on the
// same values as N calls of log_positive_normal().
template <int N>
GM_FN void log_positive_normal_batch(const double (&x)[N], double (&lg)[N]) {
  unsigned pend = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < N; ++i) {
    const uint32_t hi = f_hi(x[i]);
    if (log_is_near_one(hi)) pend |= 1u << i;
    lg[i] = log_main(hi, f_lo(x[i]));
  }
  while (pend) {
    const unsigned low = pend & (0u - pend);
    double xi = x[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 1; j < N; ++j) xi = (low == (1u << j)) ? x[j] : xi;
    const double r = log_near_one(xi);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < N; ++j) lg[j] = (low == (1u << j)) ? r : lg[j];
    pend ^= low;
  }
}

// ============================================================================
// exp
// ============================================================================
// sysdeps/ieee754/dbl-64/e_exp.c (Szabolcs Nagy): x = k ln2 / 128 + r, 2^(k/128) from
// a 128-row table {tail, scale bits}, degree-5 polynomial in r; the FMA build's
// fused operations written out, the scaled evaluation next to over- and underflow
// (specialcase(): 512 <= |x| < 1024) included; non-finite arguments and |x| >= 1024
// give IEEE's inf / 0 / NaN.
// specialcase() of e_exp.c / e_pow.c (positive scale): 2^k may be out of range although
// the result is not
GM_COLD double exp_specialcase(double tmp, uint64_t sbits, uint64_t ki) {
  if ((ki & 0x80000000ull) == 0) {                       // k > 0: scale's exponent may have overflowed
    const double sc = f_from_bits(sbits - (1009ull << 52));
    return f_mul(f_fma(sc, tmp, sc), 0x1p1009);
  }
  const double sc = f_from_bits(sbits + (1022ull << 52));     // k < 0: care in the subnormal range
  const double st = f_mul(tmp, sc);
  double y = f_add(sc, st);
  if (y < 1.0) {
    const double hi = f_add(y, 1.0);
    const double l0 = f_add(f_sub(sc, y), st);
    y = f_sub(f_add(f_add(f_add(f_sub(1.0, hi), y), l0), hi), 1.0);
    if (y == 0.0) y = 0.0;
  }
  return f_mul(y, 0x1p-1022);
}

GM_FN double exp(double x) {
  uint32_t abstop = static_cast<uint32_t>(f_bits(x) >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {
    if (abstop < 0x3c9u) return f_add(1.0, x);          // |x| < 2^-54
    if (abstop >= 0x409u) {                              // |x| >= 1024, inf, NaN
      if (f_bits(x) == 0xfff0000000000000ull) return 0.0;
      if (abstop >= 0x7ffu) return f_add(1.0, x);
      return (f_bits(x) >> 63) ? 0.0 : f_from_bits(0x7ff0000000000000ull);
    }
    abstop = 0;                                          // 512 <= |x| < 1024
  }
  const double zs = f_fma(x, GLIBC_EXP_INVLN2N, GLIBC_EXP_SHIFT);
  const uint64_t ki = f_bits(zs);
  const double kd = f_sub(zs, GLIBC_EXP_SHIFT);
  double r = f_fma(kd, GLIBC_EXP_NEGLN2HIN, x);
  r = f_fma(kd, GLIBC_EXP_NEGLN2LON, r);
  uint64_t tail_bits, sbits;
  exp_row(static_cast<int>(ki & 127u), tail_bits, sbits);
  sbits += ki << 45;
  const double p23 = f_fma(r, GLIBC_EXP_C3, GLIBC_EXP_C2);
  const double tr = f_add(r, f_from_bits(tail_bits));
  const double r2 = f_mul(r, r);
  const double p45 = f_fma(r, GLIBC_EXP_C5, GLIBC_EXP_C4);
  const double lo = f_fma(p23, r2, tr);
  const double tmp = f_fma(f_mul(r2, r2), p45, lo);
  if (abstop == 0) return exp_specialcase(tmp, sbits, ki);
  const double scale = f_from_bits(sbits);
  return f_fma(scale, tmp, scale);
}

// ============================================================================
// pow
// ============================================================================
// sysdeps/ieee754/dbl-64/e_pow.c (Szabolcs Nagy): log(x) as hi + lo with ~68 bits
// from a 128-row table {1/c, log c, tail} and a degree-7 polynomial, y log(x) as
// ehi + elo, then exp's table and polynomial with the tail folded in; the FMA
// build's fused operations written out.  x positive and normal, 2^-65 <= |y| < 2^63
// (results through over- and underflow included); everything else (zero /
// negative / subnormal / non-finite x, huge or tiny y) is
// the platform's pow -- the samplers raise uniforms to moderate powers.
// checkint() of e_pow.c: 0 = y is not an integer, 1 = odd, 2 = even
GM_FN int pow_checkint(uint64_t iy) {
  const int e = static_cast<int>(iy >> 52) & 0x7ff;
  if (e < 0x3ff) return 0;
  if (e > 0x3ff + 52) return 2;
  if (iy & ((1ull << (0x3ff + 52 - e)) - 1)) return 0;
  if (iy & (1ull << (0x3ff + 52 - e))) return 1;
  return 2;
}

GM_FN double pow_positive(uint64_t ix, double y);

GM_FN double pow(double x, double y) {
  const uint64_t ix = f_bits(x), iy = f_bits(y);
  const uint32_t topx = static_cast<uint32_t>(ix >> 52), topy = static_cast<uint32_t>(iy >> 52);
  if ((topy & 0x7ffu) - 0x3beu > 0x7fu) return pow_special(x, y);
  uint64_t ax = ix;
  bool negate = false;
  if (topx - 1u > 0x7fdu) {
    // a negative normal base with an integer exponent: the library goes on with |x| and
    // sets the sign of the scale factor for an odd exponent -- the negated result, bit for bit
    if ((topx & 0x800u) == 0 || (topx & 0x7ffu) - 1u > 0x7fdu) return pow_special(x, y);
    const int yint = pow_checkint(iy);
    if (yint == 0) return pow_special(x, y);
    ax = ix & 0x7fffffffffffffffull;
    negate = yint == 1;
  }
  const double r = pow_positive(ax, y);
  return negate ? f_neg(r) : r;
}

// x = from_bits(ix) positive and normal, 2^-65 <= |y| < 2^63
GM_FN double pow_positive(uint64_t ix, double y) {
  const double* A = GM_POW_POLY;
  // log_inline
  const uint64_t tmp = ix - 0x3fe6955500000000ull;
  const int i = static_cast<int>((tmp >> 45) & 127);
  const int k = static_cast<int>(static_cast<int64_t>(tmp) >> 52);
  const double z = f_from_bits(ix - (tmp & 0xfff0000000000000ull));
  const double kd = f_i2d(k);
  const PowLogRow row = pow_row(i);
  const double t1 = f_fma(kd, GLIBC_LOG_LN2HI, row.logc);
  const double lo1 = f_fma(kd, GLIBC_LOG_LN2LO, row.logctail);
  const double r = f_fma(z, row.invc, -1.0);
  const double ar = f_mul(r, A[0]);
  const double p12 = f_fma(r, A[2], A[1]);
  const double p34 = f_fma(r, A[4], A[3]);
  const double t2 = f_add(r, t1);
  const double lo2 = f_add(f_sub(t1, t2), r);
  const double ar2 = f_mul(r, ar);
  const double ar3 = f_mul(r, ar2);
  const double lo3 = f_fma(ar, r, -ar2);
  const double hi = f_add(t2, ar2);
  const double p56 = f_fma(r, A[6], A[5]);
  const double lo4 = f_add(f_sub(t2, hi), ar2);
  double p = f_fma(p56, ar2, p34);
  p = f_fma(ar2, p, p12);
  double lo = f_add(lo1, lo2);
  lo = f_add(lo, lo3);
  lo = f_add(lo, lo4);
  lo = f_fma(ar3, p, lo);
  const double lhi = f_add(hi, lo);
  const double llo = f_add(f_sub(hi, lhi), lo);
  // y log(x) = ehi + elo
  const double ehi = f_mul(y, lhi);
  double elo = f_fma(lhi, y, -ehi);
  elo = f_fma(y, llo, elo);
  // exp_inline
  uint32_t abstop = static_cast<uint32_t>(f_bits(ehi) >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {
    if (abstop < 0x3c9u) return f_add(1.0, ehi);         // |y log x| < 2^-54
    if (abstop >= 0x409u) return (f_bits(ehi) >> 63) ? 0.0 : f_from_bits(0x7ff0000000000000ull);
    abstop = 0;                                          // 512 <= |y log x| < 1024
  }
  const double zs = f_fma(ehi, GLIBC_EXP_INVLN2N, GLIBC_EXP_SHIFT);
  const uint64_t ki = f_bits(zs);
  const double kd2 = f_sub(zs, GLIBC_EXP_SHIFT);
  double rr = f_fma(kd2, GLIBC_EXP_NEGLN2HIN, ehi);
  rr = f_fma(kd2, GLIBC_EXP_NEGLN2LON, rr);
  rr = f_add(elo, rr);
  uint64_t tail_bits, sbits;
  exp_row(static_cast<int>(ki & 127u), tail_bits, sbits);
  sbits += ki << 45;
  const double q23 = f_fma(rr, GLIBC_EXP_C3, GLIBC_EXP_C2);
  const double tr = f_add(rr, f_from_bits(tail_bits));
  const double r2 = f_mul(rr, rr);
  const double q45 = f_fma(rr, GLIBC_EXP_C5, GLIBC_EXP_C4);
  const double lo5 = f_fma(q23, r2, tr);
  const double tmp2 = f_fma(f_mul(r2, r2), q45, lo5);
  if (abstop == 0) return exp_specialcase(tmp2, sbits, ki);
  const double scale = f_from_bits(sbits);
  return f_fma(tmp2, scale, scale);
}

// ============================================================================
// lgamma (positive arguments)
// ============================================================================
// sysdeps/ieee754/dbl-64/e_lgamma_r.c (fdlibm's lgamma): not a multiarch
// function, so the library's build has no FMA contraction -- every operator
// below is one IEEE operation, in the published order.  Its internal log() IS
// the ifunc'ed one (log above).  0 < x < 2: three polynomial / rational pieces
// around the minimum tc = 1.4616; 2 <= x < 8: rational in the fractional part plus
// log of the product (y + 2) ... (y + i - 1); 8 <= x < 2^58: Stirling in 1 / x;
// beyond: x (log x - 1).  Negative, zero, infinite and NaN arguments are the
// caller's business (samplers.cuh sends them to the toolkit's lgamma).
// 2 < x < 8 (x == 2 is answered before): rational in the fractional part, plus
// the log of (y + 2) ... (y + i - 1)
GM_FN double lgamma_2_to_8(double x) {
  double r;
  const int i = static_cast<int>(x);
  const double y = f_sub(x, f_i2d(i));
  double p = f_add(GLIBC_LG_S5, f_mul(y, GLIBC_LG_S6));
  p = f_add(GLIBC_LG_S4, f_mul(y, p));
  p = f_add(GLIBC_LG_S3, f_mul(y, p));
  p = f_add(GLIBC_LG_S2, f_mul(y, p));
  p = f_add(GLIBC_LG_S1, f_mul(y, p));
  p = f_add(GLIBC_LG_S0, f_mul(y, p));
  p = f_mul(y, p);
  double q = f_add(GLIBC_LG_R5, f_mul(y, GLIBC_LG_R6));
  q = f_add(GLIBC_LG_R4, f_mul(y, q));
  q = f_add(GLIBC_LG_R3, f_mul(y, q));
  q = f_add(GLIBC_LG_R2, f_mul(y, q));
  q = f_add(GLIBC_LG_R1, f_mul(y, q));
  q = f_add(1.0, f_mul(y, q));
  r = f_add(f_mul(0.5, y), f_div(p, q));
  if (i >= 3) {                                             // lgamma(1 + s) = log(s) + lgamma(s)
    double z = 1.0;
    if (i >= 7) z = f_mul(z, f_add(y, 6.0));
    if (i >= 6) z = f_mul(z, f_add(y, 5.0));
    if (i >= 5) z = f_mul(z, f_add(y, 4.0));
    if (i >= 4) z = f_mul(z, f_add(y, 3.0));
    z = f_mul(z, f_add(y, 2.0));
    r = f_add(r, log(z));
  }
  return r;
}

// 8 <= x < 2^58: (x - 1/2)(log x - 1) + w(1 / x)
GM_FN double lgamma_stirling(double x) {
  const double t = log_main(f_hi(x), f_lo(x));  // x >= 8: log()'s main path, no window, no special case
  const double z = f_div(1.0, x);
  const double y = f_mul(z, z);
  double w = f_add(GLIBC_LG_W5, f_mul(y, GLIBC_LG_W6));
  w = f_add(GLIBC_LG_W4, f_mul(y, w));
  w = f_add(GLIBC_LG_W3, f_mul(y, w));
  w = f_add(GLIBC_LG_W2, f_mul(y, w));
  w = f_add(GLIBC_LG_W1, f_mul(y, w));
  w = f_add(GLIBC_LG_W0, f_mul(z, w));
  return f_add(f_mul(f_sub(x, 0.5), f_sub(t, 1.0)), w);
}

GM_FN double lgamma_positive(double x) {
  const uint32_t ix = static_cast<uint32_t>(f_bits(x) >> 32) & 0x7fffffffu;
  const uint32_t lx = static_cast<uint32_t>(f_bits(x));
  if (ix < 0x3b900000u) return f_neg(log(x));                 // x < 2^-70
  if ((((ix - 0x3ff00000u) | lx) == 0) || (((ix - 0x40000000u) | lx) == 0)) return 0.0;      // 1 and 2
  double r;
  if (ix < 0x40000000u) {                                     // x < 2
    double y;
    int i;
    if (ix <= 0x3feccccc) {                                   // lgamma(x) = lgamma(x + 1) - log(x)
      r = f_neg(log(x));
      if (ix >= 0x3FE76944u) { y = f_sub(1.0, x); i = 0; }
      else if (ix >= 0x3FCDA661u) { y = f_sub(x, GLIBC_LG_TC_M1); i = 1; }
      else { y = x; i = 2; }
    } else {
      r = 0.0;
      if (ix >= 0x3FFBB4C3u) { y = f_sub(2.0, x); i = 0; }    // [1.7316, 2]
      else if (ix >= 0x3FF3B4C4u) { y = f_sub(x, GLIBC_LG_TC); i = 1; }      // [1.23, 1.73]
      else { y = f_sub(x, 1.0); i = 2; }
    }
    if (i == 0) {
      const double z = f_mul(y, y);
      double p1 = f_add(GLIBC_LG_A8, f_mul(z, GLIBC_LG_A10));
      p1 = f_add(GLIBC_LG_A6, f_mul(z, p1));
      p1 = f_add(GLIBC_LG_A4, f_mul(z, p1));
      p1 = f_add(GLIBC_LG_A2, f_mul(z, p1));
      p1 = f_add(GLIBC_LG_A0, f_mul(z, p1));
      double p2 = f_add(GLIBC_LG_A9, f_mul(z, GLIBC_LG_A11));
      p2 = f_add(GLIBC_LG_A7, f_mul(z, p2));
      p2 = f_add(GLIBC_LG_A5, f_mul(z, p2));
      p2 = f_add(GLIBC_LG_A3, f_mul(z, p2));
      p2 = f_add(GLIBC_LG_A1, f_mul(z, p2));
      p2 = f_mul(z, p2);
      const double p = f_add(f_mul(y, p1), p2);
      r = f_add(r, f_sub(p, f_mul(0.5, y)));
    } else if (i == 1) {
      const double z = f_mul(y, y);
      const double w = f_mul(z, y);
      double p1 = f_add(GLIBC_LG_T9, f_mul(w, GLIBC_LG_T12));
      p1 = f_add(GLIBC_LG_T6, f_mul(w, p1));
      p1 = f_add(GLIBC_LG_T3, f_mul(w, p1));
      p1 = f_add(GLIBC_LG_T0, f_mul(w, p1));
      double p2 = f_add(GLIBC_LG_T10, f_mul(w, GLIBC_LG_T13));
      p2 = f_add(GLIBC_LG_T7, f_mul(w, p2));
      p2 = f_add(GLIBC_LG_T4, f_mul(w, p2));
      p2 = f_add(GLIBC_LG_T1, f_mul(w, p2));
      double p3 = f_add(GLIBC_LG_T11, f_mul(w, GLIBC_LG_T14));
      p3 = f_add(GLIBC_LG_T8, f_mul(w, p3));
      p3 = f_add(GLIBC_LG_T5, f_mul(w, p3));
      p3 = f_add(GLIBC_LG_T2, f_mul(w, p3));
      const double p = f_sub(f_mul(z, p1), f_sub(GLIBC_LG_TT, f_mul(w, f_add(p2, f_mul(y, p3)))));
      r = f_add(r, f_add(GLIBC_LG_TF, p));
    } else {
      double p1 = f_add(GLIBC_LG_U4, f_mul(y, GLIBC_LG_U5));
      p1 = f_add(GLIBC_LG_U3, f_mul(y, p1));
      p1 = f_add(GLIBC_LG_U2, f_mul(y, p1));
      p1 = f_add(GLIBC_LG_U1, f_mul(y, p1));
      p1 = f_add(GLIBC_LG_U0, f_mul(y, p1));
      p1 = f_mul(y, p1);
      double p2 = f_add(GLIBC_LG_V4, f_mul(y, GLIBC_LG_V5));
      p2 = f_add(GLIBC_LG_V3, f_mul(y, p2));
      p2 = f_add(GLIBC_LG_V2, f_mul(y, p2));
      p2 = f_add(GLIBC_LG_V1, f_mul(y, p2));
      p2 = f_add(1.0, f_mul(y, p2));
      r = f_add(r, f_add(f_mul(-0.5, y), f_div(p1, p2)));
    }
    return r;
  }
  if (ix < 0x40200000u) return lgamma_2_to_8(x);               // 2 <= x < 8
  if (ix < 0x43900000u) return lgamma_stirling(x);            // 8 <= x < 2^58
  return f_mul(x, f_sub(log(x), 1.0));                        // 2^58 <= x
}

// N lgammas of positive finite arguments at once.  The three ranges of the
// function are three different programs; lanes with arguments in different
// ranges serialise them, and a density with six non-integer lgammas per
// observation (beta-binomial) pays 6 x (all three).  Here the Stirling range --
// where most arguments of the count densities live -- is evaluated for all N
// arguments without a branch (the formula is harmless, if meaningless, below 8),
// and the arguments of the other two ranges are then worked off in two loops that
// take ONE pending argument per lane per trip.  Same operations on the same
// values as N calls of lgamma_positive().
template <int N>
GM_FN void lgamma_positive_batch(const double (&x)[N], double (&out)[N]) {
  unsigned pend_mid = 0, pend_low = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < N; ++i) {
    const uint32_t ix = static_cast<uint32_t>(f_bits(x[i]) >> 32) & 0x7fffffffu;
    const uint32_t lx = static_cast<uint32_t>(f_bits(x[i]));
    const bool stirling = ix >= 0x40200000u && ix < 0x43900000u;
    const bool mid = ix >= 0x40000000u && ix < 0x40200000u && ((ix - 0x40000000u) | lx) != 0;
    if (mid) pend_mid |= 1u << i;
    if (!stirling && !mid) pend_low |= 1u << i;
    out[i] = lgamma_stirling(x[i]);
  }
  while (pend_mid) {
    const unsigned low = pend_mid & (0u - pend_mid);
    double xi = x[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 1; j < N; ++j) xi = (low == (1u << j)) ? x[j] : xi;
    const double r = lgamma_2_to_8(xi);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < N; ++j) out[j] = (low == (1u << j)) ? r : out[j];
    pend_mid ^= low;
  }
  while (pend_low) {                 // below 2, exactly 2, or 2^58 and above
    const unsigned low = pend_low & (0u - pend_low);
    double xi = x[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 1; j < N; ++j) xi = (low == (1u << j)) ? x[j] : xi;
    const double r = lgamma_positive(xi);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < N; ++j) out[j] = (low == (1u << j)) ? r : out[j];
    pend_low ^= low;
  }
}

// ============================================================================
// cos
// ============================================================================
// TAYLOR_SIN(a*a, a, da), |a| < 0.126
GM_FN double taylor_sin(double a, double da) {
  const double xx = f_mul(a, a);
  double t = f_fma(xx, GLIBC_SC_S5, GLIBC_SC_S4);
  t = f_fma(xx, t, GLIBC_SC_S3);
  t = f_fma(xx, t, GLIBC_SC_S2);
  t = f_fma(xx, t, GLIBC_SC_S1);
  const double h = f_mul(da, 0.5);
  t = f_fma(t, a, -h);
  t = f_fma(xx, t, da);
  return f_add(a, t);
}

// do_cos(a, da): cos(a + da), |a| < 0.86
GM_FN double do_cos(double a, double da) {
  if (a < 0) da = f_neg(da);
  const double aa = f_abs(a);
  const double u = f_add(aa, GLIBC_SC_BIG);
  const int k = static_cast<int>(static_cast<uint32_t>(f_bits(u)));
  const double x = f_add(f_sub(aa, f_sub(u, GLIBC_SC_BIG)), da);
  const double xx = f_mul(x, x);
  const double ps = f_fma(xx, GLIBC_SC_SN5, GLIBC_SC_SN3);
  const double s = f_fma(f_mul(x, xx), ps, x);
  double pc = f_fma(xx, GLIBC_SC_CS6, GLIBC_SC_CS4);
  pc = f_fma(xx, pc, GLIBC_SC_CS2);
  const double c = f_mul(xx, pc);
  const SinCosRow t = sincos_row(k);
  double cor = f_fma(-s, t.ssn, t.ccs);
  cor = f_fma(-c, t.cs, cor);
  cor = f_fma(-s, t.sn, cor);
  return f_add(t.cs, cor);
}

// do_sin(a, da): sin(a + da), |a| < 0.86
GM_FN double do_sin(double a, double da) {
  const double aa = f_abs(a);
  if (aa < GLIBC_SC_SIN_TAYLOR_LIMIT) return taylor_sin(a, da);
  if (a <= 0) da = f_neg(da);
  const double u = f_add(aa, GLIBC_SC_BIG);
  const int k = static_cast<int>(static_cast<uint32_t>(f_bits(u)));
  const double x = f_sub(aa, f_sub(u, GLIBC_SC_BIG));
  const double xx = f_mul(x, x);
  const double ps = f_fma(xx, GLIBC_SC_SN5, GLIBC_SC_SN3);
  const double s = f_add(x, f_fma(f_mul(x, xx), ps, da));
  double pc = f_fma(xx, GLIBC_SC_CS6, GLIBC_SC_CS4);
  pc = f_fma(xx, pc, GLIBC_SC_CS2);
  const double c = f_fma(x, da, f_mul(xx, pc));
  const SinCosRow t = sincos_row(k);
  double cor = f_fma(s, t.ccs, t.ssn);
  cor = f_fma(-c, t.sn, cor);
  cor = f_fma(s, t.cs, cor);
  const double res = f_add(t.sn, cor);
  return f_from_bits((f_bits(res) & 0x7fffffffffffffffull) | (f_bits(a) & 0x8000000000000000ull));
}

// cos(x) for |x| < 105414350 (s_sin.c: __cos, the branches below 2^26.65); the
// Box-Muller angle is in (0, 2 pi].  Larger and non-finite arguments are the
// caller's business (samplers.cuh routes them to the toolkit's cos).
GM_FN double cos_medium(double x) {
  const uint32_t k = static_cast<uint32_t>(f_bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;                       // |x| < 2^-27
  if (k < 0x3feb6000u) return do_cos(x, 0.0);            // |x| < 0.855469
  if (k < 0x400368fdu) {                                 // |x| < 2.426265
    const double y = f_sub(GLIBC_SC_HP0, f_abs(x));
    const double a = f_add(y, GLIBC_SC_HP1);
    const double da = f_add(f_sub(y, a), GLIBC_SC_HP1);
    return do_sin(a, da);
  }
  // reduce_sincos: x - n pi/2 as a + da
  const double t = f_fma(x, GLIBC_SC_HPINV, GLIBC_SC_TOINT);
  const double xn = f_sub(t, GLIBC_SC_TOINT);
  const int n = static_cast<int>(static_cast<uint32_t>(f_bits(t))) & 3;
  double y = f_fma(-xn, GLIBC_SC_MP1, x);
  y = f_fma(-xn, GLIBC_SC_MP2, y);
  const double b = f_fma(-xn, GLIBC_SC_PP3, y);
  double db = f_fma(-xn, GLIBC_SC_PP3, f_sub(y, b));
  const double a = f_fma(-xn, GLIBC_SC_PP4, b);
  const double e = f_fma(-xn, GLIBC_SC_PP4, f_sub(b, a));
  const double da = f_add(db, e);
  // do_sincos(a, da, n + 1)
  const double r = (n & 1) ? do_sin(a, da) : do_cos(a, da);
  return ((n + 1) & 2) ? f_neg(r) : r;
}

// ---------------------------------------------------------------------------
// tan (s_tan.c, the IBM routine with its multi-precision fall-backs removed, as
// shipped since glibc 2.28; operation order and fusing follow __tan_fma).
// The cauchy sampler and the Poisson Cauchy-envelope regime call tan(pi u), u in (0, 1].
// ---------------------------------------------------------------------------

// tan or -cot of the reduced argument a + da (|a| <= pi/4): cases VI / VII (and their
// copies for the 25 < |x| <= 1e8 reduction) of __tan
GM_FN double tan_reduced(double a, double da, int n) {
  const bool neg = a < 0.0;
  const double ya = neg ? f_neg(a) : a;
  if (ya <= GLIBC_TAN_G2) {
    const double a2 = f_mul(a, a);
    double t2 = f_fma(a2, GLIBC_TAN_D11, GLIBC_TAN_D9);
    t2 = f_fma(a2, t2, GLIBC_TAN_D7);
    t2 = f_fma(a2, t2, GLIBC_TAN_D5);
    t2 = f_fma(a2, t2, GLIBC_TAN_D3);
    t2 = f_fma(f_mul(a, a2), t2, da);
    const double b = f_add(a, t2);
    if (n == 0) return b;
    // -cot: EADD(a, t2, b, db); DIV2(1, 0, b, db, c, dc); -(c + dc)
    const double db = (f_abs(a) > f_abs(t2)) ? f_add(f_sub(a, b), t2) : f_add(f_sub(t2, b), a);
    const double c = f_div(1.0, b);
    const double u = f_mul(c, b);
    const double uu = f_fma(c, b, -u);
    double r = f_sub(1.0, u);
    r = f_sub(r, uu);
    r = f_add(r, 0.0);
    r = f_fma(-db, c, r);
    const double cc = f_div(r, b);
    const double z = f_add(c, cc);
    const double zz = f_add(f_sub(c, z), cc);
    return f_neg(f_add(zz, z));
  }
  const double yya = neg ? f_neg(da) : da;
  const int i = static_cast<int>(f_fma(ya, 256.0, -15.5));
  const TanRow row = tan_row(i);
  const double z = f_add(f_sub(ya, row.x), yya);
  const double z2 = f_mul(z, z);
  const double pz = f_fma(f_mul(z, z2), f_fma(z2, GLIBC_TAN_E1, GLIBC_TAN_E0), z);
  const double num = f_mul(f_add(row.f, row.g), pz);
  if (n) {
    const double y = f_sub(row.g, f_div(num, f_add(pz, row.f)));
    return neg ? y : f_neg(y);                       // -sy * y
  }
  const double y = f_add(f_div(num, f_sub(row.g, pz)), row.f);
  return neg ? f_neg(y) : y;                         // sy * y
}

GM_FN double tan(double x) {
  const uint32_t hx = static_cast<uint32_t>(f_bits(x) >> 32);
  if ((hx & 0x7ff00000u) == 0x7ff00000u) return f_sub(x, x);
  const double w = f_abs(x);
  if (w <= GLIBC_TAN_G1) return x;
  if (w <= GLIBC_TAN_G2) {
    const double x2 = f_mul(x, x);
    double t2 = f_fma(x2, GLIBC_TAN_D11, GLIBC_TAN_D9);
    t2 = f_fma(x2, t2, GLIBC_TAN_D7);
    t2 = f_fma(x2, t2, GLIBC_TAN_D5);
    t2 = f_fma(x2, t2, GLIBC_TAN_D3);
    return f_fma(f_mul(x, x2), t2, x);
  }
  if (w <= GLIBC_TAN_G3) {
    const int i = static_cast<int>(f_fma(w, 256.0, -15.5));
    const TanRow row = tan_row(i);
    const double z = f_sub(w, row.x);
    const double z2 = f_mul(z, z);
    const double pz = f_fma(f_mul(z, z2), f_fma(z2, GLIBC_TAN_E1, GLIBC_TAN_E0), z);
    const double t2 = f_div(f_mul(f_add(row.f, row.g), pz), f_sub(row.g, pz));
    const double y = f_add(t2, row.f);
    return (x < 0.0) ? f_neg(y) : y;
  }
  if (w > GLIBC_TAN_G5) return tan_huge(x);
  const double t = f_fma(x, GLIBC_SC_HPINV, GLIBC_SC_TOINT);
  const double xn = f_sub(t, GLIBC_SC_TOINT);
  const int n = static_cast<int>(static_cast<uint32_t>(f_bits(t))) & 1;
  double t1 = f_fma(-xn, GLIBC_SC_MP1, x);
  t1 = f_fma(-xn, GLIBC_SC_MP2, t1);
  if (w <= GLIBC_TAN_G4) {
    const double a = f_fma(-xn, GLIBC_TAN_MP3, t1);
    const double da = f_fma(-xn, GLIBC_TAN_MP3, f_sub(t1, a));
    return tan_reduced(a, da, n);
  }
  const double a0 = f_fma(-xn, GLIBC_SC_PP3, t1);
  double da = f_fma(-xn, GLIBC_SC_PP3, f_sub(t1, a0));
  const double a1 = f_fma(-xn, GLIBC_SC_PP4, a0);
  da = f_add(da, f_fma(-xn, GLIBC_SC_PP4, f_sub(a0, a1)));
  const double a = f_add(a1, da);
  const double db = (f_abs(a1) > f_abs(da)) ? f_add(f_sub(a1, a), da) : f_add(f_sub(da, a), a1);
  return tan_reduced(a, db, n);
}

// ---------------------------------------------------------------------------
// log1p (s_log1p.c: fdlibm's routine with glibc's split polynomial; fusing as in
// __log1p_fma).  Used by the cauchy log-density and the negative binomial's
// large-size limit.  x > -1 and finite here; the rest: IEEE special values.
// ---------------------------------------------------------------------------
GM_FN double log1p(double x) {
  const int32_t hx = static_cast<int32_t>(f_bits(x) >> 32);
  const int32_t ax = hx & 0x7fffffff;
  int k = 1;
  int32_t hu = 1;
  double f = x, c = 0.0;
  if (hx < 0x3fda827a) {                                    // x < 0.41422
    if (ax >= 0x3ff00000) {                                 // x <= -1
      if (x == -1.0) return f_div(-0x1p54, 0.0);
      const double d = f_sub(x, x);
      return f_div(d, d);
    }
    if (ax < 0x3e200000) {                                  // |x| < 2^-29
      if (ax < 0x3c900000) return x;
      return f_fma(-f_mul(x, x), 0.5, x);
    }
    if (hx > 0 || hx <= static_cast<int32_t>(0xbfd2bec3u)) k = 0;     // -0.2929 < x < 0.41422
  } else if (hx >= 0x7ff00000) {
    return f_add(x, x);
  }
  if (k != 0) {
    double u;
    if (hx < 0x43400000) {
      u = f_add(1.0, x);
      hu = static_cast<int32_t>(f_bits(u) >> 32);
      k = (hu >> 20) - 1023;
      c = (k > 0) ? f_sub(1.0, f_sub(u, x)) : f_sub(x, f_sub(u, 1.0));
      c = f_div(c, u);
    } else {
      u = x;
      hu = hx;
      k = (hu >> 20) - 1023;
    }
    hu &= 0x000fffff;
    const uint64_t lo = f_bits(u) & 0xffffffffull;
    if (hu < 0x6a09e) {
      u = f_from_bits((static_cast<uint64_t>(static_cast<uint32_t>(hu) | 0x3ff00000u) << 32) | lo);
    } else {
      k += 1;
      u = f_from_bits((static_cast<uint64_t>(static_cast<uint32_t>(hu) | 0x3fe00000u) << 32) | lo);
      hu = (0x00100000 - hu) >> 2;
    }
    f = f_sub(u, 1.0);
  }
  const double hfsq = f_mul(f_mul(0.5, f), f);
  const double kd = f_i2d(k);
  if (hu == 0) {                                            // |f| < 2^-20
    if (f == 0.0) {
      if (k == 0) return 0.0;
      return f_fma(kd, GLIBC_LOG1P_LN2_HI, f_fma(kd, GLIBC_LOG1P_LN2_LO, c));
    }
    const double R = f_mul(hfsq, f_fma(-f, GLIBC_LOG1P_TWO_THIRDS, 1.0));
    if (k == 0) return f_sub(f, R);
    return f_fma(kd, GLIBC_LOG1P_LN2_HI, -f_sub(f_sub(R, f_fma(kd, GLIBC_LOG1P_LN2_LO, c)), f));
  }
  const double s = f_div(f, f_add(2.0, f));
  const double z = f_mul(s, s);
  const double R2 = f_fma(z, GLIBC_LOG1P_LP3, GLIBC_LOG1P_LP2);
  const double R3 = f_fma(z, GLIBC_LOG1P_LP5, GLIBC_LOG1P_LP4);
  const double R4 = f_fma(z, GLIBC_LOG1P_LP7, GLIBC_LOG1P_LP6);
  const double z2 = f_mul(z, z);
  const double z4 = f_mul(z2, z2);
  const double z6 = f_mul(z4, z2);
  double R = f_fma(z, GLIBC_LOG1P_LP1, f_mul(z2, R2));
  R = f_fma(z4, R3, R);
  R = f_fma(z6, R4, R);
  const double t = f_mul(f_add(R, hfsq), s);
  if (k == 0) return f_sub(f, f_sub(hfsq, t));
  const double e = f_add(f_fma(kd, GLIBC_LOG1P_LN2_LO, c), t);
  return f_fma(kd, GLIBC_LOG1P_LN2_HI, -f_sub(f_sub(hfsq, e), f));
}

// ---------------------------------------------------------------------------
// erfc (s_erf.c: fdlibm's routine with glibc's split polynomials; plain SSE2 code in the
// library -- not an ifunc, nothing fused).  C[] is the routine's constant pool in the
// library's own order: a coefficient the code subtracts is stored as its magnitude, so the
// additions and subtractions below are the library's.  The two exponentials are e_exp.c
// (exp above).  Used by the truncated normal density.
// ---------------------------------------------------------------------------
GM_FN double erfc(double x) {
  const double* C = GM_ERFC_POOL;
  const int32_t hx = static_cast<int32_t>(f_bits(x) >> 32);
  const int32_t ix = hx & 0x7fffffff;
  if (ix >= 0x7ff00000) return f_add(f_i2d(static_cast<int>((static_cast<uint32_t>(hx) >> 31) << 1)), f_div(1.0, x));
  if (ix < 0x3feb0000) {                                    // |x| < 0.84375
    if (ix < 0x3c700000) return f_sub(1.0, x);
    const double z = f_mul(x, x);
    const double z2 = f_mul(z, z);
    const double z4 = f_mul(z2, z2);
    const double r2 = f_sub(f_mul(C[0], z), C[1]);
    const double r1 = f_add(f_mul(C[2], z), C[3]);
    const double s2 = f_add(f_mul(C[5], z), C[6]);
    const double s1 = f_add(f_mul(C[7], z), 1.0);
    const double s3 = f_add(f_mul(z, C[8]), C[9]);
    const double r = f_add(f_mul(C[4], z4), f_add(f_mul(r2, z2), r1));
    const double sd = f_add(f_mul(s3, z4), f_add(f_mul(s2, z2), s1));
    const double xy = f_mul(f_div(r, sd), x);
    if (hx < 0x3fd00000) return f_sub(1.0, f_add(xy, x));
    return f_sub(0.5, f_add(f_sub(x, 0.5), xy));
  }
  if (ix < 0x3ff40000) {                                    // 0.84375 <= |x| < 1.25
    const double s = f_sub(f_abs(x), 1.0);
    const double s2 = f_mul(s, s);
    const double s4 = f_mul(s2, s2);
    const double s6 = f_mul(s2, s4);
    double P = f_add(f_mul(f_sub(f_mul(C[10], s), C[11]), s2), f_sub(f_mul(C[12], s), C[13]));
    P = f_add(P, f_mul(f_sub(f_mul(C[14], s), C[15]), s4));
    P = f_add(P, f_mul(C[16], s6));
    double Q = f_add(f_mul(f_add(f_mul(C[17], s), C[18]), s2), f_add(f_mul(C[19], s), 1.0));
    Q = f_add(f_mul(f_add(f_mul(s, C[20]), C[21]), s4), Q);
    Q = f_add(Q, f_mul(s6, C[22]));
    const double pq = f_div(P, Q);
    if (hx < 0) return f_add(1.0, f_add(pq, C[23]));
    return f_sub(C[57], pq);
  }
  if (ix < 0x403c0000) {                                    // |x| < 28
    const double ax = f_abs(x);
    const double s = f_div(1.0, f_mul(x, x));
    const double s2 = f_mul(s, s);
    const double s4 = f_mul(s2, s2);
    const double s6 = f_mul(s2, s4);
    double R, S;
    if (ix < 0x4006db6d) {                                  // |x| < 1 / 0.35
      R = f_add(f_mul(f_sub(f_mul(C[26], s), C[27]), s2), f_sub(f_mul(C[28], s), C[29]));
      R = f_add(R, f_mul(f_sub(f_mul(C[30], s), C[31]), s4));
      R = f_add(R, f_mul(f_sub(f_mul(C[32], s), C[33]), s6));
      S = f_add(f_mul(f_add(f_mul(C[34], s), C[35]), s2), f_add(1.0, f_mul(C[36], s)));
      S = f_add(S, f_mul(f_add(f_mul(C[37], s), C[38]), s4));
      S = f_add(f_mul(f_add(f_mul(s, C[39]), C[40]), s6), S);
      S = f_add(S, f_mul(f_mul(s4, s4), C[41]));
    } else {
      if (hx < 0 && ix >= 0x40180000) return f_sub(2.0, C[25]);         // x < -6
      R = f_add(f_mul(f_sub(f_mul(C[42], s), C[43]), s2), f_sub(f_mul(C[44], s), C[45]));
      R = f_add(R, f_mul(f_sub(f_mul(C[46], s), C[47]), s4));
      R = f_add(R, f_mul(C[48], s6));
      S = f_add(f_mul(f_add(f_mul(C[49], s), C[50]), s2), f_add(1.0, f_mul(C[51], s)));
      S = f_add(S, f_mul(f_add(f_mul(C[52], s), C[53]), s4));
      S = f_add(S, f_mul(f_add(f_mul(s, C[54]), C[55]), s6));
    }
    const double z = f_from_bits(f_bits(ax) & 0xffffffff00000000ull);
    const double e1 = glibc::exp(f_sub(f_mul(f_neg(z), z), C[56]));
    const double e2 = glibc::exp(f_add(f_mul(f_sub(z, ax), f_add(z, ax)), f_div(R, S)));
    const double r = f_mul(e2, e1);
    if (hx > 0) return f_div(r, ax);
    return f_sub(2.0, f_div(r, ax));
  }
  if (hx > 0) return f_mul(C[25], C[25]);
  return f_sub(2.0, C[25]);
}

// ============================================================================
// single precision (real_type = float): logf, expf, cosf, lgammaf
// ============================================================================
// e_logf.c / e_expf.c / s_sincosf.h are Szabolcs Nagy's float routines: double
// arithmetic on the widened argument, one rounding to float at the end (FMA
// builds, fused operations explicit).  e_lgammaf_r.c is fdlibm's lgamma in float
// arithmetic (no contraction), with the same structure as lgamma_positive.

// every argument: subnormals are normalised as in the library, zero / negative / non-finite give IEEE's values
GM_FN float logf(float x) {
  uint32_t ix = ff_bits(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {      // x < 2^-126, inf or NaN
    if (ix * 2u == 0) return ff_from_bits(0xff800000u);      // log(+-0) = -inf
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) { const float d = ff_sub(x, x); return ff_div(d, d); }
    ix = ff_bits(ff_mul(x, 0x1p23f)) - (23u << 23);          // subnormal: normalise
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = static_cast<int>((tmp >> 19) & 15u);
  const int k = static_cast<int32_t>(tmp) >> 23;
  const double z = f_f2d(ff_from_bits(ix - (tmp & 0xff800000u)));
  const LogRow row = logf_row(i);
  const double* A = GM_LOGF_POLY;
  const double y0 = f_fma(f_i2d(k), GLIBC_LOGF_LN2, row.logc);
  const double r = f_fma(z, row.invc, -1.0);
  double y = f_fma(r, A[1], A[2]);
  const double r2 = f_mul(r, r);
  const double s = f_add(r, y0);
  y = f_fma(r2, A[0], y);
  return f_d2f(f_fma(r2, y, s));
}

// every argument, the over- and underflow thresholds of the library included
GM_FN float expf(float x) {
  const uint32_t abstop = (ff_bits(x) >> 20) & 0x7ffu;
  if (abstop > 0x42au) {                                     // |x| >= 88 or NaN
    if (ff_bits(x) == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8u) return ff_add(x, x);
    if (x > 0x1.62e42ep6f) return ff_from_bits(0x7f800000u);        // above log(2^128)
    if (x < -0x1.9fe368p6f) return 0.0f;                            // below log(2^-150)
    if (x < -0x1.9d1d9ep6f) return 0x1p-149f;                       // the library's "may underflow" value
  }
  const double xd = f_f2d(x);
  const double* C = GM_EXPF_POLY;
  const double zs = f_fma(GLIBC_EXPF_INVLN2N, xd, GLIBC_EXPF_SHIFT);
  const uint64_t ki = f_bits(zs);
  const double kd = f_sub(zs, GLIBC_EXPF_SHIFT);
  const double r = f_fma(GLIBC_EXPF_INVLN2N, xd, -kd);
  const double s = f_from_bits(expf_row(static_cast<int>(ki & 31u)) + (ki << 47));
  const double z = f_fma(r, C[0], C[1]);
  const double r2 = f_mul(r, r);
  double y = f_fma(r, C[2], 1.0);
  y = f_fma(z, r2, y);
  return f_d2f(f_mul(y, s));
}

// cos / sin polynomials of s_sincosf.h on a table entry p = {sign[4], hpi_inv, hpi,
// c0, c1, s1, c2, s2, c3, s3, c4}
GM_FN float sincosf_poly(double x, double x2, const double* p, int n) {
  if ((n & 1) == 0) {
    const double x3 = f_mul(x2, x);
    const double s1 = f_fma(x2, p[12], p[10]);
    const double x7 = f_mul(x2, x3);
    const double s = f_fma(x3, p[8], x);
    return f_d2f(f_fma(s1, x7, s));
  }
  const double x4 = f_mul(x2, x2);
  const double c1 = f_fma(x2, p[7], p[6]);
  const double c2 = f_fma(x2, p[13], p[11]);
  const double x6 = f_mul(x2, x4);
  const double c = f_fma(x4, p[9], c1);
  return f_d2f(f_fma(c2, x6, c));
}

// |x| < 120 here (the Box-Muller angle is below 2 pi); beyond: the platform's cosf
GM_FN float cosf(float x) {
  const double* T = GM_SINCOSF;
  const uint32_t abstop = (ff_bits(x) >> 20) & 0x7ffu;
  const double y = f_f2d(x);
  if (abstop <= 0x3f3u) {                       // |x| < pi / 4
    if (abstop <= 0x397u) return 1.0f;          // |x| < 2^-12
    return sincosf_poly(y, f_mul(y, y), T, 1);
  }
  if (abstop > 0x42eu) return cosf_large(x);
  const double r = f_mul(y, T[4]);
  const int n = (static_cast<int>(r) + 0x800000) >> 24;
  const double xr = f_fma(-f_i2d(n), T[5], y);
  const double sgn = T[n & 3];
  const double* p = (n & 2) ? T + 14 : T;
  const double x2 = f_mul(xr, xr);
  return sincosf_poly((n & 1) ? f_mul(xr, sgn) : xr, x2, p, n ^ 1);
}

// k_tanf.c (fdlibm's kernel in float arithmetic, no contraction): tan(x + y) for iy = 1,
// -1 / tan(x + y) for iy = -1, |x| <= pi / 4
GM_FN float kernel_tanf(float x, float y, int iy) {
  const float* T = GM_TANF_T;
  const uint32_t hx = ff_bits(x);
  const uint32_t ix = hx & 0x7fffffffu;
  if (ix < 0x39000000u) {                                    // |x| < 2^-13, (int) x == 0
    if ((ix | static_cast<uint32_t>(iy + 1)) == 0) return ff_div(1.0f, ff_from_bits(ix));
    if (iy == 1) return x;
    return ff_div(-1.0f, x);
  }
  const bool big = ix >= 0x3f2ca140u;                        // |x| >= 0.6744
  if (big) {
    if (hx >> 31) { x = ff_from_bits(hx ^ 0x80000000u); y = ff_from_bits(ff_bits(y) ^ 0x80000000u); }
    const float z = ff_sub(GLIBC_TANF_PIO4, x);
    const float w = ff_sub(GLIBC_TANF_PIO4LO, y);
    x = ff_add(z, w);
    y = 0.0f;
    if (ff_from_bits(ff_bits(x) & 0x7fffffffu) < 0x1p-13f) {
      const int sg = (1 - static_cast<int>((hx >> 30) & 2u)) * iy;
      return ff_mul(ff_i2f(sg), ff_sub(1.0f, ff_mul(ff_i2f(2 * iy), x)));
    }
  }
  const float z = ff_mul(x, x);
  const float w = ff_mul(z, z);
  float r = ff_add(ff_mul(T[11], w), T[9]);
  r = ff_add(ff_mul(r, w), T[7]);
  r = ff_add(ff_mul(r, w), T[5]);
  r = ff_add(ff_mul(r, w), T[3]);
  r = ff_add(ff_mul(r, w), T[1]);
  float v = ff_add(ff_mul(T[12], w), T[10]);
  v = ff_add(ff_mul(v, w), T[8]);
  v = ff_add(ff_mul(v, w), T[6]);
  v = ff_add(ff_mul(v, w), T[4]);
  v = ff_add(ff_mul(v, w), T[2]);
  v = ff_mul(z, v);
  const float s = ff_mul(z, x);
  r = ff_add(y, ff_mul(z, ff_add(ff_mul(s, ff_add(r, v)), y)));
  r = ff_add(r, ff_mul(T[0], s));
  const float wx = ff_add(x, r);
  if (big) {
    const float vi = ff_i2f(iy);
    const float q = ff_sub(ff_div(ff_mul(wx, wx), ff_add(wx, vi)), r);
    const float e = ff_sub(x, q);
    const float res = ff_sub(vi, ff_add(e, e));
    return ff_mul(ff_i2f(1 - static_cast<int>((hx >> 30) & 2u)), res);
  }
  if (iy == 1) return wx;
  // -1 / (x + r) with the quotient's error corrected
  const float zz = ff_from_bits(ff_bits(wx) & 0xfffff000u);
  const float vv = ff_sub(r, ff_sub(zz, x));
  const float a = ff_div(-1.0f, wx);
  const float t = ff_from_bits(ff_bits(a) & 0xfffff000u);
  const float ss = ff_add(1.0f, ff_mul(t, zz));
  return ff_add(t, ff_mul(a, ff_add(ss, ff_mul(t, vv))));
}

// s_tanf.c (2.39): |x| <= pi / 4 straight to the kernel; below 120 the sincosf reduction
// in double (x - n pi/2, unfused in this routine), split into a float head and tail.
// Beyond 120 and for non-finite arguments: the platform's tanf.
GM_FN float tanf(float x) {
  const uint32_t ix = ff_bits(x) & 0x7fffffffu;
  if (ix <= 0x3f490fdau) return kernel_tanf(x, 0.0f, 1);
  if (((ix >> 20) & 0x7ffu) > 0x42eu) return tanf_large(x);
  const double* P = GM_SINCOSF;
  const double xd = f_f2d(x);
  const int n = (static_cast<int>(f_mul(xd, P[4])) + 0x800000) >> 24;
  const double xr = f_sub(xd, f_mul(f_i2d(n), P[5]));
  const float y0 = f_d2f(xr);
  const float y1 = f_d2f(f_sub(xr, f_f2d(y0)));
  return kernel_tanf(y0, y1, 1 - ((n & 1) << 1));
}

// s_log1pf.c (fdlibm's log1p in float arithmetic, Horner polynomial; not an ifunc, nothing
// fused).  Constants as published: ln2_hi 0x3f317180, ln2_lo 0x3717f7d1, Lp1..Lp7.
GM_FN float log1pf(float x) {
  const float ln2_hi = ff_from_bits(0x3f317180u), ln2_lo = ff_from_bits(0x3717f7d1u);
  const float Lp1 = ff_from_bits(0x3f2aaaabu), Lp2 = ff_from_bits(0x3ecccccdu), Lp3 = ff_from_bits(0x3e924925u),
              Lp4 = ff_from_bits(0x3e638e29u), Lp5 = ff_from_bits(0x3e3a3325u), Lp6 = ff_from_bits(0x3e1cd04fu),
              Lp7 = ff_from_bits(0x3e178897u);
  const int32_t hx = static_cast<int32_t>(ff_bits(x));
  const int32_t ax = hx & 0x7fffffff;
  int k = 1;
  int32_t hu = 1;
  float f = x, c = 0.0f;
  if (hx < 0x3ed413d7) {                                    // x < 0.41422
    if (ax >= 0x3f800000) {                                 // x <= -1
      if (x == -1.0f) return ff_div(-0x1p25f, 0.0f);
      const float d = ff_sub(x, x);
      return ff_div(d, d);
    }
    if (ax < 0x31000000) {                                  // |x| < 2^-29
      if (ax < 0x24800000) return x;
      return ff_sub(x, ff_mul(ff_mul(x, x), 0.5f));
    }
    if (hx > 0 || hx <= static_cast<int32_t>(0xbe95f61fu)) k = 0;
  } else if (hx >= 0x7f800000) {
    return ff_add(x, x);
  }
  if (k != 0) {
    float u;
    if (hx < 0x5a000000) {
      u = ff_add(1.0f, x);
      hu = static_cast<int32_t>(ff_bits(u));
      k = (hu >> 23) - 127;
      c = (k > 0) ? ff_sub(1.0f, ff_sub(u, x)) : ff_sub(x, ff_sub(u, 1.0f));
      c = ff_div(c, u);
    } else {
      u = x;
      hu = hx;
      k = (hu >> 23) - 127;
    }
    hu &= 0x007fffff;
    if (hu < 0x3504f7) {
      u = ff_from_bits(static_cast<uint32_t>(hu) | 0x3f800000u);
    } else {
      k += 1;
      u = ff_from_bits(static_cast<uint32_t>(hu) | 0x3f000000u);
      hu = (0x00800000 - hu) >> 2;
    }
    f = ff_sub(u, 1.0f);
  }
  const float hfsq = ff_mul(ff_mul(0.5f, f), f);
  const float kf = ff_i2f(k);
  if (hu == 0) {                                            // |f| < 2^-20
    if (f == 0.0f) {
      if (k == 0) return 0.0f;
      return ff_add(ff_add(ff_mul(kf, ln2_lo), c), ff_mul(kf, ln2_hi));
    }
    const float R = ff_mul(hfsq, ff_sub(1.0f, ff_mul(Lp1, f)));
    if (k == 0) return ff_sub(f, R);
    return ff_sub(ff_mul(kf, ln2_hi), ff_sub(ff_sub(R, ff_add(ff_mul(kf, ln2_lo), c)), f));
  }
  const float s = ff_div(f, ff_add(2.0f, f));
  const float z = ff_mul(s, s);
  float R = ff_add(ff_mul(Lp7, z), Lp6);
  R = ff_add(ff_mul(R, z), Lp5);
  R = ff_add(ff_mul(R, z), Lp4);
  R = ff_add(ff_mul(R, z), Lp3);
  R = ff_add(ff_mul(R, z), Lp2);
  R = ff_add(ff_mul(R, z), Lp1);
  R = ff_mul(R, z);
  const float t = ff_mul(ff_add(R, hfsq), s);
  if (k == 0) return ff_sub(f, ff_sub(hfsq, t));
  const float e = ff_add(ff_add(ff_mul(kf, ln2_lo), c), t);
  return ff_sub(ff_mul(kf, ln2_hi), ff_sub(ff_sub(hfsq, e), f));
}

// e_powf.c: log2(x) in double from a 16-row table and a degree-5 polynomial, y log2(x),
// exp2 by exp2f's table.  x positive and normal, y finite and non-zero, |y log2 x| < 126
// here; everything else: the platform's powf.
// checkint() of e_powf.c: 0 = y is not an integer, 1 = odd, 2 = even
GM_FN int powf_checkint(uint32_t iy) {
  const int e = static_cast<int>(iy >> 23) & 0xff;
  if (e < 0x7f) return 0;
  if (e > 0x7f + 23) return 2;
  if (iy & ((1u << (0x7f + 23 - e)) - 1)) return 0;
  if (iy & (1u << (0x7f + 23 - e))) return 1;
  return 2;
}

GM_FN float powf(float x, float y) {
  uint32_t ix = ff_bits(x);
  const uint32_t iy = ff_bits(y);
  bool negate = false;
  if (ix - 0x00800000u > 0x7effffffu || 2u * iy - 1u > 0xfefffffeu) {
    // zero / infinite / NaN operands: IEEE-defined results, the platform's powf
    if (2u * iy - 1u > 0xfefffffeu || 2u * ix - 1u > 0xfefffffeu) return powf_special(x, y);
    if (ix & 0x80000000u) {                                  // finite x < 0: integer y only
      const int yint = powf_checkint(iy);
      if (yint == 0) { const float d = ff_sub(x, x); return ff_div(d, d); }
      negate = yint == 1;
      ix &= 0x7fffffffu;
    }
    if (ix < 0x00800000u) ix = (ff_bits(ff_mul(ff_from_bits(ix), 0x1p23f)) & 0x7fffffffu) - (23u << 23);   // subnormal: normalise
  }
  const double* A = GM_POWF_POLY;
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = static_cast<int>((tmp >> 19) & 15u);
  const uint32_t top = tmp & 0xff800000u;
  const int k = static_cast<int32_t>(top) >> 23;
  const double z = f_f2d(ff_from_bits(ix - top));
  const LogRow row = powf_row(i);
  const double r = f_fma(z, row.invc, -1.0);
  const double y0 = f_add(f_i2d(k), row.logc);
  double ya = f_fma(r, A[0], A[1]);
  const double p = f_fma(r, A[2], A[3]);
  const double r2 = f_mul(r, r);
  double q = f_fma(r, A[4], y0);
  const double r4 = f_mul(r2, r2);
  q = f_fma(r2, p, q);
  ya = f_fma(ya, r4, q);
  const double ylogx = f_mul(f_f2d(y), ya);
  if (((f_bits(ylogx) >> 47) & 0xffffu) > 0x80beu) {         // |y log2 x| >= 126
    if (ylogx > 0x1.fffffffd1d571p+6) return ff_from_bits(negate ? 0xff800000u : 0x7f800000u);
    if (ylogx <= -150.0) return ff_from_bits(negate ? 0x80000000u : 0u);
    if (ylogx < -149.0) return ff_from_bits(negate ? 0x80000001u : 0x00000001u);          // the library's "may underflow" value
  }
  const double* C = GM_EXP2F_POLY;
  const double kd0 = f_add(ylogx, GLIBC_EXP2F_SHIFT_SCALED);
  const uint64_t ki = f_bits(kd0);
  const double kd = f_sub(kd0, GLIBC_EXP2F_SHIFT_SCALED);
  const double rr = f_sub(ylogx, kd);
  const double s = f_from_bits(expf_row(static_cast<int>(ki & 31u)) + (ki << 47));
  const double zz = f_fma(rr, C[0], C[1]);
  const double rr2 = f_mul(rr, rr);
  double yy = f_fma(rr, C[2], 1.0);
  yy = f_fma(zz, rr2, yy);
  const float res = f_d2f(f_mul(yy, s));
  return negate ? ff_from_bits(ff_bits(res) ^ 0x80000000u) : res;
}

GM_FN float lgammaf_2_to_8(float x) {
  const int i = static_cast<int>(x);
  const float y = ff_sub(x, static_cast<float>(i));
  float p = ff_add(GLIBC_LGF_S5, ff_mul(y, GLIBC_LGF_S6));
  p = ff_add(GLIBC_LGF_S4, ff_mul(y, p));
  p = ff_add(GLIBC_LGF_S3, ff_mul(y, p));
  p = ff_add(GLIBC_LGF_S2, ff_mul(y, p));
  p = ff_add(GLIBC_LGF_S1, ff_mul(y, p));
  p = ff_add(GLIBC_LGF_S0, ff_mul(y, p));
  p = ff_mul(y, p);
  float q = ff_add(GLIBC_LGF_R5, ff_mul(y, GLIBC_LGF_R6));
  q = ff_add(GLIBC_LGF_R4, ff_mul(y, q));
  q = ff_add(GLIBC_LGF_R3, ff_mul(y, q));
  q = ff_add(GLIBC_LGF_R2, ff_mul(y, q));
  q = ff_add(GLIBC_LGF_R1, ff_mul(y, q));
  q = ff_add(1.0f, ff_mul(y, q));
  float r = ff_add(ff_mul(0.5f, y), ff_div(p, q));
  if (i >= 3) {
    float z = 1.0f;
    if (i >= 7) z = ff_mul(z, ff_add(y, 6.0f));
    if (i >= 6) z = ff_mul(z, ff_add(y, 5.0f));
    if (i >= 5) z = ff_mul(z, ff_add(y, 4.0f));
    if (i >= 4) z = ff_mul(z, ff_add(y, 3.0f));
    z = ff_mul(z, ff_add(y, 2.0f));
    r = ff_add(r, logf(z));
  }
  return r;
}

// x positive and finite (e_lgammaf_r.c; the thresholds are the float images of
// lgamma_positive's)
GM_FN float lgammaf_positive(float x) {
  const uint32_t ix = ff_bits(x);
  if (ix < 0x30800000u) return -logf(x);                       // x < 2^-30
  if (ix == 0x3f800000u || ix == 0x40000000u) return 0.0f;     // 1 and 2
  float r;
  if (ix < 0x40000000u) {                                      // x < 2
    float y;
    int i;
    if (ix <= 0x3f666666u) {
      r = -logf(x);
      if (ix >= 0x3f3b4a20u) { y = ff_sub(1.0f, x); i = 0; }
      else if (ix >= 0x3e6d3308u) { y = ff_sub(x, GLIBC_LGF_TC_M1); i = 1; }
      else { y = x; i = 2; }
    } else {
      r = 0.0f;
      if (ix >= 0x3fdda618u) { y = ff_sub(2.0f, x); i = 0; }
      else if (ix >= 0x3f9da620u) { y = ff_sub(x, GLIBC_LGF_TC); i = 1; }
      else { y = ff_sub(x, 1.0f); i = 2; }
    }
    if (i == 0) {
      const float z = ff_mul(y, y);
      float p1 = ff_add(GLIBC_LGF_A8, ff_mul(z, GLIBC_LGF_A10));
      p1 = ff_add(GLIBC_LGF_A6, ff_mul(z, p1));
      p1 = ff_add(GLIBC_LGF_A4, ff_mul(z, p1));
      p1 = ff_add(GLIBC_LGF_A2, ff_mul(z, p1));
      p1 = ff_add(GLIBC_LGF_A0, ff_mul(z, p1));
      float p2 = ff_add(GLIBC_LGF_A9, ff_mul(z, GLIBC_LGF_A11));
      p2 = ff_add(GLIBC_LGF_A7, ff_mul(z, p2));
      p2 = ff_add(GLIBC_LGF_A5, ff_mul(z, p2));
      p2 = ff_add(GLIBC_LGF_A3, ff_mul(z, p2));
      p2 = ff_add(GLIBC_LGF_A1, ff_mul(z, p2));
      p2 = ff_mul(z, p2);
      const float p = ff_add(ff_mul(y, p1), p2);
      r = ff_add(r, ff_sub(p, ff_mul(0.5f, y)));
    } else if (i == 1) {
      const float z = ff_mul(y, y);
      const float w = ff_mul(z, y);
      float p1 = ff_add(GLIBC_LGF_T9, ff_mul(w, GLIBC_LGF_T12));
      p1 = ff_add(GLIBC_LGF_T6, ff_mul(w, p1));
      p1 = ff_add(GLIBC_LGF_T3, ff_mul(w, p1));
      p1 = ff_add(GLIBC_LGF_T0, ff_mul(w, p1));
      float p2 = ff_add(GLIBC_LGF_T10, ff_mul(w, GLIBC_LGF_T13));
      p2 = ff_add(GLIBC_LGF_T7, ff_mul(w, p2));
      p2 = ff_add(GLIBC_LGF_T4, ff_mul(w, p2));
      p2 = ff_add(GLIBC_LGF_T1, ff_mul(w, p2));
      float p3 = ff_add(GLIBC_LGF_T11, ff_mul(w, GLIBC_LGF_T14));
      p3 = ff_add(GLIBC_LGF_T8, ff_mul(w, p3));
      p3 = ff_add(GLIBC_LGF_T5, ff_mul(w, p3));
      p3 = ff_add(GLIBC_LGF_T2, ff_mul(w, p3));
      const float p = ff_sub(ff_mul(z, p1), ff_sub(GLIBC_LGF_TT, ff_mul(w, ff_add(p2, ff_mul(y, p3)))));
      r = ff_add(r, ff_add(GLIBC_LGF_TF, p));
    } else {
      float p1 = ff_add(GLIBC_LGF_U4, ff_mul(y, GLIBC_LGF_U5));
      p1 = ff_add(GLIBC_LGF_U3, ff_mul(y, p1));
      p1 = ff_add(GLIBC_LGF_U2, ff_mul(y, p1));
      p1 = ff_add(GLIBC_LGF_U1, ff_mul(y, p1));
      p1 = ff_add(GLIBC_LGF_U0, ff_mul(y, p1));
      p1 = ff_mul(y, p1);
      float p2 = ff_add(GLIBC_LGF_V4, ff_mul(y, GLIBC_LGF_V5));
      p2 = ff_add(GLIBC_LGF_V3, ff_mul(y, p2));
      p2 = ff_add(GLIBC_LGF_V2, ff_mul(y, p2));
      p2 = ff_add(GLIBC_LGF_V1, ff_mul(y, p2));
      p2 = ff_add(1.0f, ff_mul(y, p2));
      r = ff_add(r, ff_add(ff_mul(-0.5f, y), ff_div(p1, p2)));
    }
    return r;
  }
  if (ix < 0x41000000u) return lgammaf_2_to_8(x);               // 2 <= x < 8
  if (ix < 0x5c800000u) {                                      // 8 <= x < 2^58
    const float t = logf(x);
    const float z = ff_div(1.0f, x);
    const float y = ff_mul(z, z);
    float w = ff_add(GLIBC_LGF_W5, ff_mul(y, GLIBC_LGF_W6));
    w = ff_add(GLIBC_LGF_W4, ff_mul(y, w));
    w = ff_add(GLIBC_LGF_W3, ff_mul(y, w));
    w = ff_add(GLIBC_LGF_W2, ff_mul(y, w));
    w = ff_add(GLIBC_LGF_W1, ff_mul(y, w));
    w = ff_add(GLIBC_LGF_W0, ff_mul(z, w));
    return ff_add(ff_mul(ff_sub(x, 0.5f), ff_sub(t, 1.0f)), w);
  }
  return ff_mul(x, ff_sub(logf(x), 1.0f));                     // 2^58 <= x
}

// ---- the same function for lock-step execution ---------------------------------
// cos_medium() branches four ways on the argument and then two or three ways
// on the quadrant; 32 lanes with independent angles take every path, so a warp
// pays for all of them in turn (Box-Muller fell from 163 to 64 Gdraws/s).  The
// formulation below computes the SAME operations on the same values -- every
// result bit is unchanged, checked against libm like cos_medium -- arranged so
// that all lanes run one instruction stream:
//   * both argument reductions are evaluated (3 + 11 operations) and selected;
//   * do_sin and do_cos share the table look-up, x^2, both polynomials and the
//     correction chain: with rows stored as {P, p, Q, q} = {sn, ssn, cs, ccs}
//     for sin and {cs, ccs, -sn, -ssn} for cos, both are
//         cor = s q + p;  cor = cor - c P;  cor = cor + s Q;  result = P + cor
//     (negating an operand of an FMA is exact); only s and c differ, and both
//     variants are computed (4 operations) and selected;
//   * TAYLOR_SIN (|a| < 0.126, 8% of the angles) stays a branch.
GM_FN double sincos_core(double a, double da, bool is_sin) {
  const double aa = f_abs(a);
  const bool flip = is_sin ? (a <= 0) : (a < 0);
  const double dx = flip ? f_neg(da) : da;
  const double u = f_add(aa, GLIBC_SC_BIG);
  const int k = static_cast<int>(static_cast<uint32_t>(f_bits(u)));
  const double x0 = f_sub(aa, f_sub(u, GLIBC_SC_BIG));
  const double x = f_sel(is_sin, x0, f_add(x0, dx));
  const double xx = f_mul(x, x);
  const double ps = f_fma(xx, GLIBC_SC_SN5, GLIBC_SC_SN3);
  const double xxx = f_mul(x, xx);
  double pc = f_fma(xx, GLIBC_SC_CS6, GLIBC_SC_CS4);
  pc = f_fma(xx, pc, GLIBC_SC_CS2);
  const double cc = f_mul(xx, pc);
  const double s_cos = f_fma(xxx, ps, x);
  const double s_sin = f_add(x, f_fma(xxx, ps, dx));
  const double c_sin = f_fma(x, dx, cc);
  const double s = f_sel(is_sin, s_sin, s_cos);
  const double c = f_sel(is_sin, c_sin, cc);
  const SinCosRow2 t = sincos_row2(k + (is_sin ? 110 : 0));
  double cor = f_fma(s, t.q, t.p);
  cor = f_fma(-c, t.P, cor);
  cor = f_fma(s, t.Q, cor);
  const double res = f_add(t.P, cor);
  const uint64_t sign = is_sin ? (f_bits(a) & 0x8000000000000000ull) : (f_bits(res) & 0x8000000000000000ull);
  return f_from_bits((f_bits(res) & 0x7fffffffffffffffull) | sign);
}

GM_FN double cos_medium_lockstep(double x) {
  const uint32_t k = static_cast<uint32_t>(f_bits(x) >> 32) & 0x7fffffffu;
  // |x| >= 2.426265: reduce_sincos
  const double t = f_fma(x, GLIBC_SC_HPINV, GLIBC_SC_TOINT);
  const double xn = f_sub(t, GLIBC_SC_TOINT);
  const int n = static_cast<int>(static_cast<uint32_t>(f_bits(t))) & 3;
  double y = f_fma(-xn, GLIBC_SC_MP1, x);
  y = f_fma(-xn, GLIBC_SC_MP2, y);
  const double b = f_fma(-xn, GLIBC_SC_PP3, y);
  const double db = f_fma(-xn, GLIBC_SC_PP3, f_sub(y, b));
  const double a_d = f_fma(-xn, GLIBC_SC_PP4, b);
  const double da_d = f_add(db, f_fma(-xn, GLIBC_SC_PP4, f_sub(b, a_d)));
  // 0.855469 <= |x| < 2.426265: sin(pi/2 - |x|)
  const double yc = f_sub(GLIBC_SC_HP0, f_abs(x));
  const double a_c = f_add(yc, GLIBC_SC_HP1);
  const double da_c = f_add(f_sub(yc, a_c), GLIBC_SC_HP1);
  const bool in_b = k < 0x3feb6000u, in_c = !in_b && k < 0x400368fdu;
  const double a = in_b ? x : (in_c ? a_c : a_d);
  const double da = in_b ? 0.0 : (in_c ? da_c : da_d);
  const bool is_sin = in_c || (!in_b && (n & 1));
  const bool neg = !in_b && !in_c && ((n + 1) & 2);
  double r = sincos_core(a, da, is_sin);
  if (is_sin && f_abs(a) < GLIBC_SC_SIN_TAYLOR_LIMIT) r = taylor_sin(a, da);
  if (neg) r = f_neg(r);
  // |x| < 2^-27 (s_sin.c returns 1.0 outright): do_cos gives the same bits, its
  // correction term is below 2^-55, so no test is needed here
  return r;
}

}  // namespace glibc
}  // namespace mb200
