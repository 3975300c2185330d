// density.cuh -- the (log-)densities of the reference as device functions.
//
// ref: inst/include/monty/random/density.hpp (cited per function as :lines).
// T is the reference's template parameter (double, or float for
// density_negative_binomial_mu's is_float variant, src/density.cpp:28-45).
// Literal types follow the reference so the float instantiation promotes to
// double exactly where the reference's does.  Compiled with -fmad=false.
#pragma once
#include "samplers.cuh"   // libm overloads, Lim<T>

namespace mb200 {

// log1p, erfc: glibc's s_log1p.c / s_erf.c bit for bit (glibc_math.cuh)
__device__ __forceinline__ double m_log1p(double x) { return glibc::log1p(x); }
__device__ __forceinline__ double m_erfc(double x) { return glibc::erfc(x); }
__device__ __forceinline__ float m_log1p(float x) { return glibc::log1pf(x); }      // s_log1pf.c

template <typename T> __device__ __forceinline__ T maybe_log(T x, bool lg) {   // :33-36
  return lg ? x : m_exp(x);
}
template <typename T> __device__ __forceinline__ T d_lchoose(T n, T k) {       // :38-43
  return m_lgamma(static_cast<T>(n + 1)) - m_lgamma(static_cast<T>(k + 1)) -
    m_lgamma(static_cast<T>(n - k + 1));
}
template <typename T> __device__ __forceinline__ T d_lbeta(T x, T y) {         // :45-49
  return m_lgamma(x) + m_lgamma(y) - m_lgamma(x + y);
}

template <typename T> __device__ T dens_binomial(T x, T size, T prob, bool lg) {   // :53-71
  T ret;
  if (x == 0 && size == 0) ret = 0;
  else ret = d_lchoose<T>(size, x) + x * m_log(prob) + (size - x) * m_log(1 - prob);
  return maybe_log(ret, lg);
}

template <typename T> struct NormIntegral;
template <> struct NormIntegral<double> { static constexpr double v = 0.918938533204672741780329736406; };
template <> struct NormIntegral<float> { static constexpr float v = 0.918938533204672741780329736406f; };

template <typename T> __device__ T dens_normal(T x, T mu, T sd, bool lg) {         // :73-92
  if (sd == 0) return maybe_log(x - mu == 0 ? Lim<T>::inf() : -Lim<T>::inf(), lg);
  const T dx = x - mu;
  const T ret = - dx * dx / (2 * sd * sd) - NormIntegral<T>::v - m_log(sd);
  return maybe_log(ret, lg);
}

template <typename T> __device__ T dens_negative_binomial_mu(T x, T size, T mu, bool lg) {   // :94-129
  T ret;
  if (x == 0 && size == 0) {
    ret = 0;
  } else if (x < 0 || size == 0) {
    ret = -Lim<T>::inf();
  } else if (mu == 0) {
    ret = x == 0 ? 0 : -Lim<T>::inf();
  } else {
    const T ratio = Lim<T>::eps * 100;
    if (mu < ratio * size) {
      const T log_prob = m_log(mu / (1 + mu / size));
      ret = x * log_prob - mu - m_lgamma(x + 1) + m_log1p(x * (x - 1) / (2 * size));
    } else {
      const T prob = size / (size + mu);
      if constexpr (sizeof(T) == 8) {
        // the two non-integer lgammas as one batch (same bits as two calls)
        const double arg[2] = {x + size, size};
        double lgv[2];
        m_lgamma_batch<2>(arg, lgv);
        ret = lgv[0] - lgv[1] - m_lgamma(x + 1) + size * m_log(prob) + x * m_log(1 - prob);
      } else {
        ret = m_lgamma(x + size) - m_lgamma(size) - m_lgamma(x + 1) +
          size * m_log(prob) + x * m_log(1 - prob);
      }
    }
  }
  return maybe_log(ret, lg);
}

template <typename T> __device__ T dens_negative_binomial_prob(T x, T size, T prob, bool lg) {   // :133-137
  const T mu = size * (1 - prob) / prob;
  return dens_negative_binomial_mu<T>(x, size, mu, lg);
}

template <typename T> __device__ T dens_beta_binomial_ab(T x, T size, T a, T b, bool lg) {   // :139-156
  T ret;
  if (x == 0 && size == 0) ret = 0;
  else if (x > size) ret = -Lim<T>::inf();
  else if constexpr (sizeof(T) == 8) {
    // lbeta(p, q) - lbeta(a, b) = six lgammas of non-integer arguments: one batch
    // (same bits, same order of the sums as d_lbeta: lgamma(x) + lgamma(y) - lgamma(x + y))
    const double p = x + a, q = size - x + b;
    const double arg[6] = {p, q, p + q, a, b, a + b};
    double lgv[6];
    m_lgamma_batch<6>(arg, lgv);
    ret = d_lchoose<T>(size, x) + (lgv[0] + lgv[1] - lgv[2]) - (lgv[3] + lgv[4] - lgv[5]);
  }
  else ret = d_lchoose<T>(size, x) + d_lbeta<T>(x + a, size - x + b) - d_lbeta<T>(a, b);
  return maybe_log(ret, lg);
}

template <typename T> __device__ T dens_beta_binomial_prob(T x, T size, T prob, T rho, bool lg) {   // :166-171
  const T a = prob * (1 / rho - 1);
  const T b = (1 - prob) * (1 / rho - 1);
  return dens_beta_binomial_ab<T>(x, size, a, b, lg);
}

template <typename T> __device__ T dens_poisson(T x, T lambda, bool lg) {           // :173-189
  T ret;
  if (x == 0 && lambda == 0) ret = 0;
  else ret = x * m_log(lambda) - lambda - m_lgamma(x + 1);
  return maybe_log(ret, lg);
}

template <typename T> __device__ T dens_exponential_rate(T x, T rate, bool lg) {    // :191-204
  T ret;
  if (x < 0) ret = -Lim<T>::inf(); else ret = m_log(rate) - rate * x;
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_exponential_mean(T x, T mean, bool lg) {    // :207-220
  T ret;
  if (x < 0) ret = -Lim<T>::inf(); else ret = -m_log(mean) - x / mean;
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_gamma_rate(T x, T shape, T rate, bool lg) { // :222-236
  T ret;
  if (x < 0) ret = -Lim<T>::inf();
  else ret = (shape - 1) * m_log(x) - rate * x - m_lgamma(shape) + shape * m_log(rate);
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_gamma_scale(T x, T shape, T scale, bool lg) {   // :238-252
  T ret;
  if (x < 0) ret = -Lim<T>::inf();
  else ret = (shape - 1) * m_log(x) - x / scale - m_lgamma(shape) - shape * m_log(scale);
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_uniform(T x, T min, T max, bool lg) {       // :255-263
  const T ret = (x < min || x > max) ? 0 : 1 / (max - min);
  return lg ? m_log(ret) : ret;
}
template <typename T> __device__ T dens_hypergeometric(T x, T n1, T n2, T k, bool lg) {   // :265-274
  const T ret = d_lchoose<T>(n1, x) + d_lchoose<T>(n2, k - x) - d_lchoose<T>(n1 + n2, k);
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_beta(T x, T a, T b, bool lg) {              // :276-292
  T ret;
  if (x < 0 || x > 1) ret = -Lim<T>::inf();
  else ret = (a - 1) * m_log(x) + (b - 1) * m_log(1 - x) - d_lbeta<T>(a, b);
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_truncated_normal(T x, T mu, T sd, T min, T max, bool lg) {   // :294-313
  T ret;
  if (x < min || x > max) {
    ret = -Lim<T>::inf();
  } else {
    const T d = dens_normal<T>(x, mu, sd, true);
    const T y_min = (min - mu) / sd;
    const T y_max = (max - mu) / sd;
    // -y / M_SQRT2 is a double whatever T is: std::erfc(double)
    const double z_min = 0.5 * m_erfc(-y_min / 1.41421356237309504880);
    const double z_max = 0.5 * m_erfc(-y_max / 1.41421356237309504880);
    const double z = z_max - z_min;
    ret = d - m_log(z);
  }
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_cauchy(T x, T location, T scale, bool lg) { // :315-324
  // monty::math::log(M_PI) is the double log, so the sum is a double and
  // maybe_log<double> is what gets deduced
  const T q = (x - location) / scale;
  // std::pow(q, 2): the host compiler folds a square into one multiplication (GCC does so at
  // every optimisation level above -O0, with or without -ffast-math), which is what the
  // compiled reference executes; libm's pow(q, 2.0) can differ from q * q in the last bit
  const T q2 = q * q;
  const double ret = -m_log(M_PI) - m_log(scale) - m_log1p(q2);
  return lg ? static_cast<T>(ret) : static_cast<T>(m_exp(ret));
}
template <typename T> __device__ T dens_weibull(T x, T shape, T scale, bool lg) {   // :326-340
  T ret;
  if (x < 0) ret = -Lim<T>::inf();
  else ret = m_log(shape) + (shape - 1) * m_log(x) - shape * m_log(scale) - m_pow(x / scale, shape);
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_log_normal(T x, T mulog, T sdlog, bool lg) {    // :342-353
  T ret;
  if (x < 0) ret = -Lim<T>::inf();
  else ret = dens_normal<T>(m_log(x), mulog, sdlog, true) - m_log(x);
  return maybe_log(ret, lg);
}
template <typename T> __device__ T dens_zi_wrap(T x, T pi0, T d, bool lg) {         // :355-409
  T ret;
  if (pi0 == 0) ret = d;
  else if (pi0 == 1) ret = x == 0 ? 0 : -Lim<T>::inf();
  else if (x == 0) ret = m_log(pi0 + (1 - pi0) * m_exp(d));
  else ret = m_log(1 - pi0) + d;
  return maybe_log(ret, lg);
}

}  // namespace mb200
