// deterministic_kernel.cu -- the reference's "deterministic" generator mode:
// every sampler returns the expectation of its distribution and does not
// touch the stream (host-only `rng_state.deterministic` branches of the
// reference; ref per case below).  Two faithful quirks are kept:
//   * random_real has no deterministic branch (hdr/generator.hpp:154-159):
//     monty_random_real still draws;
//   * truncated_normal with both bounds infinite draws a Box-Muller normal
//     before looking at the flag (hdr/truncated_normal.hpp:71-77).
// Not a throughput path: one thread per stream, runtime switch on the
// distribution, each stream writes its own n_samples copies.
#include "../../include/monty_b200.h"
#include "deterministic_dispatch.h"
#include "density.cuh"
#include "sample_kernel.cuh"

namespace mb200 {

// hdr/truncated_normal.hpp:55-63
template <typename T>
__device__ T truncated_normal_standard_mean(T min, T max) {
  const T m_1_sqrt_2 = 0.70710678118654746172;
  const double z_min = 0.5 * (1 + (sizeof(T) == 4 ? erff(min * m_1_sqrt_2) : erf(min * m_1_sqrt_2)));
  const double z_max = 0.5 * (1 + (sizeof(T) == 4 ? erff(max * m_1_sqrt_2) : erf(max * m_1_sqrt_2)));
  const T d_min = dens_normal<T>(min, 0, 1, false);
  const T d_max = dens_normal<T>(max, 0, 1, false);
  return (d_min - d_max) / (z_max - z_min);
}

__device__ __forceinline__ double m_tgamma(double x) { return tgamma(x); }
__device__ __forceinline__ float m_tgamma(float x) { return tgammaf(x); }

// Returns the deterministic value; err != 0 reports a validation failure.
// `rng` is used (and `used` set) only by the two quirks above.
template <typename T>
__device__ T deterministic_value(int dist, const T* p, Rng& rng, bool& used, int& err, double* vals) {
  const T a = p[0], b = p[1], c = p[2], d = p[3];
  switch (dist) {
  case MB200_DIST_REAL:
    used = true;
    return random_real<T>(rng);
  case MB200_DIST_UNIFORM: return (b - a) / 2 + a;                          // hdr/uniform.hpp:29-33
  case MB200_DIST_EXPONENTIAL_RATE: return static_cast<T>(1) / a;           // hdr/exponential.hpp:24-27
  case MB200_DIST_EXPONENTIAL_MEAN: return static_cast<T>(1) * a;
  case MB200_DIST_NORMAL_BOX_MULLER:
  case MB200_DIST_NORMAL_POLAR:
  case MB200_DIST_NORMAL_ZIGGURAT: return static_cast<T>(0) * b + a;        // hdr/normal.hpp:86-91
  case MB200_DIST_BINOMIAL: {                                               // hdr/binomial.hpp:213-227
    T n = a;
    if (n < 0) {
      if (n * n < Lim<T>::eps) {
        n = m_round(n);
      } else {
        err = kErrBinomialDetN; vals[0] = n;
        return 0;
      }
    }
    if (DistBinomial<T>::invalid(n, b)) { err = kErrParam; vals[0] = n; vals[1] = b; vals[2] = 1 - b; return 0; }
    return n > 0 && b > 0 ? n * b : 0;
  }
  case MB200_DIST_POISSON:                                                  // hdr/poisson.hpp:221-229
    if (DistPoisson<T>::invalid(a)) { err = kErrParam; return 0; }
    return a == 0 ? 0 : a;
  case MB200_DIST_GAMMA_SCALE:
  case MB200_DIST_GAMMA_RATE: {                                             // hdr/gamma.hpp:86-101, 113-121
    const T scale = dist == MB200_DIST_GAMMA_RATE ? 1 / b : b;
    if (a < 0.0 || scale < 0.0) { err = kErrParam; return 0; }
    if (a == 0 || scale == 0) return 0;
    return a * scale;
  }
  case MB200_DIST_BETA: return a / (a + b);                                 // hdr/beta.hpp:21-23
  case MB200_DIST_NEGATIVE_BINOMIAL_PROB:
  case MB200_DIST_NEGATIVE_BINOMIAL_MU: {                                   // hdr/negative_binomial.hpp:30-52
    const T prob = dist == MB200_DIST_NEGATIVE_BINOMIAL_MU ? a / (a + b) : b;
    if (DistNegativeBinomial<T, false>::invalid(a, prob)) { err = kErrParam; return 0; }
    return a > 0 ? (1 - prob) * a / prob : 0;
  }
  case MB200_DIST_HYPERGEOMETRIC: {                                         // hdr/hypergeometric.hpp:316-321 (unrounded)
    const T n = a + b;
    if (a < 0 || b < 0 || c < 0 || c > n) { err = kErrHypergeometricDet; return 0; }
    return a * c / n;
  }
  case MB200_DIST_TRUNCATED_NORMAL: {                                       // hdr/truncated_normal.hpp:65-95
    const T lo = (c - a) / b, hi = (d - a) / b;
    T z;
    if (lo == -Lim<T>::inf() && hi == Lim<T>::inf()) {
      used = true;
      z = std_normal_box_muller<T>(rng);
    } else {
      z = truncated_normal_standard_mean<T>(lo, hi);
    }
    return z * b + a;
  }
  case MB200_DIST_CAUCHY: err = kErrCauchyDeterministic; return 0;          // hdr/cauchy.hpp:19-23
  case MB200_DIST_WEIBULL:                                                  // hdr/weibull.hpp:30-35
    if (a < 0.0 || b < 0.0) { err = kErrParam; return 0; }
    return b * m_tgamma(1 + 1 / a);
  case MB200_DIST_LOG_NORMAL:                                               // hdr/log_normal.hpp:30-35
    if (!is_finite(a) || !is_finite(b) || b <= 0) { err = kErrParam; return 0; }
    return m_exp(a + (sizeof(T) == 4 ? static_cast<T>(m_pow(static_cast<double>(b), 2.0)) : static_cast<T>(m_pow(static_cast<double>(b), 2.0))) / 2);
  case MB200_DIST_BETA_BINOMIAL_PROB:
  case MB200_DIST_BETA_BINOMIAL_AB: {                                       // hdr/beta_binomial.hpp:27-49
    T aa = b, bb = c;
    if (dist == MB200_DIST_BETA_BINOMIAL_PROB) { aa = b * (1 / c - 1); bb = (1 - b) * (1 / c - 1); }
    if (!is_finite(a) || !is_finite(aa) || !is_finite(bb) || a < 0 || aa <= 0 || bb <= 0) { err = kErrParam; return 0; }
    return a * aa / (aa + bb);
  }
  case MB200_DIST_ZI_POISSON: {                                             // hdr/zi_poisson.hpp:27-36
    const bool bad = a < 0 || a > 1 || b < 0 || !is_finite(a) || (a < 1 && !is_finite(b));
    if (bad) { err = kErrParam; return 0; }
    return a < 1 ? (1 - a) * b : 0;
  }
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_PROB:
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU: {                                // hdr/zi_negative_binomial.hpp:32-58
    const T prob = dist == MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU ? b / (b + c) : c;
    const bool bad = a < 0 || a > 1 || b < 0 || prob < 0 || prob > 1 || !is_finite(a) ||
      (a < 1 && is_nan(prob) && b > 0) || (a < 1 && !is_finite(b)) || (a < 1 && b > 0 && prob == 0);
    if (bad) { err = kErrParam; return 0; }
    return (a < 1 && b > 0) ? (1 - a) * (1 - prob) * b / prob : 0;
  }
  default: err = kErrParam; return 0;
  }
}

template <typename T>
__global__ void __launch_bounds__(128)
deterministic_kernel(int dist, int n_par, const SampleArgs a) {
  const unsigned long long stream = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (stream >= a.n_streams) return;
  T par[4] = {0, 0, 0, 0};
  for (int k = 0; k < n_par; ++k) par[k] = static_cast<T>(a.par[k][a.pstride[k] ? stream : 0]);
  Rng rng;
  rng.s = load_state(a.state, stream);
  double* out = static_cast<double*>(a.out) + stream * a.n_samples;
  bool used = false;
  for (unsigned long long j = 0; j < a.n_samples; ++j) {
    int err = 0;
    double vals[3] = {0, 0, 0};
    const T y = deterministic_value<T>(dist, par, rng, used, err, vals);
    if (err) {
      report_error(a.err, stream, err, vals);
      for (; j < a.n_samples; ++j) out[j] = out_nan<double>();
      break;
    }
    out[j] = static_cast<double>(y);
  }
  if (used) store_state(a.state, stream, rng.s);
}

// hdr/multinomial.hpp:53-82 with binomial_deterministic in place of the draws
__global__ void __launch_bounds__(128)
multinomial_deterministic_kernel(unsigned long long n_streams, unsigned long long n_samples,
                                 const double* __restrict__ size, unsigned size_stride,
                                 const double* __restrict__ prob, int prob_len, unsigned prob_stride,
                                 double* __restrict__ out, ErrRec* err) {
  const unsigned long long stream = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (stream >= n_streams) return;
  const double* pr = prob + static_cast<unsigned long long>(prob_stride) * stream * prob_len;
  double p_tot0 = 0;
  for (int i = 0; i < prob_len; ++i) {
    if (pr[i] < 0) { report_error(err, stream, 5, nullptr); return; }
    p_tot0 += pr[i];
  }
  if (p_tot0 == 0) { report_error(err, stream, 6, nullptr); return; }
  Rng rng{};
  for (unsigned long long j = 0; j < n_samples; ++j) {
    double* y = out + static_cast<unsigned long long>(prob_len) * (n_samples * stream + j);
    double sz = size[size_stride ? stream : 0];
    double p_tot = p_tot0;
    for (int i = 0; i < prob_len - 1; ++i) {
      const double p = pr[i];
      if (p > 0) {
        const double q = p / p_tot;
        const double par[4] = {sz, q < 1.0 ? q : 1.0, 0, 0};
        bool used = false;
        int e = 0;
        double vals[3] = {0, 0, 0};
        const double v = deterministic_value<double>(MB200_DIST_BINOMIAL, par, rng, used, e, vals);
        if (e) { report_error(err, stream, e == kErrParam ? kErrInnerBinomial : e, vals); return; }
        y[i] = v;
        sz -= v;
        p_tot -= p;
      } else {
        y[i] = 0;
      }
    }
    y[prob_len - 1] = sz;
  }
}

cudaError_t launch_deterministic(int dist, int n_par, int real_type, const SampleArgs& a, cudaStream_t st) {
  if (a.n_streams == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((a.n_streams + 127) / 128);
  if (real_type == MB200_REAL_F32) deterministic_kernel<float><<<grid, 128, 0, st>>>(dist, n_par, a);
  else deterministic_kernel<double><<<grid, 128, 0, st>>>(dist, n_par, a);
  return cudaGetLastError();
}

cudaError_t launch_multinomial_deterministic(unsigned long long n_streams, unsigned long long n_samples,
                                             const double* size, unsigned size_stride,
                                             const double* prob, int prob_len, unsigned prob_stride,
                                             double* out, ErrRec* err, cudaStream_t st) {
  if (n_streams == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((n_streams + 127) / 128);
  multinomial_deterministic_kernel<<<grid, 128, 0, st>>>(n_streams, n_samples, size, size_stride,
                                                        prob, prob_len, prob_stride, out, err);
  return cudaGetLastError();
}

}  // namespace mb200
