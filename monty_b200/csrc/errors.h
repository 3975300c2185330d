// errors.h -- host-side formatting of the reference's error messages.
//
// Device code cannot throw, so kernels only record (first offending stream,
// code); the host re-derives the values the reference would have printed and
// formats the identical text.  Derived values (q = 1 - p, prob = size /
// (size + mu), scale = 1 / rate, ...) are recomputed in the same real_type
// arithmetic the reference uses, so "%g" prints the same digits.
#pragma once
#include <cmath>
#include <cstdio>
#include <string>

#include "../../include/monty_b200.h"

namespace mb200 {

// parameter names, ref: src/random.cpp:399-882 (the name_a.. arguments)
inline const char* const* dist_param_names(int dist) {
  static const char* const names[MB200_DIST_COUNT][MB200_MAX_PARAMS] = {
    {nullptr, nullptr, nullptr, nullptr},        // real
    {"min", "max", nullptr, nullptr},
    {"rate", nullptr, nullptr, nullptr},
    {"mean", nullptr, nullptr, nullptr},
    {"mean", "sd", nullptr, nullptr},
    {"mean", "sd", nullptr, nullptr},
    {"mean", "sd", nullptr, nullptr},
    {"size", "prob", nullptr, nullptr},
    {"lambda", nullptr, nullptr, nullptr},
    {"shape", "scale", nullptr, nullptr},
    {"shape", "rate", nullptr, nullptr},
    {"a", "b", nullptr, nullptr},
    {"size", "prob", nullptr, nullptr},
    {"size", "mu", nullptr, nullptr},
    {"n1", "n2", "k", nullptr},
    {"mean", "sd", "min", "max"},
    {"location", "scale", nullptr, nullptr},
    {"shape", "scale", nullptr, nullptr},
    {"meanlog", "sdlog", nullptr, nullptr},
    {"size", "prob", "rho", nullptr},
    {"size", "a", "b", nullptr},
    {"pi0", "lambda", nullptr, nullptr},
    {"pi0", "size", "prob", nullptr},
    {"pi0", "size", "mu", nullptr},
  };
  return names[dist];
}

inline int dist_n_params(int dist) {
  static const int np[MB200_DIST_COUNT] = {
    0, 2, 1, 1, 2, 2, 2, 2, 1, 2, 2, 2, 2, 2, 3, 4, 2, 2, 2, 3, 3, 2, 3, 3};
  return (dist >= 0 && dist < MB200_DIST_COUNT) ? np[dist] : -1;
}

// Which samplers have a validate() in the reference at all.
inline bool dist_validates(int dist) {
  switch (dist) {
  case MB200_DIST_BINOMIAL: case MB200_DIST_POISSON: case MB200_DIST_GAMMA_SCALE:
  case MB200_DIST_GAMMA_RATE: case MB200_DIST_BETA: case MB200_DIST_NEGATIVE_BINOMIAL_PROB:
  case MB200_DIST_NEGATIVE_BINOMIAL_MU: case MB200_DIST_HYPERGEOMETRIC: case MB200_DIST_WEIBULL:
  case MB200_DIST_LOG_NORMAL: case MB200_DIST_BETA_BINOMIAL_PROB: case MB200_DIST_BETA_BINOMIAL_AB:
  case MB200_DIST_ZI_POISSON: case MB200_DIST_ZI_NEGATIVE_BINOMIAL_PROB:
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU:
    return true;
  default:
    return false;
  }
}

template <typename T>
std::string format_param_error_t(int dist, const double* par) {
  char buf[256];
  const T a = static_cast<T>(par[0]), b = static_cast<T>(par[1]), c = static_cast<T>(par[2]);
  switch (dist) {
  case MB200_DIST_BINOMIAL: {                      // hdr/binomial.hpp:204-209
    const T n = std::round(a);
    snprintf(buf, sizeof buf, "Invalid call to binomial with n = %.0f, p = %g, q = %g",
             (double)n, (double)b, (double)(1 - b));
    break;
  }
  case MB200_DIST_POISSON:                         // hdr/poisson.hpp:17-23
    snprintf(buf, sizeof buf, "Invalid call to Poisson with lambda = %g", (double)a);
    break;
  case MB200_DIST_GAMMA_SCALE:                     // hdr/gamma.hpp:25-31
    snprintf(buf, sizeof buf, "Invalid call to gamma with shape = %g, scale = %g", (double)a, (double)b);
    break;
  case MB200_DIST_GAMMA_RATE: {
    const T scale = 1 / b;
    snprintf(buf, sizeof buf, "Invalid call to gamma with shape = %g, scale = %g", (double)a, (double)scale);
    break;
  }
  case MB200_DIST_BETA: {                          // gamma_scale(a, 1) then gamma_scale(b, 1)
    const T shape = (a < 0.0) ? a : b;
    snprintf(buf, sizeof buf, "Invalid call to gamma with shape = %g, scale = %g", (double)shape, 1.0);
    break;
  }
  case MB200_DIST_NEGATIVE_BINOMIAL_PROB:          // hdr/negative_binomial.hpp:16-25
    snprintf(buf, sizeof buf, "Invalid call to negative_binomial with size = %g, prob = %g",
             (double)a, (double)b);
    break;
  case MB200_DIST_NEGATIVE_BINOMIAL_MU: {
    const T prob = a / (a + b);
    snprintf(buf, sizeof buf, "Invalid call to negative_binomial with size = %g, prob = %g",
             (double)a, (double)prob);
    break;
  }
  case MB200_DIST_HYPERGEOMETRIC:                  // hdr/hypergeometric.hpp:21-27
    snprintf(buf, sizeof buf, "Invalid call to hypergeometric with n1 = %.0f, n2 = %.0f, k = %.0f",
             (double)std::round(a), (double)std::round(b), (double)std::round(c));
    break;
  case MB200_DIST_WEIBULL:                         // hdr/weibull.hpp:16-24
    snprintf(buf, sizeof buf, "Invalid call to Weibull with shape = %g, scale = %g", (double)a, (double)b);
    break;
  case MB200_DIST_LOG_NORMAL:                      // hdr/log_normal.hpp:14-22
    snprintf(buf, sizeof buf, "Invalid call to log_normal with meanlog = %g, sdlog = %g",
             (double)a, (double)b);
    break;
  case MB200_DIST_BETA_BINOMIAL_AB:                // hdr/beta_binomial.hpp:15-23
    snprintf(buf, sizeof buf, "Invalid call to beta_binomial with size = %g, a = %g, b = %g",
             (double)a, (double)b, (double)c);
    break;
  case MB200_DIST_BETA_BINOMIAL_PROB: {
    const T aa = b * (1 / c - 1);
    const T bb = (1 - b) * (1 / c - 1);
    snprintf(buf, sizeof buf, "Invalid call to beta_binomial with size = %g, a = %g, b = %g",
             (double)a, (double)aa, (double)bb);
    break;
  }
  case MB200_DIST_ZI_POISSON:                      // hdr/zi_poisson.hpp:15-23
    snprintf(buf, sizeof buf, "Invalid call to zi_poisson with pi0 = %g, lambda = %g", (double)a, (double)b);
    break;
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_PROB:       // hdr/zi_negative_binomial.hpp:15-27
    snprintf(buf, sizeof buf, "Invalid call to zi_negative_binomial with pi0 = %g, size = %g, prob = %g",
             (double)a, (double)b, (double)c);
    break;
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU: {
    const T prob = b / (b + c);
    snprintf(buf, sizeof buf, "Invalid call to zi_negative_binomial with pi0 = %g, size = %g, prob = %g",
             (double)a, (double)b, (double)prob);
    break;
  }
  default:
    snprintf(buf, sizeof buf, "Invalid parameters for distribution %d", dist);
  }
  return buf;
}

inline std::string format_param_error(int dist, int real_type, const double* par) {
  return real_type == MB200_REAL_F32 ? format_param_error_t<float>(dist, par)
                                     : format_param_error_t<double>(dist, par);
}

// data-dependent errors raised inside a composition (codes of samplers.cuh)
inline std::string format_dynamic_error(int code, const double* vals) {
  char buf[256];
  if (code == 2) {
    snprintf(buf, sizeof buf, "Invalid call to binomial with n = %.0f, p = %g, q = %g",
             vals[0], vals[1], vals[2]);
  } else if (code == 3) {
    snprintf(buf, sizeof buf, "Invalid call to Poisson with lambda = %g", vals[0]);
  } else if (code == 4) {   // hdr/cauchy.hpp:21
    snprintf(buf, sizeof buf, "Can't use Cauchy distribution deterministically; it has no mean");
  } else if (code == 7) {   // hdr/binomial.hpp:219-221
    snprintf(buf, sizeof buf, "Invalid call to binomial with n = %f", vals[0]);
  } else {
    snprintf(buf, sizeof buf, "sampler failed with code %d", code);
  }
  return buf;
}

}  // namespace mb200
