// draw_dist.cuh -- one draw of distribution `dist` (MB200_DIST_*) through the device-side
// monty::random API on the caller's generator state, whatever its type: the dispatch the
// example kernels share.
#pragma once
#include "../include/monty_b200/random.cuh"
#include "../include/monty_b200.h"

namespace mb200_examples {

template <typename T, typename G>
__device__ __forceinline__ T draw_dist(G& g, int dist, const T* p) {
  namespace mr = monty::random;
  T v;
  switch (dist) {
  case MB200_DIST_REAL: v = mr::random_real<T>(g); break;
  case MB200_DIST_UNIFORM: v = mr::uniform<T>(g, p[0], p[1]); break;
  case MB200_DIST_EXPONENTIAL_RATE: v = mr::exponential_rate<T>(g, p[0]); break;
  case MB200_DIST_EXPONENTIAL_MEAN: v = mr::exponential_mean<T>(g, p[0]); break;
  case MB200_DIST_NORMAL_BOX_MULLER: v = mr::normal<T>(g, p[0], p[1]); break;
  case MB200_DIST_NORMAL_POLAR: v = mr::normal<T, mr::algorithm::normal::polar>(g, p[0], p[1]); break;
  case MB200_DIST_NORMAL_ZIGGURAT: v = mr::normal<T, mr::algorithm::normal::ziggurat>(g, p[0], p[1]); break;
  case MB200_DIST_BINOMIAL: v = mr::binomial<T>(g, p[0], p[1]); break;
  case MB200_DIST_POISSON: v = mr::poisson<T>(g, p[0]); break;
  case MB200_DIST_GAMMA_SCALE: v = mr::gamma_scale<T>(g, p[0], p[1]); break;
  case MB200_DIST_GAMMA_RATE: v = mr::gamma_rate<T>(g, p[0], p[1]); break;
  case MB200_DIST_BETA: v = mr::beta<T>(g, p[0], p[1]); break;
  case MB200_DIST_NEGATIVE_BINOMIAL_PROB: v = mr::negative_binomial_prob<T>(g, p[0], p[1]); break;
  case MB200_DIST_NEGATIVE_BINOMIAL_MU: v = mr::negative_binomial_mu<T>(g, p[0], p[1]); break;
  case MB200_DIST_HYPERGEOMETRIC: v = mr::hypergeometric<T>(g, p[0], p[1], p[2]); break;
  case MB200_DIST_TRUNCATED_NORMAL: v = mr::truncated_normal<T>(g, p[0], p[1], p[2], p[3]); break;
  case MB200_DIST_CAUCHY: v = mr::cauchy<T>(g, p[0], p[1]); break;
  case MB200_DIST_WEIBULL: v = mr::weibull<T>(g, p[0], p[1]); break;
  case MB200_DIST_LOG_NORMAL: v = mr::log_normal<T>(g, p[0], p[1]); break;
  case MB200_DIST_BETA_BINOMIAL_PROB: v = mr::beta_binomial_prob<T>(g, p[0], p[1], p[2]); break;
  case MB200_DIST_BETA_BINOMIAL_AB: v = mr::beta_binomial_ab<T>(g, p[0], p[1], p[2]); break;
  case MB200_DIST_ZI_POISSON: v = mr::zi_poisson<T>(g, p[0], p[1]); break;
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_PROB: v = mr::zi_negative_binomial_prob<T>(g, p[0], p[1], p[2]); break;
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU: v = mr::zi_negative_binomial_mu<T>(g, p[0], p[1], p[2]); break;
  default: v = 0; break;
  }
  return v;
}

}  // namespace mb200_examples
