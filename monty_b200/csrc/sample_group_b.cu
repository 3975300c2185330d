// sample_group_b.cu -- kernels for binomial, poisson, gamma, beta, negative
// binomial and the zero-inflated Poisson.
#include "../../include/monty_b200.h"
#include "sample_dispatch.h"
#include "sample_kernel.cuh"
#include "sorted_kernel.cuh"

namespace mb200 {
template <typename T> using DGammaScale = DistGamma<T, false>;
template <typename T> using DGammaRate = DistGamma<T, true>;
template <typename T> using DNbProb = DistNegativeBinomial<T, false>;
template <typename T> using DNbMu = DistNegativeBinomial<T, true>;

bool dist_in_group_b(int dist) {
  switch (dist) {
  case MB200_DIST_BINOMIAL: case MB200_DIST_POISSON: case MB200_DIST_GAMMA_SCALE:
  case MB200_DIST_GAMMA_RATE: case MB200_DIST_BETA: case MB200_DIST_NEGATIVE_BINOMIAL_PROB:
  case MB200_DIST_NEGATIVE_BINOMIAL_MU: case MB200_DIST_ZI_POISSON:
    return true;
  default:
    return false;
  }
}

cudaError_t launch_sample_group_b(int dist, int real_type, int out_f32, bool validate_only,
                                  const SampleArgs& a, int sm_count, cudaStream_t st) {
#define MB200_CASE(id, D) case id: return launch_sample_regimes<D>(real_type, out_f32, validate_only, a, sm_count, st)
  switch (dist) {
    MB200_CASE(MB200_DIST_BINOMIAL, DistBinomial);
    MB200_CASE(MB200_DIST_POISSON, DistPoisson);
    MB200_CASE(MB200_DIST_GAMMA_SCALE, DGammaScale);
    MB200_CASE(MB200_DIST_GAMMA_RATE, DGammaRate);
    MB200_CASE(MB200_DIST_BETA, DistBeta);
    MB200_CASE(MB200_DIST_NEGATIVE_BINOMIAL_PROB, DNbProb);
    MB200_CASE(MB200_DIST_NEGATIVE_BINOMIAL_MU, DNbMu);
    MB200_CASE(MB200_DIST_ZI_POISSON, DistZiPoisson);
  default: return cudaErrorInvalidValue;
  }
#undef MB200_CASE
}
cudaError_t install_lgamma_table_group_b(const double* d, const float* f) { return lgt_install(d, f); }
}  // namespace mb200
