// sample_group_c.cu -- kernels for hypergeometric, truncated normal,
// beta-binomial and the zero-inflated negative binomials.
#include "../../include/monty_b200.h"
#include "sample_dispatch.h"
#include "sample_kernel.cuh"
#include "sorted_kernel.cuh"

namespace mb200 {
template <typename T> using DBbProb = DistBetaBinomial<T, true>;
template <typename T> using DBbAb = DistBetaBinomial<T, false>;
template <typename T> using DZinbProb = DistZiNegativeBinomial<T, false>;
template <typename T> using DZinbMu = DistZiNegativeBinomial<T, true>;

cudaError_t launch_sample_group_c(int dist, int real_type, int out_f32, bool validate_only,
                                  const SampleArgs& a, int sm_count, cudaStream_t st) {
#define MB200_CASE(id, D) case id: return launch_sample_regimes<D>(real_type, out_f32, validate_only, a, sm_count, st)
  switch (dist) {
    MB200_CASE(MB200_DIST_HYPERGEOMETRIC, DistHypergeometric);
    MB200_CASE(MB200_DIST_TRUNCATED_NORMAL, DistTruncatedNormal);
    MB200_CASE(MB200_DIST_BETA_BINOMIAL_PROB, DBbProb);
    MB200_CASE(MB200_DIST_BETA_BINOMIAL_AB, DBbAb);
    MB200_CASE(MB200_DIST_ZI_NEGATIVE_BINOMIAL_PROB, DZinbProb);
    MB200_CASE(MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU, DZinbMu);
  default: return cudaErrorInvalidValue;
  }
#undef MB200_CASE
}
cudaError_t install_lgamma_table_group_c(const double* d, const float* f) { return lgt_install(d, f); }
}  // namespace mb200
