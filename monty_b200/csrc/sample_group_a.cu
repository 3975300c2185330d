// sample_group_a.cu -- kernels for real, uniform, exponential, the three
// normals, cauchy, weibull, log_normal.
#include "../../include/monty_b200.h"
#include "sample_dispatch.h"
#include "sample_kernel.cuh"
#include "ziggurat_kernel.cuh"

#include <cstdlib>

namespace mb200 {
template <typename T> using DExpRate = DistExponential<T, false>;
template <typename T> using DExpMean = DistExponential<T, true>;
template <typename T> using DNormBM = DistNormal<T, kBoxMuller>;
template <typename T> using DNormPolar = DistNormal<T, kPolar>;
template <typename T> using DNormZig = DistNormal<T, kZiggurat>;

bool dist_in_group_a(int dist) {
  switch (dist) {
  case MB200_DIST_REAL: case MB200_DIST_UNIFORM: case MB200_DIST_EXPONENTIAL_RATE:
  case MB200_DIST_EXPONENTIAL_MEAN: case MB200_DIST_NORMAL_BOX_MULLER:
  case MB200_DIST_NORMAL_POLAR: case MB200_DIST_NORMAL_ZIGGURAT: case MB200_DIST_CAUCHY:
  case MB200_DIST_WEIBULL: case MB200_DIST_LOG_NORMAL:
    return true;
  default:
    return false;
  }
}

cudaError_t launch_sample_group_a(int dist, int real_type, int out_f32, bool validate_only,
                                  const SampleArgs& a, int sm_count, cudaStream_t st) {
  if (dist == MB200_DIST_NORMAL_ZIGGURAT && !validate_only && ziggurat_kernel_applies(a)) {
    // MB200_ZIG_GENERIC=1 keeps the lock-step kernel for A/B measurements
    static const bool generic = std::getenv("MB200_ZIG_GENERIC") != nullptr;
    if (!generic) return launch_ziggurat(real_type, out_f32, a, sm_count, st);
  }
#define MB200_CASE(id, D) case id: return launch_sample_any<D>(real_type, out_f32, validate_only, a, sm_count, st)
  switch (dist) {
    MB200_CASE(MB200_DIST_REAL, DistReal);
    MB200_CASE(MB200_DIST_UNIFORM, DistUniform);
    MB200_CASE(MB200_DIST_EXPONENTIAL_RATE, DExpRate);
    MB200_CASE(MB200_DIST_EXPONENTIAL_MEAN, DExpMean);
    MB200_CASE(MB200_DIST_NORMAL_BOX_MULLER, DNormBM);
    MB200_CASE(MB200_DIST_NORMAL_POLAR, DNormPolar);
    MB200_CASE(MB200_DIST_NORMAL_ZIGGURAT, DNormZig);
    MB200_CASE(MB200_DIST_CAUCHY, DistCauchy);
    MB200_CASE(MB200_DIST_WEIBULL, DistWeibull);
    MB200_CASE(MB200_DIST_LOG_NORMAL, DistLogNormal);
  default: return cudaErrorInvalidValue;
  }
#undef MB200_CASE
}
}  // namespace mb200
