// sample_dispatch.h -- host-callable launchers, one per group of samplers.
// The groups only exist to compile the kernel instantiations in parallel.
#pragma once
#include <cuda_runtime.h>
#include "launch_args.h"

namespace mb200 {
// Each returns cudaErrorInvalidValue if `dist` is not in its group.
cudaError_t launch_sample_group_a(int dist, int real_type, int out_f32, bool validate_only,
                                  const SampleArgs& a, int sm_count, cudaStream_t st);
cudaError_t launch_sample_group_b(int dist, int real_type, int out_f32, bool validate_only,
                                  const SampleArgs& a, int sm_count, cudaStream_t st);
cudaError_t launch_sample_group_c(int dist, int real_type, int out_f32, bool validate_only,
                                  const SampleArgs& a, int sm_count, cudaStream_t st);
bool dist_in_group_a(int dist);
// lgamma-of-integers table (samplers.cuh: m_lgamma): one installer per
// translation unit that uses it, one builder
cudaError_t install_lgamma_table_group_b(const double* d, const float* f);
cudaError_t install_lgamma_table_group_c(const double* d, const float* f);
cudaError_t install_lgamma_table_density(const double* d, const float* f);
cudaError_t launch_lgamma_table_build(double* d, float* f, int n, cudaStream_t st);
bool dist_in_group_b(int dist);
}  // namespace mb200
