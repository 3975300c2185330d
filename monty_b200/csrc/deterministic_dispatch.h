// deterministic_dispatch.h -- host-callable launchers of deterministic_kernel.cu
#pragma once
#include <cuda_runtime.h>
#include "launch_args.h"

namespace mb200 {
cudaError_t launch_deterministic(int dist, int n_par, int real_type, const SampleArgs& a, cudaStream_t st);
cudaError_t launch_multinomial_deterministic(unsigned long long n_streams, unsigned long long n_samples,
                                             const double* size, unsigned size_stride,
                                             const double* prob, int prob_len, unsigned prob_stride,
                                             double* out, ErrRec* err, cudaStream_t st);
}  // namespace mb200
