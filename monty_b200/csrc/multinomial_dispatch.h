// multinomial_dispatch.h -- host-callable launcher of multinomial_kernel.cu
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "launch_args.h"

namespace mb200 {
cudaError_t launch_multinomial(uint64_t* state, unsigned long long n_streams,
                               unsigned long long n_samples,
                               const double* size, unsigned size_stride,
                               const double* prob, int prob_len, unsigned prob_stride,
                               double* out, ErrRec* err, cudaStream_t st);
}  // namespace mb200
