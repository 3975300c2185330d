// rng_dispatch.h -- host-callable launchers of rng_kernels.cu
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mb200 {
cudaError_t launch_apply_poly(uint64_t* state, unsigned long long first, unsigned long long count,
                              const uint64_t poly[4], cudaStream_t st);
cudaError_t launch_ladder(uint64_t* state, unsigned long long anchor, unsigned long long have,
                          unsigned long long n_new, const uint64_t poly[4], cudaStream_t st);
cudaError_t launch_raw(int gen, uint64_t* state, unsigned long long n_streams,
                       unsigned long long n_samples, uint64_t* out, cudaStream_t st);
cudaError_t launch_xoshiro_run(int gen, uint64_t seed, int n, uint64_t* values_dev,
                               uint64_t* state_dev, int* n_words, cudaStream_t st);
}  // namespace mb200
