// density_dispatch.h -- host-callable launchers of density_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include "launch_args.h"

namespace mb200 {
int density_grid(unsigned long long n, int sm_count);
cudaError_t launch_density(int dens, int real_type, const DensityArgs& a, int grid, cudaStream_t st);
cudaError_t launch_reduce_partials(const double* partial, int n, double* sum, cudaStream_t st);
}  // namespace mb200
