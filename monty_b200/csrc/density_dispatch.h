// density_dispatch.h -- host-callable launchers of density_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include "launch_args.h"

namespace mb200 {
// upper bound of the number of per-block partials a launch writes
int density_max_partials(int sm_count);
// *grid_out = number of partials written to a.partial (when it is non-null)
cudaError_t launch_density(int dens, int real_type, const DensityArgs& a, int sm_count, int* grid_out, cudaStream_t st);
cudaError_t launch_reduce_partials(const double* partial, int n, double* sum, cudaStream_t st);
}  // namespace mb200
