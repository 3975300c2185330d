// density_kernels.cu -- elementwise (log-)density kernels with a fused,
// deterministic block/grid sum (the log-likelihood of SURVEY config C5).
//
// ref: src/density.cpp:6-254: `for i < n: ret[i] = density::<d>(x[i], p[i]...)`
// with every argument indexed by i (no recycling); x and the size-like
// arguments of the count distributions arrive as int32, the rest as double.
//
// One grid-stride pass: thread t evaluates elements t, t + stride, ... and
// keeps a running double sum; the block reduces with warp shuffles and writes
// one partial per block; reduce_partials_kernel adds the partials in a fixed
// order, so the sum is reproducible run to run for a given launch geometry.
#include "../../include/monty_b200.h"
#include "density.cuh"
#include "density_dispatch.h"
#include "launch_args.h"

namespace mb200 {

constexpr int kDensBlock = 256;

template <typename T>
__device__ __forceinline__ T ld_arg(const void* p, int is_i32, unsigned long long i) {
  return is_i32 ? static_cast<T>(static_cast<const int*>(p)[i])
                : static_cast<T>(static_cast<const double*>(p)[i]);
}

// number of parameters of density DENS (ref: src/density.cpp signatures)
template <int DENS> struct DensNP {
  static constexpr int value =
    (DENS == MB200_DENS_POISSON || DENS == MB200_DENS_EXPONENTIAL_RATE || DENS == MB200_DENS_EXPONENTIAL_MEAN) ? 1 :
    (DENS == MB200_DENS_BETA_BINOMIAL_PROB || DENS == MB200_DENS_BETA_BINOMIAL_AB || DENS == MB200_DENS_HYPERGEOMETRIC ||
     DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU || DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB) ? 3 :
    (DENS == MB200_DENS_TRUNCATED_NORMAL) ? 4 : 2;
};

template <int DENS, typename T>
__device__ __forceinline__ T density_compute(T x, const T* p, bool lg) {
  switch (DENS) {
  case MB200_DENS_BINOMIAL: return dens_binomial<T>(x, p[0], p[1], lg);
  case MB200_DENS_NORMAL: return dens_normal<T>(x, p[0], p[1], lg);
  case MB200_DENS_NEGATIVE_BINOMIAL_MU: return dens_negative_binomial_mu<T>(x, p[0], p[1], lg);
  case MB200_DENS_NEGATIVE_BINOMIAL_PROB: return dens_negative_binomial_prob<T>(x, p[0], p[1], lg);
  case MB200_DENS_BETA_BINOMIAL_PROB: return dens_beta_binomial_prob<T>(x, p[0], p[1], p[2], lg);
  case MB200_DENS_BETA_BINOMIAL_AB: return dens_beta_binomial_ab<T>(x, p[0], p[1], p[2], lg);
  case MB200_DENS_POISSON: return dens_poisson<T>(x, p[0], lg);
  case MB200_DENS_UNIFORM: return dens_uniform<T>(x, p[0], p[1], lg);
  case MB200_DENS_EXPONENTIAL_RATE: return dens_exponential_rate<T>(x, p[0], lg);
  case MB200_DENS_EXPONENTIAL_MEAN: return dens_exponential_mean<T>(x, p[0], lg);
  case MB200_DENS_GAMMA_RATE: return dens_gamma_rate<T>(x, p[0], p[1], lg);
  case MB200_DENS_GAMMA_SCALE: return dens_gamma_scale<T>(x, p[0], p[1], lg);
  case MB200_DENS_BETA: return dens_beta<T>(x, p[0], p[1], lg);
  case MB200_DENS_HYPERGEOMETRIC: return dens_hypergeometric<T>(x, p[0], p[1], p[2], lg);
  case MB200_DENS_LOG_NORMAL: return dens_log_normal<T>(x, p[0], p[1], lg);
  case MB200_DENS_TRUNCATED_NORMAL: return dens_truncated_normal<T>(x, p[0], p[1], p[2], p[3], lg);
  case MB200_DENS_CAUCHY: return dens_cauchy<T>(x, p[0], p[1], lg);
  case MB200_DENS_WEIBULL: return dens_weibull<T>(x, p[0], p[1], lg);
  case MB200_DENS_ZI_POISSON:
    return dens_zi_wrap<T>(x, p[0], dens_poisson<T>(x, p[1], true), lg);
  case MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU:
    return dens_zi_wrap<T>(x, p[0], dens_negative_binomial_mu<T>(x, p[1], p[2], true), lg);
  case MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB:
    return dens_zi_wrap<T>(x, p[0], dens_negative_binomial_prob<T>(x, p[1], p[2], true), lg);
  default: return 0;
  }
}

// Four observations per thread per trip, all their loads issued before the
// first use: the count densities became short once lgamma of integers is a
// table look-up, and with one element in flight per thread the kernel waited
// on HBM latency (ncu: long-scoreboard stall 11.9 per issue,
// profiles/r01_density_poisson_table_ncu_summary.txt).  Element order per
// thread and the reduction tree are fixed, so the sum stays reproducible.
// Only for the light densities: the lgamma-heavy ones (non-integer arguments)
// are FP64-pipe bound and lose occupancy to the extra registers.
template <int DENS> struct DensUnroll {
  static constexpr int value =
    (DENS == MB200_DENS_POISSON || DENS == MB200_DENS_NORMAL || DENS == MB200_DENS_UNIFORM ||
     DENS == MB200_DENS_EXPONENTIAL_RATE || DENS == MB200_DENS_EXPONENTIAL_MEAN || DENS == MB200_DENS_CAUCHY ||
     DENS == MB200_DENS_LOG_NORMAL || DENS == MB200_DENS_WEIBULL || DENS == MB200_DENS_ZI_POISSON) ? 4 : 1;
};

template <int DENS, typename T>
__global__ void __launch_bounds__(kDensBlock)
density_kernel(const DensityArgs a) {
  constexpr int NPD = DensNP<DENS>::value;
  constexpr int kDensUnroll = DensUnroll<DENS>::value;
  double acc = 0.0;
  const bool lg = a.log != 0;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kDensBlock;
  for (unsigned long long i0 = static_cast<unsigned long long>(blockIdx.x) * kDensBlock + threadIdx.x;
       i0 < a.n; i0 += stride * kDensUnroll) {
    T x[kDensUnroll], p[kDensUnroll][NPD];
#pragma unroll
    for (int u = 0; u < kDensUnroll; ++u) {
      const unsigned long long i = i0 + u * stride;
      if (i < a.n) {
        x[u] = ld_arg<T>(a.x, a.x_is_i32, i);
#pragma unroll
        for (int k = 0; k < NPD; ++k) p[u][k] = ld_arg<T>(a.par[k], a.par_is_i32[k], i);
      }
    }
#pragma unroll
    for (int u = 0; u < kDensUnroll; ++u) {
      const unsigned long long i = i0 + u * stride;
      if (i < a.n) {
        const double v = static_cast<double>(density_compute<DENS, T>(x[u], p[u], lg));
        if (a.out) a.out[i] = v;
        acc += v;
      }
    }
  }
  if (a.partial) {
    __shared__ double warp_sum[kDensBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kDensBlock / 32; ++w) s += warp_sum[w];
      a.partial[blockIdx.x] = s;
    }
  }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ sum) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *sum = sh[0];
}

int density_grid(unsigned long long n, int sm_count) {
  const unsigned long long need = (n + kDensBlock - 1) / kDensBlock;
  const unsigned long long cap = static_cast<unsigned long long>(sm_count) * 8;
  return static_cast<int>(need < cap ? (need ? need : 1) : cap);
}

template <int DENS>
static cudaError_t launch_density_t(int real_type, const DensityArgs& a, int grid, cudaStream_t st) {
  if (real_type == MB200_REAL_F32) density_kernel<DENS, float><<<grid, kDensBlock, 0, st>>>(a);
  else density_kernel<DENS, double><<<grid, kDensBlock, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_density(int dens, int real_type, const DensityArgs& a, int grid, cudaStream_t st) {
  switch (dens) {
#define C(id) case id: return launch_density_t<id>(real_type, a, grid, st)
    C(MB200_DENS_BINOMIAL); C(MB200_DENS_NORMAL); C(MB200_DENS_NEGATIVE_BINOMIAL_MU);
    C(MB200_DENS_NEGATIVE_BINOMIAL_PROB); C(MB200_DENS_BETA_BINOMIAL_PROB);
    C(MB200_DENS_BETA_BINOMIAL_AB); C(MB200_DENS_POISSON); C(MB200_DENS_UNIFORM);
    C(MB200_DENS_EXPONENTIAL_RATE); C(MB200_DENS_EXPONENTIAL_MEAN); C(MB200_DENS_GAMMA_RATE);
    C(MB200_DENS_GAMMA_SCALE); C(MB200_DENS_BETA); C(MB200_DENS_HYPERGEOMETRIC);
    C(MB200_DENS_LOG_NORMAL); C(MB200_DENS_TRUNCATED_NORMAL); C(MB200_DENS_CAUCHY);
    C(MB200_DENS_WEIBULL); C(MB200_DENS_ZI_POISSON); C(MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU);
    C(MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB);
#undef C
  default: return cudaErrorInvalidValue;
  }
}

cudaError_t install_lgamma_table_density(const double* d, const float* f) { return lgt_install(d, f); }

cudaError_t launch_reduce_partials(const double* partial, int n, double* sum, cudaStream_t st) {
  reduce_partials_kernel<<<1, 256, 0, st>>>(partial, n, sum);
  return cudaGetLastError();
}

}  // namespace mb200
