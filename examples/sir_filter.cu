// sir_filter.cu -- the reference test-suite's stochastic SIR model and its
// bootstrap particle filter (tests/testthat/helper-sir-filter.R:56-117) as a
// dust2-style consumer of the device-side monty::random API.
//
//   sir_step (:99-110)   p_SI = 1 - exp(-beta I / N dt), p_IR = 1 - exp(-gamma dt),
//                        n_SI ~ binomial(S, p_SI), n_IR ~ binomial(I, p_IR)
//   sir_filter (:56-88)  per observation: run the steps, noise ~ exponential(rate),
//                        lambda = cases + noise, log weight = dpois(incidence, lambda),
//                        ll += log(mean(w)) + max, systematic resampling with one
//                        uniform from the filter stream (dust_resample_weight :113-117)
//
// B200 shape: one thread per particle; the particle's generator state and its
// compartments stay in registers across all steps up to the next observation,
// and the exponential draw and the Poisson log-density are fused into the same
// kernel -- one state round trip (32 + 32 B) per observation instead of one
// per draw.  The weight normalisation, prefix sum and gather run as three
// small kernels (the particle count is the only parallelism there).
#include <cuda_runtime.h>

#include <vector>

#include "../include/monty_b200/random.cuh"
#include "examples.h"

namespace {
namespace mr = monty::random;

constexpr int kScanThreads = 1024;

struct SirPars { double N, beta, gamma, exp_noise; };

__global__ void sir_run_kernel(uint64_t* states, size_t n, double* comp /* [4][n] */, SirPars q,
                               double from, double to, double incidence, double* logw) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  auto g = mr::load_state(states, i);
  double S = comp[i], I = comp[n + i], R = comp[2 * n + i], cases = comp[3 * n + i];
  const double dt = 1;
  for (double time = from + dt; time <= to; time += dt) {       // seq(from, to, by = dt)[-1]
    const double p_SI = 1 - exp(-q.beta * I / q.N * dt);
    const double p_IR = 1 - exp(-q.gamma * dt);
    const double n_SI = mr::binomial<double>(g, S, p_SI);
    const double n_IR = mr::binomial<double>(g, I, p_IR);
    const double carried = fmod(time, 1.0) == 0 ? 0 : cases;
    S = S - n_SI;
    I = I + n_SI - n_IR;
    R = R + n_IR;
    cases = carried + n_SI;
  }
  const double noise = mr::exponential_rate<double>(g, q.exp_noise);
  logw[i] = monty::density::poisson<double>(incidence, cases + noise, true);
  mr::store_state(states, i, g);
  comp[i] = S; comp[n + i] = I; comp[2 * n + i] = R; comp[3 * n + i] = cases;
}

__device__ double block_reduce(double v, bool take_max, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = 16; d > 0; d >>= 1) {
    const double o = __shfl_down_sync(0xffffffffu, v, d);
    v = take_max ? fmax(v, o) : v + o;
  }
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? scratch[lane] : (take_max ? -INFINITY : 0.0);
    for (int d = 16; d > 0; d >>= 1) {
      const double o = __shfl_down_sync(0xffffffffu, v, d);
      v = take_max ? fmax(v, o) : v + o;
    }
    if (lane == 0) scratch[0] = v;
  }
  __syncthreads();
  const double r = scratch[0];
  __syncthreads();
  return r;
}

// One block.  w = exp(logw - max), ll += log(mean(w)) + max, cum = cumsum(w / sum(w)),
// u = one uniform from the filter stream.
__global__ void __launch_bounds__(kScanThreads)
sir_weights_kernel(const double* logw, size_t n, double* cum, double* ll, uint64_t* filter_state, double* u_out) {
  __shared__ double scratch[32];
  __shared__ double offs[kScanThreads];
  const size_t per = (n + kScanThreads - 1) / kScanThreads;
  const size_t lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  double m = -INFINITY;
  for (size_t i = lo; i < hi; ++i) m = fmax(m, logw[i]);
  m = block_reduce(m, true, scratch);
  double s = 0;
  for (size_t i = lo; i < hi; ++i) s += exp(logw[i] - m);
  const double total = block_reduce(s, false, scratch);
  double c = 0;
  for (size_t i = lo; i < hi; ++i) c += exp(logw[i] - m) / total;
  offs[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0;
    for (int t = 0; t < kScanThreads; ++t) { const double v = offs[t]; offs[t] = run; run += v; }
    *ll += log(total / static_cast<double>(n)) + m;
    auto g = mr::load_state(filter_state, 0);
    *u_out = mr::random_real<double>(g);
    mr::store_state(filter_state, 0, g);
  }
  __syncthreads();
  double run = offs[threadIdx.x];
  for (size_t i = lo; i < hi; ++i) { run += exp(logw[i] - m) / total; cum[i] = run; }
}

// k = findInterval(u / n + j / n, cum) (number of entries <= x), then gather
__global__ void sir_resample_kernel(const double* cum, size_t n, const double* u, const double* src, double* dst) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (j >= n) return;
  const double nn = static_cast<double>(n);
  const double x = *u / nn + static_cast<double>(j) * (1 / nn);
  size_t lo = 0, hi = n;                       // first index with cum[idx] > x
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (cum[mid] <= x) lo = mid + 1; else hi = mid;
  }
  const size_t k = lo < n ? lo : n - 1;
  for (int c = 0; c < 4; ++c) dst[c * n + j] = src[c * n + k];
}

}  // namespace

extern "C" int mb200_example_sir_filter(mb200_rng* system, mb200_rng* filter,
                                        double N, double I0, double beta, double gamma, double exp_noise,
                                        const double* data_time, const double* data_incidence, size_t n_data,
                                        int resample, double* log_likelihood, double* final_state_host) {
  if (!system || !filter || !log_likelihood) return MB200_ERR_ARG;
  if (cudaSetDevice(mb200_rng_device(system)) != cudaSuccess) return MB200_ERR_CUDA;
  if (mb200_rng_sync(system) != MB200_OK || mb200_rng_sync(filter) != MB200_OK) return MB200_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(mb200_rng_cuda_stream(system));
  const size_t n = mb200_rng_n_streams(system);
  double *comp = nullptr, *comp2 = nullptr, *logw = nullptr, *cum = nullptr, *scalars = nullptr;
  bool ok = cudaMalloc(&comp, 4 * n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&comp2, 4 * n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&logw, n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&cum, n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&scalars, 2 * sizeof(double)) == cudaSuccess;
  if (ok) {
    std::vector<double> init(4 * n, 0.0);
    for (size_t i = 0; i < n; ++i) { init[i] = N - I0; init[n + i] = I0; }
    ok = cudaMemcpyAsync(comp, init.data(), 4 * n * sizeof(double), cudaMemcpyHostToDevice, st) == cudaSuccess &&
         cudaMemsetAsync(scalars, 0, 2 * sizeof(double), st) == cudaSuccess &&
         cudaStreamSynchronize(st) == cudaSuccess;
  }
  const SirPars q{N, beta, gamma, exp_noise};
  const unsigned block = 128, grid = static_cast<unsigned>((n + block - 1) / block);
  double time = 0;
  for (size_t d = 0; ok && d < n_data; ++d) {
    sir_run_kernel<<<grid, block, 0, st>>>(mb200_rng_state_dev(system), n, comp, q, time, data_time[d],
                                           data_incidence[d], logw);
    sir_weights_kernel<<<1, kScanThreads, 0, st>>>(logw, n, cum, scalars, mb200_rng_state_dev(filter), scalars + 1);
    if (resample) {
      sir_resample_kernel<<<grid, block, 0, st>>>(cum, n, scalars + 1, comp, comp2);
      double* t = comp; comp = comp2; comp2 = t;
    }
    time = data_time[d];
    ok = cudaGetLastError() == cudaSuccess;
  }
  if (ok) {
    ok = cudaMemcpyAsync(log_likelihood, scalars, sizeof(double), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (ok && final_state_host) {
      ok = cudaMemcpyAsync(final_state_host, comp, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    }
    ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
  }
  cudaFree(comp); cudaFree(comp2); cudaFree(logw); cudaFree(cum); cudaFree(scalars);
  return ok ? MB200_OK : MB200_ERR_CUDA;
}
