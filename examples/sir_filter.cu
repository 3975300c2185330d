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
// per draw.  The weight normalisation, prefix sum and gather run as five small
// grid-wide kernels per observation.
#include <cuda_runtime.h>

#include <vector>

#include "../include/monty_b200/random.cuh"
#include "examples.h"

namespace {
namespace mr = monty::random;

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

// ---- weights: four small grid-wide kernels per observation -----------------
constexpr int kWThreads = 256;
constexpr int kWPer = 8;                          // consecutive particles per thread
constexpr int kWChunk = kWThreads * kWPer;        // particles per block

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

// 1. per-block maximum of the log weights
__global__ void __launch_bounds__(kWThreads)
sir_max_kernel(const double* logw, size_t n, double* part_max) {
  __shared__ double scratch[32];
  const size_t base = static_cast<size_t>(blockIdx.x) * kWChunk;
  double m = -INFINITY;
  for (int k = 0; k < kWPer; ++k) {
    const size_t i = base + k * kWThreads + threadIdx.x;       // coalesced
    if (i < n) m = fmax(m, logw[i]);
  }
  m = block_reduce(m, true, scratch);
  if (threadIdx.x == 0) part_max[blockIdx.x] = m;
}

// 2. w = exp(logw - max) into `cum` (unnormalised), per-block sums
__global__ void __launch_bounds__(kWThreads)
sir_weight_kernel(const double* logw, size_t n, const double* part_max, int n_blocks, double* cum, double* part_sum,
                  double* max_out) {
  __shared__ double scratch[32];
  double m = -INFINITY;
  for (int b = threadIdx.x; b < n_blocks; b += kWThreads) m = fmax(m, part_max[b]);
  m = block_reduce(m, true, scratch);
  const size_t base = static_cast<size_t>(blockIdx.x) * kWChunk;
  double s = 0;
  for (int k = 0; k < kWPer; ++k) {
    const size_t i = base + k * kWThreads + threadIdx.x;
    if (i < n) { const double w = exp(logw[i] - m); cum[i] = w; s += w; }
  }
  s = block_reduce(s, false, scratch);
  if (threadIdx.x == 0) { part_sum[blockIdx.x] = s; if (blockIdx.x == 0) *max_out = m; }
}

// 3. one block: exclusive scan of the block sums, ll += log(mean(w)) + max, the filter's uniform
__global__ void __launch_bounds__(1024)
sir_total_kernel(double* part_sum, int n_blocks, size_t n, const double* max_in, double* scalars, uint64_t* filter_state) {
  __shared__ double sh[1024];
  const int per = (n_blocks + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, n_blocks);
  double s = 0;
  for (int b = lo; b < hi; ++b) s += part_sum[b];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0;
    for (int t = 0; t < 1024; ++t) { const double v = sh[t]; sh[t] = run; run += v; }
    scalars[2] = run;                                               // sum of the weights
    scalars[0] += log(run / static_cast<double>(n)) + *max_in;      // log-likelihood so far
    auto g = mr::load_state(filter_state, 0);
    scalars[1] = mr::random_real<double>(g);
    mr::store_state(filter_state, 0, g);
  }
  __syncthreads();
  double run = sh[threadIdx.x];
  for (int b = lo; b < hi; ++b) { const double v = part_sum[b]; part_sum[b] = run; run += v; }
}

// 4. cum = cumsum(w) / sum(w): scan inside the block's chunk + the block's offset
__global__ void __launch_bounds__(kWThreads)
sir_cumsum_kernel(double* cum, size_t n, const double* part_off, const double* scalars) {
  __shared__ double sh[kWThreads];
  const size_t base = static_cast<size_t>(blockIdx.x) * kWChunk + static_cast<size_t>(threadIdx.x) * kWPer;
  double w[kWPer];
  double s = 0;
  for (int k = 0; k < kWPer; ++k) { w[k] = base + k < n ? cum[base + k] : 0.0; s += w[k]; }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < kWThreads; d <<= 1) {                       // Hillis-Steele inclusive scan
    const double v = threadIdx.x >= d ? sh[threadIdx.x - d] : 0.0;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  double run = part_off[blockIdx.x] + sh[threadIdx.x] - s;
  const double total = scalars[2];
  for (int k = 0; k < kWPer; ++k) {
    run += w[k];
    if (base + k < n) cum[base + k] = run / total;
  }
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
  double *comp = nullptr, *comp2 = nullptr, *logw = nullptr, *cum = nullptr, *scalars = nullptr, *part = nullptr;
  const int n_wblocks = static_cast<int>((n + kWChunk - 1) / kWChunk);
  bool ok = cudaMalloc(&comp, 4 * n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&comp2, 4 * n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&logw, n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&cum, n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&part, 2 * static_cast<size_t>(n_wblocks) * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&scalars, 4 * sizeof(double)) == cudaSuccess;
  if (ok) {
    std::vector<double> init(4 * n, 0.0);
    for (size_t i = 0; i < n; ++i) { init[i] = N - I0; init[n + i] = I0; }
    ok = cudaMemcpyAsync(comp, init.data(), 4 * n * sizeof(double), cudaMemcpyHostToDevice, st) == cudaSuccess &&
         cudaMemsetAsync(scalars, 0, 4 * sizeof(double), st) == cudaSuccess &&
         cudaStreamSynchronize(st) == cudaSuccess;
  }
  const SirPars q{N, beta, gamma, exp_noise};
  const unsigned block = 128, grid = static_cast<unsigned>((n + block - 1) / block);
  double time = 0;
  for (size_t d = 0; ok && d < n_data; ++d) {
    sir_run_kernel<<<grid, block, 0, st>>>(mb200_rng_state_dev(system), n, comp, q, time, data_time[d],
                                           data_incidence[d], logw);
    double* part_max = part;
    double* part_sum = part + n_wblocks;
    sir_max_kernel<<<n_wblocks, kWThreads, 0, st>>>(logw, n, part_max);
    sir_weight_kernel<<<n_wblocks, kWThreads, 0, st>>>(logw, n, part_max, n_wblocks, cum, part_sum, scalars + 3);
    sir_total_kernel<<<1, 1024, 0, st>>>(part_sum, n_wblocks, n, scalars + 3, scalars, mb200_rng_state_dev(filter));
    if (resample) sir_cumsum_kernel<<<n_wblocks, kWThreads, 0, st>>>(cum, n, part_sum, scalars);
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
  cudaFree(comp); cudaFree(comp2); cudaFree(logw); cudaFree(cum); cudaFree(scalars); cudaFree(part);
  return ok ? MB200_OK : MB200_ERR_CUDA;
}
