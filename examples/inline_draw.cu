// inline_draw.cu -- every sampler of the device-side monty::random API
// (include/monty_b200/random.cuh) called the way a model's update() calls it:
// one draw per stream, state loaded into registers, drawn from, stored back.
// Used by tests/test_gpu_facade.py to prove that the inline API produces the
// draws of the bulk C-ABI path (which the oracle tests pin to the reference).
#include <cuda_runtime.h>

#include "../include/monty_b200/random.cuh"
#include "draw_dist.cuh"
#include "examples.h"

namespace {

struct InlineArgs {
  uint64_t* states;
  size_t n;
  const double* par[4];
  unsigned stride[4];
  double* out;
  int dist;
};

template <typename T>
__global__ void inline_draw_kernel(const InlineArgs a) {
  namespace mr = monty::random;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= a.n) return;
  T p[4];
  for (int k = 0; k < 4; ++k) p[k] = a.par[k] ? static_cast<T>(a.par[k][a.stride[k] ? i : 0]) : T(0);
  auto g = mr::load_state(a.states, i);
  const T v = mb200_examples::draw_dist<T>(g, a.dist, p);
  mr::store_state(a.states, i, g);
  a.out[i] = static_cast<double>(v);
}

}  // namespace

extern "C" int mb200_example_inline_draw(mb200_rng* rng, int dist, int real_type,
                                         const double* const* params_host, const size_t* param_len,
                                         double* out_host) {
  if (!rng || !out_host) return MB200_ERR_ARG;
  if (cudaSetDevice(mb200_rng_device(rng)) != cudaSuccess) return MB200_ERR_CUDA;
  if (mb200_rng_sync(rng) != MB200_OK) return MB200_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(mb200_rng_cuda_stream(rng));
  InlineArgs a{};
  a.states = mb200_rng_state_dev(rng);
  a.n = mb200_rng_n_streams(rng);
  a.dist = dist;
  double* d_par[4] = {nullptr, nullptr, nullptr, nullptr};
  int rc = MB200_OK;
  for (int k = 0; k < 4 && rc == MB200_OK; ++k) {
    if (!params_host || !params_host[k] || param_len[k] == 0) continue;
    if (cudaMalloc(&d_par[k], param_len[k] * sizeof(double)) != cudaSuccess ||
        cudaMemcpyAsync(d_par[k], params_host[k], param_len[k] * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) {
      rc = MB200_ERR_CUDA;
    }
    a.par[k] = d_par[k];
    a.stride[k] = param_len[k] == 1 ? 0u : 1u;
  }
  if (rc == MB200_OK && cudaMalloc(&a.out, a.n * sizeof(double)) != cudaSuccess) rc = MB200_ERR_CUDA;
  if (rc == MB200_OK) {
    const unsigned block = 128, grid = static_cast<unsigned>((a.n + block - 1) / block);
    if (real_type == MB200_REAL_F32) inline_draw_kernel<float><<<grid, block, 0, st>>>(a);
    else inline_draw_kernel<double><<<grid, block, 0, st>>>(a);
    if (cudaGetLastError() != cudaSuccess ||
        cudaMemcpyAsync(out_host, a.out, a.n * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      rc = MB200_ERR_CUDA;
    }
  }
  for (double* p : d_par) if (p) cudaFree(p);
  if (a.out) cudaFree(a.out);
  return rc;
}

// ---- monty::random::multinomial inside a kernel ----------------------------------
namespace {
template <typename T>
__global__ void inline_multinomial_kernel(uint64_t* states, size_t n, const double* __restrict__ size, unsigned size_stride,
                                          const double* __restrict__ prob, int prob_len, unsigned prob_stride,
                                          double* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto g = monty::random::load_state(states, i);
  T pr[16], ret[16];
  for (int k = 0; k < prob_len; ++k) pr[k] = static_cast<T>(prob[static_cast<size_t>(prob_stride) * i * prob_len + k]);
  monty::random::multinomial<T>(g, static_cast<T>(size[size_stride ? i : 0]), pr, prob_len, ret);
  for (int k = 0; k < prob_len; ++k) out[i * prob_len + k] = static_cast<double>(ret[k]);
  monty::random::store_state(states, i, g);
}
}  // namespace

extern "C" int mb200_example_inline_multinomial(mb200_rng* rng, int real_type,
                                                const double* size_host, size_t size_len,
                                                const double* prob_host, int prob_len, size_t prob_cols,
                                                double* out_host) {
  if (!rng || !size_host || !prob_host || !out_host || prob_len < 1 || prob_len > 16) return MB200_ERR_ARG;
  if (cudaSetDevice(mb200_rng_device(rng)) != cudaSuccess) return MB200_ERR_CUDA;
  if (mb200_rng_sync(rng) != MB200_OK) return MB200_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(mb200_rng_cuda_stream(rng));
  const size_t n = mb200_rng_n_streams(rng);
  if ((size_len != 1 && size_len != n) || (prob_cols != 1 && prob_cols != n)) return MB200_ERR_ARG;
  double *d_size = nullptr, *d_prob = nullptr, *d_out = nullptr;
  int rc = MB200_OK;
  if (cudaMalloc(&d_size, size_len * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&d_prob, prob_cols * prob_len * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&d_out, n * prob_len * sizeof(double)) != cudaSuccess ||
      cudaMemcpyAsync(d_size, size_host, size_len * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(d_prob, prob_host, prob_cols * prob_len * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) {
    rc = MB200_ERR_CUDA;
  }
  if (rc == MB200_OK) {
    const unsigned block = 128, grid = static_cast<unsigned>((n + block - 1) / block);
    uint64_t* states = mb200_rng_state_dev(rng);
    if (real_type == MB200_REAL_F32) {
      inline_multinomial_kernel<float><<<grid, block, 0, st>>>(states, n, d_size, size_len == 1 ? 0u : 1u, d_prob, prob_len,
                                                               prob_cols == 1 ? 0u : 1u, d_out);
    } else {
      inline_multinomial_kernel<double><<<grid, block, 0, st>>>(states, n, d_size, size_len == 1 ? 0u : 1u, d_prob, prob_len,
                                                                prob_cols == 1 ? 0u : 1u, d_out);
    }
    if (cudaGetLastError() != cudaSuccess ||
        cudaMemcpyAsync(out_host, d_out, n * prob_len * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      rc = MB200_ERR_CUDA;
    }
  }
  if (d_size) cudaFree(d_size);
  if (d_prob) cudaFree(d_prob);
  if (d_out) cudaFree(d_out);
  return rc;
}

// ---- probe of the samplers' own math routines ----------------------------------
namespace {
__global__ void math_probe_kernel(int fn, const double* __restrict__ x, double* __restrict__ y, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double r;
  switch (fn) {
    case 0: r = mb200::m_log(v); break;
    case 1: r = mb200::m_log_normal(v); break;
    case 2: r = mb200::cos_0_2pi(v); break;
    case 3: r = mb200::sqrt_normal(v); break;
    case 4: r = sqrt(v); break;
    case 5: r = monty::random::contraction_is_off() ? 1.0 : 0.0; break;
    case 6: r = mb200::m_lgamma(v); break;
    case 7: r = mb200::m_exp(v); break;
    case 8: r = mb200::m_pow(v, x[i ^ 1]); break;        // exponent: the neighbouring element
    case 9: r = mb200::m_tan(v); break;
    case 10: r = mb200::m_log1p(v); break;
    case 11: r = mb200::m_erfc(v); break;
    case 12: r = static_cast<double>(mb200::m_tan(static_cast<float>(v))); break;
    default: r = 0.0; break;
  }
  y[i] = r;
}
}  // namespace

extern "C" int mb200_example_math_probe(int fn, const double* x_host, double* y_host, size_t n) {
  if (!x_host || !y_host || fn < 0 || fn > 12) return MB200_ERR_ARG;
  double *dx = nullptr, *dy = nullptr;
  int rc = MB200_OK;
  if (cudaMalloc(&dx, n * sizeof(double)) != cudaSuccess || cudaMalloc(&dy, n * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(dx, x_host, n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    rc = MB200_ERR_CUDA;
  }
  if (rc == MB200_OK) {
    math_probe_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(fn, dx, dy, n);
    if (cudaGetLastError() != cudaSuccess || cudaMemcpy(y_host, dy, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = MB200_ERR_CUDA;
    }
  }
  if (dx) cudaFree(dx);
  if (dy) cudaFree(dy);
  return rc;
}

// ---- monty::density::zi_* through the device-side API ---------------------------
namespace {
__global__ void zi_density_kernel(int which, const double* __restrict__ x, const double* __restrict__ pi0,
                                  const double* __restrict__ a, const double* __restrict__ b, int lg,
                                  double* __restrict__ y, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r;
  switch (which) {
    case 0: r = monty::density::zi_poisson<double>(x[i], pi0[i], a[i], lg != 0); break;
    case 1: r = monty::density::zi_negative_binomial_prob<double>(x[i], pi0[i], a[i], b[i], lg != 0); break;
    default: r = monty::density::zi_negative_binomial_mu<double>(x[i], pi0[i], a[i], b[i], lg != 0); break;
  }
  y[i] = r;
}
}  // namespace

extern "C" int mb200_example_zi_density(int which, const double* x_host, const double* pi0_host, const double* a_host,
                                        const double* b_host, int lg, double* y_host, size_t n) {
  if (!x_host || !pi0_host || !a_host || !y_host || which < 0 || which > 2 || (which > 0 && !b_host)) return MB200_ERR_ARG;
  double* d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const double* h[4] = {x_host, pi0_host, a_host, b_host ? b_host : a_host};
  int rc = MB200_OK;
  for (int k = 0; k < 5 && rc == MB200_OK; ++k) {
    if (cudaMalloc(&d[k], n * sizeof(double)) != cudaSuccess) rc = MB200_ERR_CUDA;
    else if (k < 4 && cudaMemcpy(d[k], h[k], n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) rc = MB200_ERR_CUDA;
  }
  if (rc == MB200_OK) {
    zi_density_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(which, d[0], d[1], d[2], d[3], lg, d[4], n);
    if (cudaGetLastError() != cudaSuccess || cudaMemcpy(y_host, d[4], n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = MB200_ERR_CUDA;
    }
  }
  for (int k = 0; k < 5; ++k) if (d[k]) cudaFree(d[k]);
  return rc;
}
