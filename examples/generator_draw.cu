// generator_draw.cu -- the device-side monty::random API on the reference's OTHER
// generators: a model compiled for real_type = float draws from
// monty::random::generator<float> = xoshiro128+ (hdr/random.hpp:31-47), and any of the
// twelve state types can be named explicitly.  Streams are seeded by
// monty_b200::device_prng<G> exactly like the reference's prng<G>(n, seed); each stream
// then makes n_samples draws of one distribution inside a kernel.
// tests/test_gpu_generators.py compares the draws and the final states with the
// reference's own templates instantiated on the same generator (oracle/ref_harness.cpp).
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "../include/monty_b200/device_prng.cuh"
#include "draw_dist.cuh"
#include "examples.h"

namespace {

struct GenArgs {
  size_t n, n_samples;
  const double* par[4];
  unsigned stride[4];
  double* out;
  int dist;
};

template <typename T, class G>
__global__ void generator_draw_kernel(typename G::int_type* states, const GenArgs a) {
  namespace mr = monty::random;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= a.n) return;
  T p[4];
  for (int k = 0; k < 4; ++k) p[k] = a.par[k] ? static_cast<T>(a.par[k][a.stride[k] ? i : 0]) : T(0);
  G g = mr::load_state_of<G>(states, i);
  for (size_t j = 0; j < a.n_samples; ++j) {
    a.out[i * a.n_samples + j] = static_cast<double>(mb200_examples::draw_dist<T>(g, a.dist, p));
  }
  mr::store_state_of<G>(states, i, g);
}

template <typename T, class G>
int run(uint64_t seed, int jumps, int long_jumps, GenArgs a, double* out_host, void* state_out) {
  try {
    monty_b200::device_prng<G> rng(a.n, seed);
    for (int k = 0; k < jumps; ++k) rng.jump();
    for (int k = 0; k < long_jumps; ++k) rng.long_jump();
    if (a.n_samples > 0) {
      generator_draw_kernel<T, G><<<static_cast<unsigned>((a.n + 127) / 128), 128>>>(rng.device_state(), a);
      if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return MB200_ERR_CUDA;
      if (cudaMemcpy(out_host, a.out, a.n * a.n_samples * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return MB200_ERR_CUDA;
    }
    if (state_out) {
      const auto st = rng.export_state();
      std::memcpy(state_out, st.data(), st.size() * sizeof(typename G::int_type));
    }
    return MB200_OK;
  } catch (const std::exception&) {
    return MB200_ERR_CUDA;
  }
}

}  // namespace

extern "C" int mb200_example_generator_draw(int gen, int dist, int real_type, uint64_t seed, size_t n_streams,
                                            size_t n_samples, int jumps, int long_jumps,
                                            const double* const* params_host, const size_t* param_len,
                                            double* out_host, void* state_out) {
  namespace mr = monty::random;
  if (n_streams == 0 || (!out_host && n_samples > 0)) return MB200_ERR_ARG;
  GenArgs a{};
  a.n = n_streams;
  a.n_samples = n_samples;
  a.dist = dist;
  double* d_par[4] = {nullptr, nullptr, nullptr, nullptr};
  int rc = MB200_OK;
  for (int k = 0; k < 4 && rc == MB200_OK; ++k) {
    if (!params_host || !params_host[k] || param_len[k] == 0) continue;
    if (cudaMalloc(&d_par[k], param_len[k] * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(d_par[k], params_host[k], param_len[k] * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
      rc = MB200_ERR_CUDA;
    }
    a.par[k] = d_par[k];
    a.stride[k] = param_len[k] == 1 ? 0u : 1u;
  }
  if (rc == MB200_OK && n_samples > 0 && cudaMalloc(&a.out, n_streams * n_samples * sizeof(double)) != cudaSuccess) rc = MB200_ERR_CUDA;
  if (rc == MB200_OK) {
    const bool f32 = real_type == MB200_REAL_F32;
    // the default generators of both real types, and one of each other family
    switch (gen) {
    case MB200_GEN_XOSHIRO128PLUS:            // generator<float>
      rc = f32 ? run<float, mr::generator<float>>(seed, jumps, long_jumps, a, out_host, state_out)
               : run<double, mr::xoshiro128plus>(seed, jumps, long_jumps, a, out_host, state_out);
      break;
    case MB200_GEN_XOSHIRO256PLUS:            // generator<double>, through the generic container
      rc = f32 ? run<float, mr::xoshiro256plus>(seed, jumps, long_jumps, a, out_host, state_out)
               : run<double, mr::generator<double>>(seed, jumps, long_jumps, a, out_host, state_out);
      break;
    case MB200_GEN_XOSHIRO256STARSTAR:
      rc = f32 ? MB200_ERR_ARG : run<double, mr::xoshiro256starstar>(seed, jumps, long_jumps, a, out_host, state_out);
      break;
    case MB200_GEN_XOROSHIRO128PLUSPLUS:
      rc = f32 ? MB200_ERR_ARG : run<double, mr::xoroshiro128plusplus>(seed, jumps, long_jumps, a, out_host, state_out);
      break;
    case MB200_GEN_XOSHIRO512STARSTAR:
      rc = f32 ? MB200_ERR_ARG : run<double, mr::xoshiro512starstar>(seed, jumps, long_jumps, a, out_host, state_out);
      break;
    case MB200_GEN_XOSHIRO128STARSTAR:
      rc = f32 ? run<float, mr::xoshiro128starstar>(seed, jumps, long_jumps, a, out_host, state_out) : MB200_ERR_ARG;
      break;
    default:
      rc = MB200_ERR_ARG;
    }
  }
  for (double* p : d_par) if (p) cudaFree(p);
  if (a.out) cudaFree(a.out);
  return rc;
}

// device_prng<G>(n, seed_words): the reference's seed-vector constructor (hdr/prng.hpp:38-53: as
// many full states as the words give, the rest by jump() from the last one), then optionally
// import_state() of the exported states into a second container and one jump() there --
// export / import round trip.  words: n_words values of the generator's int_type (uint32 for
// xoshiro128, uint64 otherwise); state_out: int_type[n_streams][state words].
namespace {
template <class G>
int seed_vector_run(const void* words, size_t n_words, size_t n_streams, int reimport_and_jump, void* state_out) {
  using W = typename G::int_type;
  try {
    std::vector<W> seed(static_cast<const W*>(words), static_cast<const W*>(words) + n_words);
    monty_b200::device_prng<G> a(n_streams, seed);
    std::vector<W> st = a.export_state();
    if (reimport_and_jump) {
      monty_b200::device_prng<G> b(n_streams, static_cast<uint64_t>(1));
      b.import_state(st);
      b.jump();
      st = b.export_state();
    }
    std::memcpy(state_out, st.data(), st.size() * sizeof(W));
    return MB200_OK;
  } catch (const std::exception&) {
    return MB200_ERR_ARG;
  }
}
}  // namespace

extern "C" int mb200_example_generator_seed_vector(int gen, const void* words, size_t n_words, size_t n_streams,
                                                   int reimport_and_jump, void* state_out) {
  namespace mr = monty::random;
  if (!words || !state_out || n_streams == 0) return MB200_ERR_ARG;
  switch (gen) {
  case MB200_GEN_XOSHIRO128PLUS: return seed_vector_run<mr::xoshiro128plus>(words, n_words, n_streams, reimport_and_jump, state_out);
  case MB200_GEN_XOSHIRO256PLUS: return seed_vector_run<mr::xoshiro256plus>(words, n_words, n_streams, reimport_and_jump, state_out);
  case MB200_GEN_XOROSHIRO128PLUSPLUS: return seed_vector_run<mr::xoroshiro128plusplus>(words, n_words, n_streams, reimport_and_jump, state_out);
  case MB200_GEN_XOSHIRO512STARSTAR: return seed_vector_run<mr::xoshiro512starstar>(words, n_words, n_streams, reimport_and_jump, state_out);
  default: return MB200_ERR_ARG;
  }
}
