// monty_b200/device_prng.cuh -- the reference's multi-stream container
// (inst/include/monty/random/prng.hpp:17-131) for ANY of the twelve generators, with the
// states in HBM as int_type[size()][G::size()] -- the exported layout, read and written
// by kernels through monty::random::load_state_of<G> / store_state_of<G>.
//
// prng.hpp (plain C++, on top of the C-ABI) is the xoshiro256+ container behind the R
// layer; this one is what a model compiled for real_type = float uses, whose
// monty::random::generator<float> is xoshiro128+ (hdr/random.hpp:31-47).  Header-only CUDA
// C++ (compile with nvcc).  Same constructor forms and members as the reference:
//
//   prng<G>(n, int seed)                  device_prng<G>(n, seed)
//   prng<G>(n, vector<int_type> seed)     device_prng<G>(n, seed_words): as many full states
//                                         as the words give, then jump() from the last one
//                                         (hdr/prng.hpp:38-53)
//   size(), state_size(), jump(), long_jump(), export_state(), import_state(), state ptr
//
// Seeding walks the jump chain on the host (stream i = i jumps from the seed, as the
// reference does serially); jump() / long_jump() of all streams run as one kernel.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "random.cuh"

namespace monty_b200 {

namespace detail {
template <class G, bool LONG>
__global__ void device_prng_jump_kernel(typename G::int_type* states, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G g = monty::random::load_state_of<G>(states, i);
  if (LONG) monty::random::long_jump(g); else monty::random::jump(g);
  monty::random::store_state_of<G>(states, i, g);
}
}  // namespace detail

template <class G>
class device_prng {
public:
  using rng_state = G;
  using int_type = typename G::int_type;

  device_prng(size_t n, uint64_t seed, int device = 0) : n_(n), device_(device) {
    G g = monty::random::seed<G>(seed);
    std::vector<int_type> words;
    for (size_t k = 0; k < G::size(); ++k) words.push_back(g[k]);
    init(words);
  }
  device_prng(size_t n, const std::vector<int_type>& seed, int device = 0) : n_(n), device_(device) { init(seed); }
  ~device_prng() { if (d_) cudaFree(d_); }
  device_prng(const device_prng&) = delete;
  device_prng& operator=(const device_prng&) = delete;

  size_t size() const { return n_; }
  size_t state_size() const { return n_ * G::size(); }
  void jump() { launch<false>(); }
  void long_jump() { launch<true>(); }

  std::vector<int_type> export_state() const {
    std::vector<int_type> s(state_size());
    check(cudaMemcpy(s.data(), d_, s.size() * sizeof(int_type), cudaMemcpyDeviceToHost));
    return s;
  }
  void import_state(const std::vector<int_type>& s) {
    if (s.size() != state_size()) throw std::runtime_error("import_state: wrong length");
    check(cudaMemcpy(d_, s.data(), s.size() * sizeof(int_type), cudaMemcpyHostToDevice));
  }
  int_type* device_state() { return d_; }

private:
  static void check(cudaError_t e) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("device_prng: ") + cudaGetErrorString(e));
  }
  // hdr/prng.hpp:38-53
  void init(const std::vector<int_type>& seed) {
    constexpr size_t len = G::size();
    if (seed.size() == 0 || seed.size() % len != 0) throw std::runtime_error("seed length must be a multiple of the state length");
    std::vector<int_type> host(n_ * len);
    const size_t given = seed.size() / len < n_ ? seed.size() / len : n_;
    for (size_t i = 0; i < given * len; ++i) host[i] = seed[i];
    G g{};
    for (size_t i = given; i < n_; ++i) {
      if (i == given) for (size_t k = 0; k < len; ++k) g[k] = host[(i - 1) * len + k];
      monty::random::jump(g);
      for (size_t k = 0; k < len; ++k) host[i * len + k] = g[k];
    }
    check(cudaSetDevice(device_));
    check(cudaMalloc(&d_, host.size() * sizeof(int_type)));
    check(cudaMemcpy(d_, host.data(), host.size() * sizeof(int_type), cudaMemcpyHostToDevice));
  }
  template <bool LONG> void launch() {
    if (n_ == 0) return;
    check(cudaSetDevice(device_));
    detail::device_prng_jump_kernel<G, LONG><<<static_cast<unsigned>((n_ + 127) / 128), 128>>>(d_, n_);
    check(cudaGetLastError());
    check(cudaDeviceSynchronize());
  }
  size_t n_;
  int device_;
  int_type* d_ = nullptr;
};

}  // namespace monty_b200
