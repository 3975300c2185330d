// monty_b200/prng.hpp -- host-side `prng`: the reference's multi-stream
// container (inst/include/monty/random/prng.hpp:17-131) with the states
// living in HBM.  Same constructor forms and members; header-only, on top of
// the C-ABI (include/monty_b200.h, link with -lmonty_b200).
//
//   reference                               here
//   prng<T>(n, int seed, deterministic)     prng(n, seed, deterministic, device)
//   prng<T>(n, vector<int_type> seed, ...)  prng(n, std::vector<uint64_t>, ...)  (k full states, then jump from the last)
//   size(), state_size()                    same
//   jump(), long_jump()                     same (every stream; one kernel)
//   state(i)                                device_state() + stream index: the states are device
//                                           memory; kernels load them with monty::random::load_state
//   export_state() / import_state(v)        same, uint64 words in the reference's layout
//   deterministic()                         same
//
// The device array is uint64[size()][4] in exactly the exported layout, so a
// model kernel (see examples/sir_filter.cu) does
//     auto g = monty::random::load_state(states, i);  ...draw...;  store_state(states, i, g);
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../monty_b200.h"

namespace monty_b200 {

class prng {
public:
  using int_type = uint64_t;

  prng(size_t n, uint64_t seed, bool deterministic = false, int device = 0) {
    check(mb200_rng_create(seed, nullptr, 0, n, deterministic ? 1 : 0, 0, 0, device, &h_));
  }
  prng(size_t n, const std::vector<uint64_t>& seed, bool deterministic = false, int device = 0) {
    check(mb200_rng_create(0, reinterpret_cast<const uint8_t*>(seed.data()), seed.size() * 8, n,
                           deterministic ? 1 : 0, 0, 0, device, &h_));
  }
  ~prng() { if (h_) mb200_rng_free(h_); }
  prng(const prng&) = delete;
  prng& operator=(const prng&) = delete;
  prng(prng&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }

  size_t size() const { return mb200_rng_n_streams(h_); }
  size_t state_size() const { return size() * 4; }
  bool deterministic() const { return mb200_rng_deterministic(h_) != 0; }
  void jump() { check(mb200_rng_jump(h_, 1)); }
  void long_jump() { check(mb200_rng_long_jump(h_, 1)); }

  std::vector<uint64_t> export_state() const {
    std::vector<uint64_t> s(state_size());
    check(mb200_rng_state(h_, reinterpret_cast<uint8_t*>(s.data()), s.size() * 8));
    return s;
  }
  void import_state(const std::vector<uint64_t>& s) {
    check(mb200_rng_set_state(h_, reinterpret_cast<const uint8_t*>(s.data()), s.size() * 8));
  }

  // device side
  uint64_t* device_state() { return mb200_rng_state_dev(h_); }
  void* cuda_stream() { return mb200_rng_cuda_stream(h_); }
  void sync() { check(mb200_rng_sync(h_)); }
  mb200_rng* handle() { return h_; }

private:
  static void check(int rc) {
    if (rc != MB200_OK) throw std::runtime_error(mb200_last_error());
  }
  mb200_rng* h_ = nullptr;
};

}  // namespace monty_b200
