// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A C-ABI driver around the UNMODIFIED reference headers, compiled from where
// they lie under /root/reference/inst/include (see oracle/Makefile, target
// _ref/libmonty_ref.so).  No reference source is copied into this repository:
// this file only #includes the reference's public headers and restates the
// 20-line driver loops of the cpp11 layer (which cannot be built here: no R,
// no cpp11):
//
//   loops over streams x samples     ref: src/random.cpp:89-195, 211-338
//   parameter recycling (len 1 | n)  ref: src/random.cpp:21-42
//   state export / import            ref: src/random.cpp:356-379
//   jump / long_jump n times         ref: src/random.cpp:382-396
//   golden protocol                  ref: src/test_rng.cpp:21-38
//   density loops                    ref: src/density.cpp:6-254
//   multinomial loops                ref: src/random.cpp:884-934
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the resulting library.  The product
// (monty_b200/) never does.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

// The single R dependency on the arithmetic path (ref: beta_binomial.hpp:16).
#define R_FINITE(x) std::isfinite(x)

#include <monty/random/random.hpp>
#include <monty/random/density.hpp>

#include "../include/monty_b200.h"

namespace mr = monty::random;
namespace md = monty::density;

static thread_local std::string g_err;

using rng256p = mr::xoshiro256plus;

namespace {

struct input {  // ref: src/random.cpp:21-42
  const double* data;
  bool fixed;
  double operator[](size_t i) const { return data[fixed ? 0 : i]; }
};

template <typename real_type, typename G = rng256p>
real_type draw_one(int dist, G& s, const input* p, size_t i) {
  using T = real_type;
  const T a = p[0].data ? static_cast<T>(p[0][i]) : 0;
  const T b = p[1].data ? static_cast<T>(p[1][i]) : 0;
  const T c = p[2].data ? static_cast<T>(p[2][i]) : 0;
  const T d = p[3].data ? static_cast<T>(p[3][i]) : 0;
  switch (dist) {
  case MB200_DIST_REAL: return mr::random_real<T>(s);
  case MB200_DIST_UNIFORM: return mr::uniform<T>(s, a, b);
  case MB200_DIST_EXPONENTIAL_RATE: return mr::exponential_rate<T>(s, a);
  case MB200_DIST_EXPONENTIAL_MEAN: return mr::exponential_mean<T>(s, a);
  case MB200_DIST_NORMAL_BOX_MULLER:
    return mr::normal<T, mr::algorithm::normal::box_muller>(s, a, b);
  case MB200_DIST_NORMAL_POLAR:
    return mr::normal<T, mr::algorithm::normal::polar>(s, a, b);
  case MB200_DIST_NORMAL_ZIGGURAT:
    return mr::normal<T, mr::algorithm::normal::ziggurat>(s, a, b);
  case MB200_DIST_BINOMIAL: return mr::binomial<T>(s, a, b);
  case MB200_DIST_POISSON: return mr::poisson<T>(s, a);
  case MB200_DIST_GAMMA_SCALE: return mr::gamma_scale<T>(s, a, b);
  case MB200_DIST_GAMMA_RATE: return mr::gamma_rate<T>(s, a, b);
  case MB200_DIST_BETA: return mr::beta<T>(s, a, b);
  case MB200_DIST_NEGATIVE_BINOMIAL_PROB:
    return mr::negative_binomial_prob<T>(s, a, b);
  case MB200_DIST_NEGATIVE_BINOMIAL_MU:
    return mr::negative_binomial_mu<T>(s, a, b);
  case MB200_DIST_HYPERGEOMETRIC: return mr::hypergeometric<T>(s, a, b, c);
  case MB200_DIST_TRUNCATED_NORMAL:
    return mr::truncated_normal<T>(s, a, b, c, d);
  case MB200_DIST_CAUCHY: return mr::cauchy<T>(s, a, b);
  case MB200_DIST_WEIBULL: return mr::weibull<T>(s, a, b);
  case MB200_DIST_LOG_NORMAL: return mr::log_normal<T>(s, a, b);
  case MB200_DIST_BETA_BINOMIAL_PROB:
    return mr::beta_binomial_prob<T>(s, a, b, c);
  case MB200_DIST_BETA_BINOMIAL_AB: return mr::beta_binomial_ab<T>(s, a, b, c);
  case MB200_DIST_ZI_POISSON: return mr::zi_poisson<T>(s, a, b);
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_PROB:
    return mr::zi_negative_binomial_prob<T>(s, a, b, c);
  case MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU:
    return mr::zi_negative_binomial_mu<T>(s, a, b, c);
  default: throw std::runtime_error("unknown distribution id");
  }
}

const int k_n_params[MB200_DIST_COUNT] = {
  0, 2, 1, 1, 2, 2, 2, 2, 1, 2, 2, 2, 2, 2, 3, 4, 2, 2, 2, 3, 3, 2, 3, 3};

void load_states(const uint64_t* state, size_t n, bool deterministic,
                 std::vector<rng256p>& out) {
  out.resize(n);
  for (size_t i = 0; i < n; ++i) {
    std::memcpy(out[i].state, state + 4 * i, 32);
    out[i].deterministic = deterministic;
  }
}

void store_states(const std::vector<rng256p>& in, uint64_t* state) {
  for (size_t i = 0; i < in.size(); ++i) {
    std::memcpy(state + 4 * i, in[i].state, 32);
  }
}

template <typename real_type>
int sample_impl(int dist, std::vector<rng256p>& st, size_t n_samples,
                const input* p, double* out, int n_threads) {
  const size_t n_streams = st.size();
  if (n_threads <= 1) {
    // exactly the serial stream-outer / sample-inner loop of
    // src/random.cpp:211-228; an exception leaves earlier streams advanced
    for (size_t i = 0; i < n_streams; ++i) {
      for (size_t j = 0; j < n_samples; ++j) {
        out[n_samples * i + j] = draw_one<real_type>(dist, st[i], p, i);
      }
    }
    return 0;
  }
  // the standalone-OpenMP usage pattern, ref: inst/random/openmp/rnguse.cpp:19-24
  int failed = 0;
  std::string msg;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
  for (long long i = 0; i < (long long)n_streams; ++i) {
    try {
      for (size_t j = 0; j < n_samples; ++j) {
        out[n_samples * i + j] = draw_one<real_type>(dist, st[i], p, i);
      }
    } catch (const std::exception& e) {
#ifdef _OPENMP
#pragma omp critical
#endif
      { failed = 1; msg = e.what(); }
    }
  }
  if (failed) throw std::runtime_error(msg);
  return 0;
}

template <typename G>
void run_protocol(uint64_t seed, int n, uint64_t* values, uint64_t* state_out,
                  int* n_words) {
  G s = mr::seed<G>(seed);
  for (int i = 0; i < 3 * n; ++i) {  // ref: src/test_rng.cpp:27-35
    if (i == n - 1) {
      mr::jump(s);
    } else if (i == 2 * n - 1) {
      mr::long_jump(s);
    }
    values[i] = static_cast<uint64_t>(mr::next(s));
  }
  *n_words = static_cast<int>(G::size());
  for (size_t k = 0; k < G::size(); ++k) state_out[k] = s.state[k];
}

template <typename G>
void raw_impl(uint64_t* state, size_t n_streams, size_t n_samples, uint64_t* out) {
  for (size_t i = 0; i < n_streams; ++i) {
    G s;
    std::memcpy(s.state, state + 4 * i, 32);
    for (size_t j = 0; j < n_samples; ++j) out[n_samples * i + j] = mr::next(s);
    std::memcpy(state + 4 * i, s.state, 32);
  }
}

template <typename T>
struct arg {  // one density argument: int32 or double vector, indexed by i
  const void* p; int is_i32;
  T operator[](size_t i) const {
    return is_i32 ? static_cast<T>(static_cast<const int32_t*>(p)[i])
                  : static_cast<T>(static_cast<const double*>(p)[i]);
  }
};

template <typename T>
double density_one(int dens, size_t i, const arg<T>& x, const arg<T>* q, bool lg) {
  switch (dens) {
  case MB200_DENS_BINOMIAL: return md::binomial<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_NORMAL: return md::normal<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_NEGATIVE_BINOMIAL_MU:
    return md::negative_binomial_mu<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_NEGATIVE_BINOMIAL_PROB:
    return md::negative_binomial_prob<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_BETA_BINOMIAL_PROB:
    return md::beta_binomial_prob<T>(x[i], q[0][i], q[1][i], q[2][i], lg);
  case MB200_DENS_BETA_BINOMIAL_AB:
    return md::beta_binomial_ab<T>(x[i], q[0][i], q[1][i], q[2][i], lg);
  case MB200_DENS_POISSON: return md::poisson<T>(x[i], q[0][i], lg);
  case MB200_DENS_UNIFORM: return md::uniform<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_EXPONENTIAL_RATE: return md::exponential_rate<T>(x[i], q[0][i], lg);
  case MB200_DENS_EXPONENTIAL_MEAN: return md::exponential_mean<T>(x[i], q[0][i], lg);
  case MB200_DENS_GAMMA_RATE: return md::gamma_rate<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_GAMMA_SCALE: return md::gamma_scale<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_BETA: return md::beta<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_HYPERGEOMETRIC:
    return md::hypergeometric<T>(x[i], q[0][i], q[1][i], q[2][i], lg);
  case MB200_DENS_LOG_NORMAL: return md::log_normal<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_TRUNCATED_NORMAL:
    return md::truncated_normal<T>(x[i], q[0][i], q[1][i], q[2][i], q[3][i], lg);
  case MB200_DENS_CAUCHY: return md::cauchy<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_WEIBULL: return md::weibull<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_ZI_POISSON: return md::zi_poisson<T>(x[i], q[0][i], q[1][i], lg);
  case MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU:
    return md::zi_negative_binomial_mu<T>(x[i], q[0][i], q[1][i], q[2][i], lg);
  case MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB:
    return md::zi_negative_binomial_prob<T>(x[i], q[0][i], q[1][i], q[2][i], lg);
  default: throw std::runtime_error("unknown density id");
  }
}

template <typename T>
void density_impl(int dens, size_t n, const void* x, int x_is_i32,
                  const void* const* params, const int* param_is_i32, int lg,
                  double* out, int n_threads) {
  arg<T> ax{x, x_is_i32};
  arg<T> q[MB200_MAX_PARAMS];
  for (int k = 0; k < MB200_MAX_PARAMS; ++k) {
    q[k] = arg<T>{params ? params[k] : nullptr, param_is_i32 ? param_is_i32[k] : 0};
  }
  if (n_threads <= 1) {
    for (size_t i = 0; i < n; ++i) out[i] = density_one<T>(dens, i, ax, q, lg != 0);
  } else {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (long long i = 0; i < (long long)n; ++i)
      out[i] = density_one<T>(dens, i, ax, q, lg != 0);
  }
}

}  // namespace

// prng<G>(n, seed) -> jumps -> n_samples draws per stream, for any generator
// ref: hdr/prng.hpp:29-73, src/random.cpp:211-228
template <typename real_type, typename G>
int generator_draw_impl(int dist, uint64_t seed, size_t n_streams, size_t n_samples,
                        int jumps, int long_jumps, const input* p, double* out, void* state_out) {
  mr::prng<G> r(n_streams, mr::seed_data<G>(seed), false);
  for (int k = 0; k < jumps; ++k) r.jump();
  for (int k = 0; k < long_jumps; ++k) r.long_jump();
  for (size_t i = 0; i < n_streams; ++i) {
    for (size_t j = 0; j < n_samples; ++j) {
      out[n_samples * i + j] = draw_one<real_type, G>(dist, r.state(i), p, i);
    }
  }
  if (state_out) {
    const auto v = r.export_state();
    std::memcpy(state_out, v.data(), v.size() * sizeof(typename G::int_type));
  }
  return 0;
}

// prng<G>(n, seed vector) [-> export -> import into another prng -> jump] -> export
// ref: hdr/prng.hpp:38-53, 92-122
template <typename G>
int seed_vector_impl(const void* words, size_t n_words, size_t n_streams, int reimport_and_jump, void* state_out) {
  using W = typename G::int_type;
  std::vector<W> seed(static_cast<const W*>(words), static_cast<const W*>(words) + n_words);
  mr::prng<G> a(n_streams, seed, false);
  std::vector<W> st = a.export_state();
  if (reimport_and_jump) {
    mr::prng<G> b(n_streams, 1, false);
    b.import_state(st);
    b.jump();
    st = b.export_state();
  }
  std::memcpy(state_out, st.data(), st.size() * sizeof(W));
  return 0;
}

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

int ref_generator_seed_vector(int gen, const void* words, size_t n_words, size_t n_streams,
                              int reimport_and_jump, void* state_out) {
  try {
    switch (gen) {
    case MB200_GEN_XOSHIRO128PLUS: return seed_vector_impl<mr::xoshiro128plus>(words, n_words, n_streams, reimport_and_jump, state_out);
    case MB200_GEN_XOSHIRO256PLUS: return seed_vector_impl<mr::xoshiro256plus>(words, n_words, n_streams, reimport_and_jump, state_out);
    case MB200_GEN_XOROSHIRO128PLUSPLUS: return seed_vector_impl<mr::xoroshiro128plusplus>(words, n_words, n_streams, reimport_and_jump, state_out);
    case MB200_GEN_XOSHIRO512STARSTAR: return seed_vector_impl<mr::xoshiro512starstar>(words, n_words, n_streams, reimport_and_jump, state_out);
    default: break;
    }
  } catch (const std::exception& e) { g_err = e.what(); return -3; }
  g_err = "unsupported generator";
  return -2;
}

// the reference's samplers instantiated on its other generators (default for float:
// xoshiro128+, hdr/random.hpp:31-47); mirrors examples/generator_draw.cu
int ref_generator_draw(int gen, int dist, int real_type, uint64_t seed, size_t n_streams,
                       size_t n_samples, int jumps, int long_jumps,
                       const double* const* params, const size_t* param_len,
                       double* out, void* state_out) {
  input p[MB200_MAX_PARAMS] = {};
  const int np = (dist >= 0 && dist < MB200_DIST_COUNT) ? k_n_params[dist] : -1;
  if (np < 0) { g_err = "unknown distribution id"; return -2; }
  for (int k = 0; k < np; ++k) p[k] = input{params[k], param_len[k] == 1};
  const bool f32 = real_type == MB200_REAL_F32;
  try {
#define GD(T, G) return generator_draw_impl<T, G>(dist, seed, n_streams, n_samples, jumps, long_jumps, p, out, state_out)
    switch (gen) {
    case MB200_GEN_XOSHIRO128PLUS:
      if (f32) GD(float, mr::generator<float>); else GD(double, mr::xoshiro128plus);
    case MB200_GEN_XOSHIRO256PLUS:
      if (f32) GD(float, mr::xoshiro256plus); else GD(double, mr::generator<double>);
    case MB200_GEN_XOSHIRO256STARSTAR: if (!f32) GD(double, mr::xoshiro256starstar); break;
    case MB200_GEN_XOROSHIRO128PLUSPLUS: if (!f32) GD(double, mr::xoroshiro128plusplus); break;
    case MB200_GEN_XOSHIRO512STARSTAR: if (!f32) GD(double, mr::xoshiro512starstar); break;
    case MB200_GEN_XOSHIRO128STARSTAR: if (f32) GD(float, mr::xoshiro128starstar); break;
    default: break;
    }
#undef GD
  } catch (const std::exception& e) { g_err = e.what(); return -3; }
  g_err = "unsupported generator / real type";
  return -2;
}

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// prng<xoshiro256plus>(n_streams, seed_data(seed))   ref: prng.hpp:29-53
int ref_seed_states(uint64_t seed, size_t n_streams, uint64_t* out) {
  try {
    mr::prng<rng256p> r(n_streams, mr::seed_data<rng256p>(seed), false);
    auto v = r.export_state();
    std::memcpy(out, v.data(), v.size() * sizeof(uint64_t));
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// prng<xoshiro256plus>(n_streams, raw seed words)    ref: prng.hpp:38-53
int ref_seed_states_raw(const uint64_t* words, size_t n_words, size_t n_streams,
                        uint64_t* out) {
  try {
    std::vector<uint64_t> seed(words, words + n_words);
    mr::prng<rng256p> r(n_streams, seed, false);
    auto v = r.export_state();
    std::memcpy(out, v.data(), v.size() * sizeof(uint64_t));
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// every stream jumps n times                          ref: src/random.cpp:382-396
int ref_jump(uint64_t* state, size_t n_streams, int n, int is_long) {
  std::vector<rng256p> st;
  load_states(state, n_streams, false, st);
  for (int k = 0; k < n; ++k) {
    for (auto& s : st) {
      if (is_long) mr::long_jump(s); else mr::jump(s);
    }
  }
  store_states(st, state);
  return 0;
}

int ref_n_params(int dist) {
  return (dist >= 0 && dist < MB200_DIST_COUNT) ? k_n_params[dist] : -1;
}

// monty_random_n_<dist>; state is updated in place    ref: src/random.cpp:211-338
int ref_sample(int dist, int real_type, uint64_t* state, size_t n_streams,
               int deterministic, size_t n_samples,
               const double* const* params, const size_t* param_len,
               double* out, int n_threads) {
  std::vector<rng256p> st;
  load_states(state, n_streams, deterministic != 0, st);
  input p[MB200_MAX_PARAMS] = {};
  const int np = ref_n_params(dist);
  if (np < 0) { g_err = "unknown distribution id"; return -2; }
  for (int k = 0; k < np; ++k) {
    if (param_len[k] != 1 && param_len[k] != n_streams) {
      g_err = "bad parameter length"; return -2;
    }
    p[k] = input{params[k], param_len[k] == 1};
  }
  int rc = 0;
  try {
    if (real_type == MB200_REAL_F32) {
      sample_impl<float>(dist, st, n_samples, p, out, n_threads);
    } else {
      sample_impl<double>(dist, st, n_samples, p, out, n_threads);
    }
  } catch (const std::exception& e) { g_err = e.what(); rc = -3; }
  store_states(st, state);
  return rc;
}

// ref: src/random.cpp:884-934
int ref_multinomial(uint64_t* state, size_t n_streams, size_t n_samples,
                    const double* size, size_t size_len,
                    const double* prob, size_t prob_len, size_t prob_cols,
                    double* out) {
  std::vector<rng256p> st;
  load_states(state, n_streams, false, st);
  int rc = 0;
  try {
    for (size_t i = 0; i < n_streams; ++i) {
      const double* pr = prob + (prob_cols == 1 ? 0 : i * prob_len);
      const double sz = size[size_len == 1 ? 0 : i];
      for (size_t j = 0; j < n_samples; ++j) {
        double* y = out + prob_len * (n_samples * i + j);
        mr::multinomial<double>(st[i], sz, pr, static_cast<int>(prob_len), y);
      }
    }
  } catch (const std::exception& e) { g_err = e.what(); rc = -3; }
  store_states(st, state);
  return rc;
}

// n_samples next() per stream with the given scrambler on a 4 x u64 state
int ref_raw(int gen, uint64_t* state, size_t n_streams, size_t n_samples,
            uint64_t* out) {
  switch (gen) {
  case MB200_GEN_XOSHIRO256PLUS: raw_impl<mr::xoshiro256plus>(state, n_streams, n_samples, out); return 0;
  case MB200_GEN_XOSHIRO256PLUSPLUS: raw_impl<mr::xoshiro256plusplus>(state, n_streams, n_samples, out); return 0;
  case MB200_GEN_XOSHIRO256STARSTAR: raw_impl<mr::xoshiro256starstar>(state, n_streams, n_samples, out); return 0;
  default: g_err = "raw: generator must be a xoshiro256 variant"; return -2;
  }
}

// golden protocol for all 12 generators                ref: src/test_rng.cpp:21-38
int ref_xoshiro_run(int gen, uint64_t seed, int n, uint64_t* values,
                    uint64_t* state_out, int* n_words) {
  switch (gen) {
  case MB200_GEN_XOSHIRO256PLUS: run_protocol<mr::xoshiro256plus>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOSHIRO256PLUSPLUS: run_protocol<mr::xoshiro256plusplus>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOSHIRO256STARSTAR: run_protocol<mr::xoshiro256starstar>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOSHIRO128PLUS: run_protocol<mr::xoshiro128plus>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOSHIRO128PLUSPLUS: run_protocol<mr::xoshiro128plusplus>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOSHIRO128STARSTAR: run_protocol<mr::xoshiro128starstar>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOROSHIRO128PLUS: run_protocol<mr::xoroshiro128plus>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOROSHIRO128PLUSPLUS: run_protocol<mr::xoroshiro128plusplus>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOROSHIRO128STARSTAR: run_protocol<mr::xoroshiro128starstar>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOSHIRO512PLUS: run_protocol<mr::xoshiro512plus>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOSHIRO512PLUSPLUS: run_protocol<mr::xoshiro512plusplus>(seed, n, values, state_out, n_words); break;
  case MB200_GEN_XOSHIRO512STARSTAR: run_protocol<mr::xoshiro512starstar>(seed, n, values, state_out, n_words); break;
  default: g_err = "unknown generator id"; return -2;
  }
  return 0;
}

// ref: src/density.cpp:6-254
int ref_density(int dens, int real_type, size_t n, const void* x, int x_is_i32,
                const void* const* params, const int* param_is_i32, int lg,
                double* out, int n_threads) {
  try {
    if (real_type == MB200_REAL_F32) {
      density_impl<float>(dens, n, x, x_is_i32, params, param_is_i32, lg, out, n_threads);
    } else {
      density_impl<double>(dens, n, x, x_is_i32, params, param_is_i32, lg, out, n_threads);
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -3; }
}

// table dump used by tools/extract_tables.py to generate the constant tables
// (data, not source) the product and the C oracle carry
void ref_tables(double* zig_x /*257*/, double* zig_y /*256*/,
                float* tail_f /*128*/, double* tail_d /*10*/) {
  for (int i = 0; i < 257; ++i) zig_x[i] = mr::ziggurat::x[i];
  for (int i = 0; i < 256; ++i) zig_y[i] = mr::ziggurat::y[i];
  for (int i = 0; i < 128; ++i) tail_f[i] = mr::k_tail_values_f[i];
  for (int i = 0; i < 10; ++i) tail_d[i] = mr::k_tail_values_d[i];
}

// The stochastic SIR model and bootstrap particle filter of the reference's
// own test-suite (tests/testthat/helper-sir-filter.R:56-117, an R script),
// restated over the reference's C++ samplers and densities: per particle two
// binomial draws per step, an exponential observation noise and a Poisson
// log-density per observation; systematic resampling with one uniform from
// the filter stream (dust_resample_weight, :113-117).  Checker for
// examples/sir_filter.cu.
int ref_sir_filter(uint64_t* system_state, size_t n, uint64_t* filter_state,
                   double N, double I0, double beta, double gamma, double exp_noise,
                   const double* data_time, const double* data_incidence, size_t n_data,
                   int resample, double* ll_out, double* final_state /* [4][n] */) {
  try {
    std::vector<rng256p> sys, fil;
    load_states(system_state, n, false, sys);
    load_states(filter_state, 1, false, fil);
    std::vector<double> S(n, N - I0), I(n, I0), R(n, 0.0), cases(n, 0.0), logw(n), w(n), cum(n);
    std::vector<double> S2(n), I2(n), R2(n), C2(n);
    const double dt = 1;
    double time = 0, ll = 0;
    for (size_t d = 0; d < n_data; ++d) {
      const double to = data_time[d];
      for (size_t i = 0; i < n; ++i) {
        for (double t = time + dt; t <= to; t += dt) {                    // sir_run / sir_step
          const double p_SI = 1 - std::exp(-beta * I[i] / N * dt);
          const double p_IR = 1 - std::exp(-gamma * dt);
          const double n_SI = mr::binomial<double>(sys[i], S[i], p_SI);
          const double n_IR = mr::binomial<double>(sys[i], I[i], p_IR);
          const double carried = std::fmod(t, 1.0) == 0 ? 0 : cases[i];
          S[i] = S[i] - n_SI;
          I[i] = I[i] + n_SI - n_IR;
          R[i] = R[i] + n_IR;
          cases[i] = carried + n_SI;
        }
        const double noise = mr::exponential_rate<double>(sys[i], exp_noise);
        logw[i] = md::poisson(data_incidence[d], cases[i] + noise, true);
      }
      double m = -INFINITY, total = 0;
      for (size_t i = 0; i < n; ++i) m = std::max(m, logw[i]);
      for (size_t i = 0; i < n; ++i) { w[i] = std::exp(logw[i] - m); total += w[i]; }
      ll += std::log(total / static_cast<double>(n)) + m;
      const double u = mr::random_real<double>(fil[0]);
      if (resample) {
        double run = 0;
        for (size_t i = 0; i < n; ++i) { run += w[i] / total; cum[i] = run; }
        const double nn = static_cast<double>(n);
        for (size_t j = 0; j < n; ++j) {
          const double x = u / nn + static_cast<double>(j) * (1 / nn);
          size_t k = std::upper_bound(cum.begin(), cum.end(), x) - cum.begin();   // findInterval
          if (k >= n) k = n - 1;
          S2[j] = S[k]; I2[j] = I[k]; R2[j] = R[k]; C2[j] = cases[k];
        }
        S.swap(S2); I.swap(I2); R.swap(R2); cases.swap(C2);
      }
      time = to;
    }
    store_states(sys, system_state);
    store_states(fil, filter_state);
    *ll_out = ll;
    if (final_state) {
      for (size_t i = 0; i < n; ++i) {
        final_state[i] = S[i]; final_state[n + i] = I[i]; final_state[2 * n + i] = R[i]; final_state[3 * n + i] = cases[i];
      }
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

}  // extern "C"
