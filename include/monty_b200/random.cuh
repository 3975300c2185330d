// monty_b200/random.cuh -- the device-side `monty::random` / `monty::density`
// API: the per-stream samplers and log-densities as __device__ functions with
// the reference's names, template parameters and argument order, for kernels
// that draw inline (a dust2 / odin2 style model `update`) instead of through
// the bulk C-ABI calls of include/monty_b200.h.
//
// Mirrors (reference file:line, "hdr/" = inst/include/monty/random/):
//   monty::random::generator<real_type>            hdr/random.hpp:31-47
//   seed / jump / long_jump                        hdr/generator.hpp:59-141
//   random_int / random_real                       hdr/generator.hpp:154-183
//   uniform, exponential_rate / _mean, normal / random_normal (box_muller,
//   polar, ziggurat), binomial, poisson, gamma_scale / _rate, beta,
//   negative_binomial_prob / _mu, hypergeometric, truncated_normal, cauchy,
//   weibull, log_normal, beta_binomial_prob / _ab, zi_poisson,
//   zi_negative_binomial_prob / _mu                hdr/<dist>.hpp
//   monty::density::<dist>(x, params..., log)      inst/include/monty/density.hpp
// The C++ names are the ones odin2 generates (R/dsl-distributions.R:57-340).
//
// Every function draws from the caller's state exactly the words the
// reference draws, in the same order, and computes the same operations in the
// same order (detail/samplers.cuh; compile with -fmad=false like the library:
// the reference is built without FMA contraction).  The state type holds the
// four words of xoshiro256+ in registers; `state[i]` reads and writes them in
// the reference's exported layout (hdr/prng.hpp:92-105), which is also the
// layout of the device array behind monty_b200::prng (prng.hpp in this
// directory) -- load with load_state(), store with store_state().
//
// Differences from the reference, all forced by the device:
//   * no deterministic mode (the reference compiles it out of device code too:
//     `#ifndef __CUDA_ARCH__`);
//   * invalid parameters: the reference prints and __trap()s, which kills the
//     CUDA context; here the draw returns NaN (define MONTY_B200_TRAP_ON_ERROR
//     for the reference's behaviour).
// `generator<float>` is xoshiro128+ and `generator<double>` xoshiro256+, as in the
// reference (hdr/random.hpp:31-47); every sampler works on any of the twelve state
// types.  (The bulk C-ABI behind the R layer is xoshiro256+ for every real_type, like
// src/random.cpp:10.)
#pragma once
#include <cstdio>

#include "detail/samplers.cuh"
#include "detail/generators.cuh"
#include "detail/density.cuh"
#include "detail/jump_constants.cuh"

namespace monty {
namespace random {

namespace detail {
#include "detail/monty_tables.inc"
static __device__ const double f_zig_x[257] = { MONTY_ZIG_X_VALUES };
static __device__ const double2 f_zig_xy[256] = { MONTY_ZIG_XY_VALUES };
static __device__ const double f_tail_d[10] = { MONTY_TAIL_D_VALUES };
static __device__ const float f_tail_f[128] = { MONTY_TAIL_F_VALUES };

// tables read straight from global memory (L1-cached): inline draws have no
// shared-memory staging phase
__device__ __forceinline__ mb200::Ctx ctx() {
  mb200::Ctx c;
  c.zig_xy_saddr = 0;
  c.zig_xy = f_zig_xy;
  c.zig_x = f_zig_x;
  c.zig_x2s = nullptr;
  c.tail_d = f_tail_d;
  c.tail_f = f_tail_f;
  c.err = 0;
  return c;
}

template <typename T> __device__ __forceinline__ T nan_value();
template <> __device__ __forceinline__ double nan_value<double>() { return __longlong_as_double(0x7ff8000000000000ll); }
template <> __device__ __forceinline__ float nan_value<float>() { return __int_as_float(0x7fc00000); }

__device__ __forceinline__ void fail(const char* what) {
#ifdef MONTY_B200_TRAP_ON_ERROR
  printf("Invalid call to %s\n", what);
  __trap();
#else
  (void)what;
#endif
}
}  // namespace detail

// True when this translation unit really was compiled without FMA contraction
// (the #error in detail/samplers.cuh only checks that the consumer SAID so):
// a * b + c is evaluated on values for which the fused and the two-rounding
// results differ.  Call it once from a test kernel of the consuming project.
__device__ __forceinline__ bool contraction_is_off() {
  volatile double a = 1.0 + 0x1p-30, b = 1.0 - 0x1p-30, c = -1.0;
  const double x = a, y = b, z = c;
  return x * y + z == 0.0;          // exact product is 1 - 2^-60: rounds to 1, fused gives -2^-60
}

// ---- generator state -------------------------------------------------------
// hdr/xoshiro_state.hpp:20-39: the words of one stream, `state[i]` in the reference's
// exported order (hdr/prng.hpp:92-105).  All twelve generators of the reference
// (hdr/xoshiro128.hpp, xoroshiro128.hpp, xoshiro256.hpp, xoshiro512.hpp).
template <class State, int SCR> struct xoshiro_state {
  using int_type = typename mb200::GenTraits<State>::int_type;
  using state_type = State;
  static constexpr int scrambler = SCR;
  State s;
  __host__ __device__ static constexpr size_t size() { return mb200::GenTraits<State>::N; }
  __host__ __device__ int_type& operator[](size_t i) { return mb200::gen_words(s)[i]; }
  __host__ __device__ const int_type& operator[](size_t i) const { return mb200::gen_words(const_cast<State&>(s))[i]; }
};
using xoshiro256plus = xoshiro_state<mb200::X256, mb200::kPlus>;
using xoshiro256plusplus = xoshiro_state<mb200::X256, mb200::kPlusPlus>;
using xoshiro256starstar = xoshiro_state<mb200::X256, mb200::kStarStar>;
using xoshiro128plus = xoshiro_state<mb200::X128, mb200::kPlus>;
using xoshiro128plusplus = xoshiro_state<mb200::X128, mb200::kPlusPlus>;
using xoshiro128starstar = xoshiro_state<mb200::X128, mb200::kStarStar>;
using xoroshiro128plus = xoshiro_state<mb200::XR128, mb200::kPlus>;
using xoroshiro128plusplus = xoshiro_state<mb200::XR128, mb200::kPlusPlus>;
using xoroshiro128starstar = xoshiro_state<mb200::XR128, mb200::kStarStar>;
using xoshiro512plus = xoshiro_state<mb200::X512, mb200::kPlus>;
using xoshiro512plusplus = xoshiro_state<mb200::X512, mb200::kPlusPlus>;
using xoshiro512starstar = xoshiro_state<mb200::X512, mb200::kStarStar>;

// hdr/random.hpp:31-47: xoshiro128+ for float, xoshiro256+ otherwise
template <typename real_type> struct default_rng_helper { using type = xoshiro256plus; };
template <> struct default_rng_helper<float> { using type = xoshiro128plus; };
template <typename real_type> using generator = typename default_rng_helper<real_type>::type;

namespace detail {
// the samplers' view of a state type: xoshiro256+ takes the batch kernels' generator
// (written on 32-bit halves), the others the plain restatements of generators.cuh
template <class G> struct rng_of { using type = mb200::GenRng<typename G::state_type, G::scrambler>; };
template <> struct rng_of<xoshiro256plus> { using type = mb200::Rng; };
}  // namespace detail

// state <-> the device array of monty_b200::prng (uint64[n_streams][4]),
// 2 x 128-bit accesses per stream
__device__ __forceinline__ xoshiro256plus load_state(const uint64_t* states, size_t stream) {
  return xoshiro256plus{mb200::load_state(states, stream)};
}
__device__ __forceinline__ void store_state(uint64_t* states, size_t stream, const xoshiro256plus& g) {
  mb200::store_state(states, stream, g.s);
}
// any generator: int_type[n_streams][size()], the layout of monty_b200::device_prng<G>
// (device_prng.cuh) and of the reference's export_state()
template <class G>
__device__ __forceinline__ G load_state_of(const typename G::int_type* states, size_t stream) {
  G g;
  for (size_t i = 0; i < G::size(); ++i) g[i] = states[stream * G::size() + i];
  return g;
}
template <class G>
__device__ __forceinline__ void store_state_of(typename G::int_type* states, size_t stream, const G& g) {
  for (size_t i = 0; i < G::size(); ++i) states[stream * G::size() + i] = g[i];
}

// hdr/utils.hpp:72-77, hdr/generator.hpp:114-141
__host__ __device__ inline uint64_t splitmix64(uint64_t seed) {
  uint64_t z = (seed + 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
template <typename rng_state_type>
__host__ __device__ inline rng_state_type seed(uint64_t seed) {
  rng_state_type g;
  for (size_t i = 0; i < rng_state_type::size(); ++i) {
    seed = splitmix64(seed);
    g[i] = static_cast<typename rng_state_type::int_type>(seed);
  }
  return g;
}

// hdr/generator.hpp:59-106
template <class State, int SCR>
__host__ __device__ inline void jump(xoshiro_state<State, SCR>& g) { mb200::gen_jump<State, SCR, false>(g.s); }
template <class State, int SCR>
__host__ __device__ inline void long_jump(xoshiro_state<State, SCR>& g) { mb200::gen_jump<State, SCR, true>(g.s); }

// hdr/generator.hpp:154-183
template <typename T, typename rng_state_type>
__device__ __forceinline__ T random_int(rng_state_type& g) {
  static_assert(sizeof(T) <= sizeof(typename rng_state_type::int_type), "requested integer too wide");
  typename detail::rng_of<rng_state_type>::type r{g.s};
  const auto value = r.next();
  g.s = r.s;
  return static_cast<T>(value);
}
template <typename real_type, typename rng_state_type>
__device__ __forceinline__ real_type random_real(rng_state_type& g) {
  typename detail::rng_of<rng_state_type>::type r{g.s};
  const real_type v = mb200::random_real<real_type>(r);
  g.s = r.s;
  return v;
}

namespace detail {
// prepare + one draw of sampler D on the caller's state
template <class D, typename real_type, typename rng_state_type, int NP>
__device__ __forceinline__ real_type draw(rng_state_type& g, const real_type (&par)[NP], const char* what) {
  mb200::Ctx c = ctx();
  typename D::Prep q;
  if (D::prepare(par, q, c) != mb200::kOk) {
    fail(what);
    return nan_value<real_type>();
  }
  typename rng_of<rng_state_type>::type r{g.s};
  const real_type v = D::draw(r, q, c);
  g.s = r.s;
  if (D::DYN_ERR && c.err) {
    fail(what);
    return nan_value<real_type>();
  }
  return v;
}
}  // namespace detail

namespace algorithm {
enum class normal { box_muller, polar, ziggurat };    // hdr/normal.hpp:14-20
}

#define MONTY_B200_SAMPLER(NAME, DIST, NP, PARAMS, ARGS)                                   \
  template <typename real_type, typename rng_state_type>                                   \
  __device__ __forceinline__ real_type NAME(rng_state_type& g, PARAMS) {                   \
    const real_type par[NP] = {ARGS};                                                      \
    return detail::draw<DIST, real_type, rng_state_type, NP>(g, par, #NAME);               \
  }
#define P1(a) real_type a
#define P2(a, b) real_type a, real_type b
#define P3(a, b, c) real_type a, real_type b, real_type c
#define P4(a, b, c, d) real_type a, real_type b, real_type c, real_type d
#define A2(a, b) a, b
#define A3(a, b, c) a, b, c
#define A4(a, b, c, d) a, b, c, d
#define D1(X) mb200::X<real_type>
#define D2(X, F) mb200::X<real_type, F>

MONTY_B200_SAMPLER(uniform, D1(DistUniform), 2, P2(min, max), A2(min, max))                         // hdr/uniform.hpp:23-35
MONTY_B200_SAMPLER(exponential_rate, D2(DistExponential, false), 1, P1(rate), rate)                 // hdr/exponential.hpp:34-46
MONTY_B200_SAMPLER(exponential_mean, D2(DistExponential, true), 1, P1(mean), mean)                  // hdr/exponential.hpp:49-61
MONTY_B200_SAMPLER(binomial, D1(DistBinomial), 2, P2(n, p), A2(n, p))                               // hdr/binomial.hpp:284-299
MONTY_B200_SAMPLER(poisson, D1(DistPoisson), 1, P1(lambda), lambda)                                 // hdr/poisson.hpp:209-240
MONTY_B200_SAMPLER(gamma_scale, D2(DistGamma, false), 2, P2(shape, scale), A2(shape, scale))        // hdr/gamma.hpp:82-111
MONTY_B200_SAMPLER(gamma_rate, D2(DistGamma, true), 2, P2(shape, rate), A2(shape, rate))            // hdr/gamma.hpp:113-121
MONTY_B200_SAMPLER(beta, D1(DistBeta), 2, P2(a, b), A2(a, b))                                       // hdr/beta.hpp:12-28
MONTY_B200_SAMPLER(negative_binomial_prob, D2(DistNegativeBinomial, false), 2, P2(size, prob), A2(size, prob))   // hdr/negative_binomial.hpp:30-41
MONTY_B200_SAMPLER(negative_binomial_mu, D2(DistNegativeBinomial, true), 2, P2(size, mu), A2(size, mu))          // hdr/negative_binomial.hpp:43-52
MONTY_B200_SAMPLER(hypergeometric, D1(DistHypergeometric), 3, P3(n1, n2, k), A3(n1, n2, k))        // hdr/hypergeometric.hpp:343-366
MONTY_B200_SAMPLER(truncated_normal, D1(DistTruncatedNormal), 4, P4(mean, sd, min, max), A4(mean, sd, min, max)) // hdr/truncated_normal.hpp:89-95
MONTY_B200_SAMPLER(cauchy, D1(DistCauchy), 2, P2(location, scale), A2(location, scale))            // hdr/cauchy.hpp:12-25
MONTY_B200_SAMPLER(weibull, D1(DistWeibull), 2, P2(shape, scale), A2(shape, scale))                // hdr/weibull.hpp:26-38
MONTY_B200_SAMPLER(log_normal, D1(DistLogNormal), 2, P2(mulog, sdlog), A2(mulog, sdlog))           // hdr/log_normal.hpp:26-37
MONTY_B200_SAMPLER(beta_binomial_prob, D2(DistBetaBinomial, true), 3, P3(size, prob, rho), A3(size, prob, rho))  // hdr/beta_binomial.hpp:40-49
MONTY_B200_SAMPLER(beta_binomial_ab, D2(DistBetaBinomial, false), 3, P3(size, a, b), A3(size, a, b))             // hdr/beta_binomial.hpp:27-38
MONTY_B200_SAMPLER(zi_poisson, D1(DistZiPoisson), 2, P2(pi0, lambda), A2(pi0, lambda))             // hdr/zi_poisson.hpp:27-41
MONTY_B200_SAMPLER(zi_negative_binomial_prob, D2(DistZiNegativeBinomial, false), 3, P3(pi0, size, prob), A3(pi0, size, prob))  // hdr/zi_negative_binomial.hpp:32-46
MONTY_B200_SAMPLER(zi_negative_binomial_mu, D2(DistZiNegativeBinomial, true), 3, P3(pi0, size, mu), A3(pi0, size, mu))        // hdr/zi_negative_binomial.hpp:48-58

#undef MONTY_B200_SAMPLER
#undef P1
#undef P2
#undef P3
#undef P4
#undef A2
#undef A3
#undef A4
#undef D1
#undef D2

// hdr/multinomial.hpp:53-82: chained conditional binomials; `prob` need not be
// normalised; `ret` receives prob_len counts.  On invalid input (a negative
// probability, no positive one, or an inner binomial the reference would
// reject) every entry of `ret` is NaN and the state is left where the failure
// happened, unless MONTY_B200_TRAP_ON_ERROR turns it into the reference's
// fatal error.
template <typename real_type, typename rng_state_type, typename T, typename U>
__device__ inline void multinomial(rng_state_type& g, real_type size, const T& prob, int prob_len, U ret) {
  real_type p_tot = 0;
  bool ok = true;
  for (int i = 0; i < prob_len; ++i) {
    if (prob[i] < 0) ok = false;
    p_tot += prob[i];
  }
  if (!ok || p_tot == 0) {
    detail::fail("multinomial");
    for (int i = 0; i < prob_len; ++i) ret[i] = detail::nan_value<real_type>();
    return;
  }
  for (int i = 0; i < prob_len - 1; ++i) {
    if (prob[i] > 0) {
      const real_type q = static_cast<real_type>(prob[i]) / p_tot;
      const real_type pi = q < static_cast<real_type>(1) ? q : static_cast<real_type>(1);
      const real_type draw = binomial<real_type>(g, size, pi);
      if (draw != draw) {
        for (int k = 0; k < prob_len; ++k) ret[k] = detail::nan_value<real_type>();
        return;
      }
      ret[i] = draw;
      size -= draw;
      p_tot -= prob[i];
    } else {
      ret[i] = 0;
    }
  }
  ret[prob_len - 1] = size;
}

// hdr/normal.hpp:23-93: the algorithm is a template parameter, box_muller the
// default (which gamma and truncated_normal use implicitly)
template <typename real_type, algorithm::normal A = algorithm::normal::box_muller, typename rng_state_type>
__device__ __forceinline__ real_type random_normal(rng_state_type& g) {
  typename detail::rng_of<rng_state_type>::type r{g.s};
  real_type z;
  if (A == algorithm::normal::box_muller) {
    z = mb200::std_normal_box_muller<real_type>(r);
  } else if (A == algorithm::normal::polar) {
    z = mb200::std_normal_polar<real_type>(r);
  } else {
    const mb200::Ctx c = detail::ctx();
    z = mb200::std_normal_ziggurat<real_type>(r, c);
  }
  g.s = r.s;
  return z;
}
template <typename real_type, algorithm::normal A = algorithm::normal::box_muller, typename rng_state_type>
__device__ __forceinline__ real_type normal(rng_state_type& g, real_type mean, real_type sd) {
  return random_normal<real_type, A>(g) * sd + mean;
}

}  // namespace random

// ---- log-densities (inst/include/monty/density.hpp) ---------------------------
namespace density {
#define MONTY_B200_DENSITY(NAME, IMPL, PARAMS, ARGS)                                        \
  template <typename T> __device__ __forceinline__ T NAME(T x, PARAMS, bool log) {          \
    return mb200::IMPL<T>(x, ARGS, log);                                                    \
  }
#define Q1(a) T a
#define Q2(a, b) T a, T b
#define Q3(a, b, c) T a, T b, T c
#define Q4(a, b, c, d) T a, T b, T c, T d
#define B2(a, b) a, b
#define B3(a, b, c) a, b, c
#define B4(a, b, c, d) a, b, c, d
MONTY_B200_DENSITY(binomial, dens_binomial, Q2(size, prob), B2(size, prob))
MONTY_B200_DENSITY(normal, dens_normal, Q2(mu, sd), B2(mu, sd))
MONTY_B200_DENSITY(negative_binomial_mu, dens_negative_binomial_mu, Q2(size, mu), B2(size, mu))
MONTY_B200_DENSITY(negative_binomial_prob, dens_negative_binomial_prob, Q2(size, prob), B2(size, prob))
MONTY_B200_DENSITY(beta_binomial_ab, dens_beta_binomial_ab, Q3(size, a, b), B3(size, a, b))
MONTY_B200_DENSITY(beta_binomial_prob, dens_beta_binomial_prob, Q3(size, prob, rho), B3(size, prob, rho))
MONTY_B200_DENSITY(poisson, dens_poisson, Q1(lambda), lambda)
MONTY_B200_DENSITY(exponential_rate, dens_exponential_rate, Q1(rate), rate)
MONTY_B200_DENSITY(exponential_mean, dens_exponential_mean, Q1(mean), mean)
MONTY_B200_DENSITY(gamma_rate, dens_gamma_rate, Q2(shape, rate), B2(shape, rate))
MONTY_B200_DENSITY(gamma_scale, dens_gamma_scale, Q2(shape, scale), B2(shape, scale))
MONTY_B200_DENSITY(uniform, dens_uniform, Q2(min, max), B2(min, max))
MONTY_B200_DENSITY(hypergeometric, dens_hypergeometric, Q3(n1, n2, k), B3(n1, n2, k))
MONTY_B200_DENSITY(beta, dens_beta, Q2(a, b), B2(a, b))
MONTY_B200_DENSITY(truncated_normal, dens_truncated_normal, Q4(mu, sd, min, max), B4(mu, sd, min, max))
MONTY_B200_DENSITY(cauchy, dens_cauchy, Q2(location, scale), B2(location, scale))
MONTY_B200_DENSITY(weibull, dens_weibull, Q2(shape, scale), B2(shape, scale))
MONTY_B200_DENSITY(log_normal, dens_log_normal, Q2(mulog, sdlog), B2(mulog, sdlog))
#undef MONTY_B200_DENSITY
// inst/include/monty/random/density.hpp:355-409: zero-inflated count densities,
// log(pi0 + (1 - pi0) f(0)) at x == 0 and log(1 - pi0) + log f(x) elsewhere
template <typename T> __device__ __forceinline__ T zi_poisson(T x, T pi0, T lambda, bool log) {
  return mb200::dens_zi_wrap<T>(x, pi0, mb200::dens_poisson<T>(x, lambda, true), log);
}
template <typename T> __device__ __forceinline__ T zi_negative_binomial_prob(T x, T pi0, T size, T prob, bool log) {
  return mb200::dens_zi_wrap<T>(x, pi0, mb200::dens_negative_binomial_prob<T>(x, size, prob, true), log);
}
template <typename T> __device__ __forceinline__ T zi_negative_binomial_mu(T x, T pi0, T size, T mu, bool log) {
  return mb200::dens_zi_wrap<T>(x, pi0, mb200::dens_negative_binomial_mu<T>(x, size, mu, true), log);
}
#undef Q1
#undef Q2
#undef Q3
#undef Q4
#undef B2
#undef B3
#undef B4
}  // namespace density
}  // namespace monty
