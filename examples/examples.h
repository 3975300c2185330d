/* examples.h -- C entry points of monty_b200/libmonty_b200_examples.so:
 * consumers of the device-side monty::random API (include/monty_b200/random.cuh)
 * written the way dust2 / odin2 models consume the reference's headers. */
#ifndef MONTY_B200_EXAMPLES_H
#define MONTY_B200_EXAMPLES_H
#include <stddef.h>
#include <stdint.h>

#include "../include/monty_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One inline draw per stream of distribution `dist` (MB200_DIST_*) through
 * monty::random::<dist><real_type>(state, params...): the same draws as
 * mb200_rng_sample(..., n_samples = 1, ...).  params_host[k] has length 1 or
 * n_streams.  out_host receives n_streams doubles. */
int mb200_example_inline_draw(mb200_rng* rng, int dist, int real_type,
                              const double* const* params_host, const size_t* param_len,
                              double* out_host);

/* The device-side API on the reference's other generators (generator_draw.cu): n_streams
 * streams seeded like prng<G>(n_streams, seed) by monty_b200::device_prng<G>, then `jumps`
 * jump()s and `long_jumps` long_jump()s of every stream, then n_samples draws of `dist` per
 * stream through monty::random::<dist><real_type>(state, ...).  gen (MB200_GEN_*):
 * XOSHIRO128PLUS (f32: generator<float>; f64 too), XOSHIRO256PLUS (generator<double>; f32
 * too), XOSHIRO128STARSTAR (f32), XOSHIRO256STARSTAR / XOROSHIRO128PLUSPLUS /
 * XOSHIRO512STARSTAR (f64).  out_host: n_streams x n_samples doubles, stream-major;
 * state_out (optional): the final states, int_type[n_streams][state words] as exported. */
int mb200_example_generator_draw(int gen, int dist, int real_type, uint64_t seed, size_t n_streams,
                                 size_t n_samples, int jumps, int long_jumps,
                                 const double* const* params_host, const size_t* param_len,
                                 double* out_host, void* state_out);

/* monty_b200::device_prng<G>(n_streams, seed words) -- the reference's seed-vector constructor
 * (prng.hpp:38-53) -- exported; with reimport_and_jump the states are imported into a second
 * container, jumped once and exported from there.  gen: XOSHIRO128PLUS (uint32 words),
 * XOSHIRO256PLUS, XOROSHIRO128PLUSPLUS, XOSHIRO512STARSTAR (uint64 words). */
int mb200_example_generator_seed_vector(int gen, const void* words, size_t n_words, size_t n_streams,
                                        int reimport_and_jump, void* state_out);

/* Bootstrap particle filter of the stochastic SIR model of the reference's
 * test-suite (tests/testthat/helper-sir-filter.R:56-117): n_particles =
 * n_streams of `system`, one stream in `filter` for the systematic
 * resampling.  Between observations every particle runs its dt = 1 steps in
 * registers: two binomial draws per step, then the exponential observation
 * noise and the Poisson log-density, fused in one kernel.
 * data_time/data_incidence: n_data observations (times are integers, increasing).
 * final_state_host: optional, 4 x n_particles doubles (S, I, R, cases per particle).
 * resample = 0 skips the resampling (trajectory parity tests). */
int mb200_example_sir_filter(mb200_rng* system, mb200_rng* filter,
                             double N, double I0, double beta, double gamma, double exp_noise,
                             const double* data_time, const double* data_incidence, size_t n_data,
                             int resample, double* log_likelihood, double* final_state_host);

/* One multinomial draw per stream through monty::random::multinomial<real_type>
 * (state, size, prob, prob_len, ret) inside a kernel: the same draws as
 * mb200_rng_multinomial(..., n_samples = 1, ...).  size_host: 1 or n_streams
 * values; prob_host: prob_len (<= 16) values per column, prob_cols = 1 or
 * n_streams columns; out_host: n_streams x prob_len doubles, stream-major.
 * Invalid input gives a row of NaN (the reference's fatal error). */
int mb200_example_inline_multinomial(mb200_rng* rng, int real_type,
                                     const double* size_host, size_t size_len,
                                     const double* prob_host, int prob_len, size_t prob_cols,
                                     double* out_host);

/* Evaluates one of the samplers' own math routines (include/monty_b200/detail/samplers.cuh) on n
 * host doubles, for the parity tests: fn 0 = m_log, 1 = m_log_normal
 * (positive normal arguments), 2 = cos_0_2pi (0 <= x <= 2 pi), 3 = sqrt_normal
 * (positive normal arguments), 4 = the toolkit's sqrt, 5 = 1.0 when the examples
 * library was compiled without FMA contraction (monty::random::contraction_is_off),
 * 6 = m_lgamma, 7 = m_exp, 8 = m_pow(x[i], x[i ^ 1]) (n even). */
int mb200_example_math_probe(int fn, const double* x_host, double* y_host, size_t n);

/* monty::density::zi_poisson (which = 0; a = lambda), zi_negative_binomial_prob
 * (1; a = size, b = prob), zi_negative_binomial_mu (2; a = size, b = mu) through
 * the device-side API, elementwise on n host doubles. */
int mb200_example_zi_density(int which, const double* x_host, const double* pi0_host, const double* a_host,
                             const double* b_host, int lg, double* y_host, size_t n);

#ifdef __cplusplus
}
#endif
#endif
