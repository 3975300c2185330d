/*
 * monty_b200.h -- C-ABI of the B200-native implementation of monty's
 * multi-stream RNG / sampler / log-density hot path.
 *
 * This is the drop-in boundary: every entry point below replaces one
 * function (or one family of functions) of the reference's cpp11 layer,
 * cited as "ref: file:line" relative to the mrc-ide/monty source tree.
 * The cpp11 glue that a maintainer of the reference would write on top of
 * these is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary.  Every function
 *     returns MB200_OK (0) or a negative error code; the message (with the
 *     reference's exact text where the reference has one) is available from
 *     mb200_last_error() (thread-local) and, for handle-bound calls, from
 *     mb200_rng_last_error(handle).
 *   - pointers named *_host are host memory (R vectors); pointers named
 *     *_dev are device memory on the handle's device.
 *   - stream state is held on the device in the reference's exact layout:
 *     n_streams x 4 x uint64 little-endian, stream-major
 *     (ref: inst/include/monty/random/prng.hpp:92-105, src/random.cpp:356-364).
 *   - there is no CPU fallback: if no CUDA device is usable, creation fails
 *     with MB200_ERR_CUDA.
 */
#ifndef MONTY_B200_H
#define MONTY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------ */
enum {
  MB200_OK = 0,
  MB200_ERR_CUDA = -1,      /* CUDA runtime / driver failure, no device      */
  MB200_ERR_ARG = -2,       /* bad argument at the boundary (length, NULL)   */
  MB200_ERR_PARAM = -3,     /* invalid distribution parameter (reference's
                               fatal_error text in last_error)               */
  MB200_ERR_STATE = -4,     /* bad state / seed length                       */
  MB200_ERR_UNSUPPORTED = -5
};

/* ---- real_type selector (ref: template parameter real_type of every
 *      monty::random::<dist><real_type>) --------------------------------- */
enum {
  MB200_REAL_F64 = 0,
  MB200_REAL_F32 = 1
};

/* ---- generator selector for the raw / seeding entry points
 *      (ref: inst/include/monty/random/xoshiro256.hpp:14-19 and the 128/512
 *      siblings).  Samplers run on XOSHIRO256PLUS, the reference's
 *      default_rng64 (ref: src/random.cpp:10). ---------------------------- */
enum {
  MB200_GEN_XOSHIRO256PLUS = 0,
  MB200_GEN_XOSHIRO256PLUSPLUS = 1,
  MB200_GEN_XOSHIRO256STARSTAR = 2,
  MB200_GEN_XOSHIRO128PLUS = 3,
  MB200_GEN_XOSHIRO128PLUSPLUS = 4,
  MB200_GEN_XOSHIRO128STARSTAR = 5,
  MB200_GEN_XOROSHIRO128PLUS = 6,
  MB200_GEN_XOROSHIRO128PLUSPLUS = 7,
  MB200_GEN_XOROSHIRO128STARSTAR = 8,
  MB200_GEN_XOSHIRO512PLUS = 9,
  MB200_GEN_XOSHIRO512PLUSPLUS = 10,
  MB200_GEN_XOSHIRO512STARSTAR = 11,
  MB200_GEN_COUNT = 12
};

/* ---- distribution ids.  One per sampler family behind
 *      monty_random_<dist> / monty_random_n_<dist>
 *      (ref: src/random.cpp:399-882).  The parameter order is the order of
 *      the reference's R-level arguments. --------------------------------- */
enum {
  MB200_DIST_REAL = 0,                  /* ()                 random.cpp:401 */
  MB200_DIST_UNIFORM = 1,               /* min, max           uniform.hpp:23 */
  MB200_DIST_EXPONENTIAL_RATE = 2,      /* rate           exponential.hpp:44 */
  MB200_DIST_EXPONENTIAL_MEAN = 3,      /* mean           exponential.hpp:54 */
  MB200_DIST_NORMAL_BOX_MULLER = 4,     /* mean, sd            normal.hpp:77 */
  MB200_DIST_NORMAL_POLAR = 5,          /* mean, sd                          */
  MB200_DIST_NORMAL_ZIGGURAT = 6,       /* mean, sd                          */
  MB200_DIST_BINOMIAL = 7,              /* size, prob       binomial.hpp:284 */
  MB200_DIST_POISSON = 8,               /* lambda            poisson.hpp:209 */
  MB200_DIST_GAMMA_SCALE = 9,           /* shape, scale         gamma.hpp:82 */
  MB200_DIST_GAMMA_RATE = 10,           /* shape, rate         gamma.hpp:113 */
  MB200_DIST_BETA = 11,                 /* a, b                  beta.hpp:12 */
  MB200_DIST_NEGATIVE_BINOMIAL_PROB = 12,/* size, prob negative_binomial.hpp:30 */
  MB200_DIST_NEGATIVE_BINOMIAL_MU = 13, /* size, mu  negative_binomial.hpp:43 */
  MB200_DIST_HYPERGEOMETRIC = 14,       /* n1, n2, k  hypergeometric.hpp:343 */
  MB200_DIST_TRUNCATED_NORMAL = 15,     /* mean, sd, min, max
                                                     truncated_normal.hpp:89 */
  MB200_DIST_CAUCHY = 16,               /* location, scale     cauchy.hpp:12 */
  MB200_DIST_WEIBULL = 17,              /* shape, scale       weibull.hpp:26 */
  MB200_DIST_LOG_NORMAL = 18,           /* meanlog, sdlog  log_normal.hpp:26 */
  MB200_DIST_BETA_BINOMIAL_PROB = 19,   /* size, prob, rho
                                                        beta_binomial.hpp:40 */
  MB200_DIST_BETA_BINOMIAL_AB = 20,     /* size, a, b   beta_binomial.hpp:27 */
  MB200_DIST_ZI_POISSON = 21,           /* pi0, lambda     zi_poisson.hpp:27 */
  MB200_DIST_ZI_NEGATIVE_BINOMIAL_PROB = 22, /* pi0, size, prob
                                                 zi_negative_binomial.hpp:32 */
  MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU = 23,   /* pi0, size, mu
                                                 zi_negative_binomial.hpp:49 */
  MB200_DIST_COUNT = 24
};

/* ---- density ids, one per density_<dist> of the reference
 *      (ref: src/density.cpp:6-254, inst/include/monty/random/density.hpp).
 *      "x:i32" marks the densities whose x (and listed size-like arguments)
 *      arrive as R integers (int32) in the reference. ---------------------- */
enum {
  MB200_DENS_BINOMIAL = 0,              /* x:i32 size:i32 prob               */
  MB200_DENS_NORMAL = 1,                /* x mu sd                           */
  MB200_DENS_NEGATIVE_BINOMIAL_MU = 2,  /* x:i32 size mu                     */
  MB200_DENS_NEGATIVE_BINOMIAL_PROB = 3,/* x:i32 size prob                   */
  MB200_DENS_BETA_BINOMIAL_PROB = 4,    /* x:i32 size:i32 prob rho           */
  MB200_DENS_BETA_BINOMIAL_AB = 5,      /* x:i32 size:i32 a b                */
  MB200_DENS_POISSON = 6,               /* x:i32 lambda                      */
  MB200_DENS_UNIFORM = 7,               /* x min max                         */
  MB200_DENS_EXPONENTIAL_RATE = 8,      /* x rate                            */
  MB200_DENS_EXPONENTIAL_MEAN = 9,      /* x mean                            */
  MB200_DENS_GAMMA_RATE = 10,           /* x shape rate                      */
  MB200_DENS_GAMMA_SCALE = 11,          /* x shape scale                     */
  MB200_DENS_BETA = 12,                 /* x a b                             */
  MB200_DENS_HYPERGEOMETRIC = 13,       /* x:i32 n1:i32 n2:i32 k:i32         */
  MB200_DENS_LOG_NORMAL = 14,           /* x mulog sdlog                     */
  MB200_DENS_TRUNCATED_NORMAL = 15,     /* x mu sd min max                   */
  MB200_DENS_CAUCHY = 16,               /* x location scale                  */
  MB200_DENS_WEIBULL = 17,              /* x shape scale                     */
  MB200_DENS_ZI_POISSON = 18,           /* x:i32 pi0 lambda                  */
  MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU = 19,  /* x:i32 pi0 size mu             */
  MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB = 20,/* x:i32 pi0 size prob           */
  MB200_DENS_COUNT = 21
};

#define MB200_MAX_PARAMS 4

typedef struct mb200_rng mb200_rng;   /* opaque: owns the device state array */

/* ---- library / device -------------------------------------------------- */

/* Version string of this library. */
const char* mb200_version(void);
/* Number of usable CUDA devices (0 if none; never falls back to the CPU). */
int mb200_device_count(void);
/* Thread-local message of the last failing call on this thread. */
const char* mb200_last_error(void);
/* Number of kernel launches issued by this library in this process so far
 * (used by bench.py to report gpu_launches). */
uint64_t mb200_launch_count(void);

/* ---- generator lifecycle ------------------------------------------------
 * ref: monty_rng_alloc, src/random.cpp:341-353; prng<T>::prng,
 *      prng.hpp:29-53; as_rng_seed, inst/include/monty/r/random.hpp:47-67.
 *
 * seed_state_host/seed_state_len: optional raw seed, k x 32 bytes holding k
 *   full xoshiro256 states ("raw vector" seed form); when seed_state_len is
 *   0 the integer `seed` is expanded with splitmix64 (generator.hpp:114-141).
 *   Streams beyond the k supplied states are produced by jump() from the
 *   previous stream, exactly as prng.hpp:45-51.  A length that is not a
 *   multiple of 32 fails with the reference's message
 *   "Expected raw vector of length as multiple of 32 for 'seed'".
 * stream_offset: the handle holds streams [stream_offset, stream_offset +
 *   n_streams) of the notional generator prng(stream_offset + n_streams,
 *   seed); used to shard one generator over several GPUs bit-exactly.
 * long_jumps: number of long_jump()s applied to the seed state before the
 *   jump ladder (monty_rng_distributed_state semantics, R/random.R:161-171).
 * device: CUDA device ordinal.
 */
int mb200_rng_create(uint64_t seed,
                     const uint8_t* seed_state_host, size_t seed_state_len,
                     size_t n_streams, int deterministic,
                     uint64_t stream_offset, uint64_t long_jumps,
                     int device, mb200_rng** out);
void mb200_rng_free(mb200_rng* rng);
size_t mb200_rng_n_streams(const mb200_rng* rng);
int mb200_rng_deterministic(const mb200_rng* rng);
int mb200_rng_device(const mb200_rng* rng);
const char* mb200_rng_last_error(const mb200_rng* rng);

/* ref: cpp_monty_rng_state, src/random.cpp:356-364. out_host receives
 * 32 * n_streams bytes. */
int mb200_rng_state(mb200_rng* rng, uint8_t* out_host, size_t len);
/* ref: cpp_monty_rng_set_state, src/random.cpp:366-379; len must be exactly
 * 32 * n_streams or the call fails with
 * "'value' must be a raw vector of length %d (but was %d)". */
int mb200_rng_set_state(mb200_rng* rng, const uint8_t* value_host, size_t len);
/* ref: cpp_monty_rng_jump / cpp_monty_rng_long_jump, src/random.cpp:382-396:
 * every stream jumps n times. */
int mb200_rng_jump(mb200_rng* rng, int n);
int mb200_rng_long_jump(mb200_rng* rng, int n);
/* Device pointer to the live state array (n_streams x 4 x uint64), for
 * device-side consumers (the dust2-style seam).  Owned by the handle. */
uint64_t* mb200_rng_state_dev(mb200_rng* rng);
/* The CUDA stream (cudaStream_t) all work of this handle is ordered on. */
void* mb200_rng_cuda_stream(mb200_rng* rng);
/* Block the calling thread until all queued work of the handle finished;
 * reports any deferred sampler error. */
int mb200_rng_sync(mb200_rng* rng);

/* ---- sampling -----------------------------------------------------------
 * ref: monty_random_sample_1_k / monty_random_sample_n_k,
 *      src/random.cpp:89-195 and 211-338.
 *
 * params_host[k] is parameter k of the distribution: a double vector of
 * length param_len[k] which must be 1 (broadcast) or n_streams, otherwise
 * the call fails with the reference's message
 * "Expected '<name>' to have length <n> or 1, not <len>" (random.cpp:21-42).
 * out_host receives n_samples * n_streams doubles laid out
 * out[n_samples * stream + sample] (random.cpp:220-224).
 * Invalid distribution parameters follow the reference's semantics: streams
 * before the first offending stream are advanced, the offender and later
 * streams are untouched, and the call fails with MB200_ERR_PARAM and the
 * reference's message text.
 * One documented divergence: a DATA-DEPENDENT failure inside a composition
 * (negative_binomial -> Poisson on the gamma draw, beta_binomial -> binomial on
 * the beta draw, multinomial's inner binomials: errors that depend on what was
 * drawn, not on the parameters) is detected after the launch.  The message is
 * the reference's, for the smallest failing stream, but by then every valid
 * stream of the call has advanced through all its draws, including streams
 * after the failing one, and the failing stream's output is NaN from the
 * failure on; the reference stops at the failing draw and leaves later streams
 * untouched (src/random.cpp:211-228).  These failures need parameters at the
 * edge of the double range (a gamma draw that overflows); the parity tests
 * cover the message text, not the state of the later streams.
 */
int mb200_rng_sample(mb200_rng* rng, int dist, int real_type, size_t n_samples,
                     const double* const* params_host, const size_t* param_len,
                     double* out_host);

/* Device-resident variant: parameters and output stay in HBM.
 * params_dev[k] points to doubles on the handle's device; param_len[k] is 1
 * or n_streams.  out_dev is double* when out_f32 == 0 and float* when
 * out_f32 != 0 (only allowed with MB200_REAL_F32).  Asynchronous on the
 * handle's stream; parameter errors are reported by the next
 * mb200_rng_sync().  Invalid streams are skipped (state untouched, output
 * NaN); all valid streams advance.
 * Ordering contract: the launch is ordered on the handle's own non-blocking
 * stream (mb200_rng_cuda_stream) and on nothing else.  A caller that produces
 * params_dev or consumes out_dev on another stream must order the two itself
 * (an event recorded on the producer stream and waited on by the handle's
 * stream before the call; the reverse after it) -- monty_b200/random.py's
 * sample_dev does exactly that with torch's current stream. */
int mb200_rng_sample_dev(mb200_rng* rng, int dist, int real_type,
                         size_t n_samples,
                         const double* const* params_dev,
                         const size_t* param_len,
                         void* out_dev, int out_f32);

/* ref: cpp_monty_random_multinomial / cpp_monty_random_n_multinomial,
 *      src/random.cpp:884-934; multinomial.hpp:53-82.
 * size_host: length size_len (1 or n_streams).  prob_host: prob_len x
 * prob_cols doubles, column-major, prob_cols is 1 or n_streams.
 * out_host: prob_len x n_samples x n_streams doubles. */
int mb200_rng_multinomial(mb200_rng* rng, size_t n_samples,
                          const double* size_host, size_t size_len,
                          const double* prob_host, size_t prob_len,
                          size_t prob_cols, double* out_host);

/* Raw generator output for golden-vector checks
 * (ref: test_xoshiro_run, src/test_rng.cpp:21-38; next(),
 * xoshiro256.hpp:65-103): n_samples successive next() values per stream,
 * out[n_samples * stream + sample], as uint64 (32-bit generators are
 * zero-extended).  `gen` selects the scrambler applied to the handle's
 * xoshiro256 state. */
int mb200_rng_raw(mb200_rng* rng, int gen, size_t n_samples, uint64_t* out_host);

/* Stand-alone generator check for all 12 generators of the reference
 * (ref: tests/testthat/xoshiro-ref/<name>, src/test_rng.cpp:21-38): seed ->
 * 3*n next() values on the device with a jump before value n and a long_jump
 * before value 2*n (1-based: i == n-1 and i == 2n-1 in the 0-based loop);
 * writes the values (zero-extended to uint64) and the final state words. */
int mb200_xoshiro_run(int gen, uint64_t seed, int n, int device,
                      uint64_t* out_values_host /* 3n */,
                      uint64_t* out_state_host /* <= 8 words */,
                      int* out_state_words);

/* ---- densities ----------------------------------------------------------
 * ref: src/density.cpp:6-254 (one loop per density), density.hpp.
 * x_host: int32 vector for the "x:i32" densities, double otherwise
 * (x_is_i32 says which was passed; double x is accepted for every density).
 * params_host[k]: parameter k; param_is_i32[k] != 0 means it is an int32
 * vector (the reference's cpp11::integers arguments), else double.  Every
 * argument is indexed by i (no recycling, as in the reference).
 * out_host (may be NULL) receives n doubles.  sum_out (may be NULL) receives
 * the sum of the n values (the log-likelihood when log != 0), reduced on the
 * device. */
int mb200_density(int dens, int real_type, size_t n,
                  const void* x_host, int x_is_i32,
                  const void* const* params_host, const int* param_is_i32,
                  int log, double* out_host, double* sum_out, int device);

/* Device-resident variant: all pointers are device pointers on `device`;
 * out_dev may be NULL (reduce-only); sum_dev (device double, may be NULL) is
 * overwritten with the sum.  `cuda_stream` may be NULL (default stream). */
int mb200_density_dev(int dens, int real_type, size_t n,
                      const void* x_dev, int x_is_i32,
                      const void* const* params_dev, const int* param_is_i32,
                      int log, double* out_dev, double* sum_dev,
                      int device, void* cuda_stream);

/* ---- named entry points, 1:1 with the reference's registered routines ----
 * (ref: src/cpp11.cpp:552-638).  Each is a thin wrapper over the generic
 * calls above; `n_samples == 0` selects the one-draw-per-stream form
 * (monty_random_<dist>), otherwise monty_random_n_<dist>. */
#define MB200_DECL_SAMPLER0(name) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, double* out_host);
#define MB200_DECL_SAMPLER1(name, a) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, \
                          const double* a, size_t a##_len, double* out_host);
#define MB200_DECL_SAMPLER2(name, a, b) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, \
                          const double* a, size_t a##_len, \
                          const double* b, size_t b##_len, double* out_host);
#define MB200_DECL_SAMPLER3(name, a, b, c) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, \
                          const double* a, size_t a##_len, \
                          const double* b, size_t b##_len, \
                          const double* c, size_t c##_len, double* out_host);
#define MB200_DECL_SAMPLER4(name, a, b, c, d) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, \
                          const double* a, size_t a##_len, \
                          const double* b, size_t b##_len, \
                          const double* c, size_t c##_len, \
                          const double* d, size_t d##_len, double* out_host);

MB200_DECL_SAMPLER0(real)                                   /* random.cpp:401,407 */
MB200_DECL_SAMPLER1(exponential_rate, rate)                 /* random.cpp:416,423 */
MB200_DECL_SAMPLER1(exponential_mean, mean)                 /* random.cpp:431,438 */
MB200_DECL_SAMPLER1(poisson, lambda)                        /* random.cpp:446,453 */
MB200_DECL_SAMPLER2(beta, a, b)                             /* random.cpp:461,469 */
MB200_DECL_SAMPLER2(binomial, size, prob)                   /* random.cpp:478,487 */
MB200_DECL_SAMPLER2(cauchy, location, scale)                /* random.cpp:498,507 */
MB200_DECL_SAMPLER2(gamma_scale, shape, scale)
MB200_DECL_SAMPLER2(gamma_rate, shape, rate)
MB200_DECL_SAMPLER2(negative_binomial_prob, size, prob)
MB200_DECL_SAMPLER2(negative_binomial_mu, size, mu)
MB200_DECL_SAMPLER2(normal_box_muller, mean, sd)
MB200_DECL_SAMPLER2(normal_polar, mean, sd)
MB200_DECL_SAMPLER2(normal_ziggurat, mean, sd)
MB200_DECL_SAMPLER2(uniform, min, max)
MB200_DECL_SAMPLER2(weibull, shape, scale)
MB200_DECL_SAMPLER2(log_normal, meanlog, sdlog)
MB200_DECL_SAMPLER2(zi_poisson, pi0, lambda)
MB200_DECL_SAMPLER3(beta_binomial_prob, size, prob, rho)
MB200_DECL_SAMPLER3(beta_binomial_ab, size, a, b)
MB200_DECL_SAMPLER3(hypergeometric, n1, n2, k)
MB200_DECL_SAMPLER3(zi_negative_binomial_prob, pi0, size, prob)
MB200_DECL_SAMPLER3(zi_negative_binomial_mu, pi0, size, mu)
MB200_DECL_SAMPLER4(truncated_normal, mean, sd, min, max)   /* random.cpp:858,871 */

#ifdef __cplusplus
}
#endif
#endif /* MONTY_B200_H */
