/* oracle/monty_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the reference algorithms on the hot path of
 * mrc-ide/monty: the xoshiro/xoroshiro generators with splitmix64 seeding and
 * jump / long_jump, the multi-stream container semantics, every per-stream
 * sampler behind monty_random_* (double and float real_type) and the
 * densities of density.hpp.  It is the parity checker for the CUDA product in
 * monty_b200/ and is NEVER on the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.
 *
 * Pinning.  This restatement is itself pinned against
 *   (1) the 12 golden files of the reference (tests/golden/xoshiro-ref/,
 *       copied from tests/testthat/xoshiro-ref), and
 *   (2) the unmodified reference headers compiled as oracle/_ref/
 *       libmonty_ref.so (bit-for-bit: raw output, states, every sampler in
 *       double and float, every density),
 * by tests/test_oracle.py.  Where the arithmetic goes through libm the
 * reference itself is pinned only by this image's glibc 2.39.
 *
 * Citations are to the reference tree; "hdr/" = inst/include/monty/random/.
 * Build: see oracle/Makefile (-O2 -ffp-contract=off, no fast-math).
 */
#define _GNU_SOURCE
#include <limits.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/monty_b200.h"
#include "oracle_tables.inc"

static const double orc_zig_x[257] = { MONTY_ZIG_X_VALUES };
static const double orc_zig_y[256] = { MONTY_ZIG_Y_VALUES };
static const float orc_tail_f[128] = { MONTY_TAIL_F_VALUES };
static const double orc_tail_d[10] = { MONTY_TAIL_D_VALUES };

/* ---- error reporting (the reference throws; hdr/cuda_compatibility.hpp:49-57) */
static _Thread_local int orc_failed;
static _Thread_local char orc_msg[256];
static char orc_last[256];
static void orc_fail(const char* fmt, ...) {
  if (orc_failed) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(orc_msg, sizeof orc_msg, fmt, ap);
  va_end(ap);
  orc_failed = 1;
}
#define ORC_FAIL(...) orc_fail(__VA_ARGS__)
const char* orc_last_error(void) { return orc_last; }

/* ---- generators ------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); } /* hdr/utils.hpp:46-49 */
static inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); } /* hdr/utils.hpp:51-54 */

/* hdr/utils.hpp:72-77 */
static inline uint64_t splitmix64(uint64_t seed) {
  uint64_t z = (seed += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

/* One stream of the sampler generator, xoshiro256+ (hdr/xoshiro_state.hpp:20-39). */
typedef struct { uint64_t s[4]; int deterministic; } x256;

enum { SCR_PLUS, SCR_PLUSPLUS, SCR_STARSTAR };

/* A generic generator state wide enough for all 12 variants.  Words are held
 * as uint64 (32-bit generators use the low halves). */
typedef struct { uint64_t w[8]; } gstate;
typedef struct {
  int family;   /* 0: xoshiro256, 1: xoshiro128, 2: xoroshiro128, 3: xoshiro512 */
  int scr;
  int n_words;
  int bits;     /* width of a word */
  uint64_t jump[8], long_jump[8];
} gen_desc;

/* jump / long-jump polynomials: hdr/xoshiro256.hpp:25-63, hdr/xoshiro128.hpp,
 * hdr/xoroshiro128.hpp, hdr/xoshiro512.hpp (same tables as Vigna's public
 * domain originals at prng.di.unimi.it) */
static const gen_desc k_gens[MB200_GEN_COUNT] = {
#define X256(scr) {0, scr, 4, 64, \
  {0x180ec6d33cfd0abaull, 0xd5a61266f0c9392cull, 0xa9582618e03fc9aaull, 0x39abdc4529b1661cull}, \
  {0x76e15d3efefdcbbfull, 0xc5004e441c522fb3ull, 0x77710069854ee241ull, 0x39109bb02acbe635ull}}
  X256(SCR_PLUS), X256(SCR_PLUSPLUS), X256(SCR_STARSTAR),
#define X128(scr) {1, scr, 4, 32, \
  {0x8764000bu, 0xf542d2d3u, 0x6fa035c3u, 0x77f2db5bu}, \
  {0xb523952eu, 0x0b6f099fu, 0xccf5a0efu, 0x1c580662u}}
  X128(SCR_PLUS), X128(SCR_PLUSPLUS), X128(SCR_STARSTAR),
  /* xoroshiro128: + and ** share constants, ++ has its own */
  {2, SCR_PLUS, 2, 64, {0xdf900294d8f554a5ull, 0x170865df4b3201fcull},
                       {0xd2a98b26625eee7bull, 0xdddf9b1090aa7ac1ull}},
  {2, SCR_PLUSPLUS, 2, 64, {0x2bd7a6a6e99c2ddcull, 0x0992ccaf6a6fca05ull},
                           {0x360fd5f2cf8d5d99ull, 0x9c6e6877736c46e3ull}},
  {2, SCR_STARSTAR, 2, 64, {0xdf900294d8f554a5ull, 0x170865df4b3201fcull},
                           {0xd2a98b26625eee7bull, 0xdddf9b1090aa7ac1ull}},
#define X512(scr) {3, scr, 8, 64, \
  {0x33ed89b6e7a353f9ull, 0x760083d7955323beull, 0x2837f2fbb5f22faeull, 0x4b8c5674d309511cull, \
   0xb11ac47a7ba28c25ull, 0xf1be7667092bcc1cull, 0x53851efdb6df0aafull, 0x1ebbc8b23eaf25dbull}, \
  {0x11467fef8f921d28ull, 0xa2a819f2e79c8ea8ull, 0xa8299fc284b3959aull, 0xb4d347340ca63ee1ull, \
   0x1cb0940bedbff6ceull, 0xd956c5c4fa1f8e17ull, 0x915e38fd4eda93bcull, 0x5b3ccdfa5d7daca5ull}}
  X512(SCR_PLUS), X512(SCR_PLUSPLUS), X512(SCR_STARSTAR),
};

/* next() for each family: hdr/xoshiro256.hpp:65-103, hdr/xoshiro128.hpp:63-100,
 * hdr/xoroshiro128.hpp:68-102, hdr/xoshiro512.hpp:74-123 */
static uint64_t gen_next(const gen_desc* g, gstate* st) {
  uint64_t* s = st->w;
  uint64_t result = 0;
  switch (g->family) {
  case 0: {
    if (g->scr == SCR_PLUS) result = s[0] + s[3];
    else if (g->scr == SCR_PLUSPLUS) result = rotl64(s[0] + s[3], 23) + s[0];
    else result = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
  }
  case 1: {
    uint32_t a = (uint32_t)s[0], b = (uint32_t)s[1], c = (uint32_t)s[2], d = (uint32_t)s[3];
    uint32_t r;
    if (g->scr == SCR_PLUS) r = a + d;
    else if (g->scr == SCR_PLUSPLUS) r = rotl32(a + d, 7) + a;
    else r = rotl32(b * 5, 7) * 9;
    const uint32_t t = b << 9;
    c ^= a; d ^= b; b ^= c; a ^= d;
    c ^= t;
    d = rotl32(d, 11);
    s[0] = a; s[1] = b; s[2] = c; s[3] = d;
    return r;
  }
  case 2: {
    const uint64_t s0 = s[0];
    uint64_t s1 = s[1];
    if (g->scr == SCR_PLUS) result = s0 + s1;
    else if (g->scr == SCR_PLUSPLUS) result = rotl64(s0 + s1, 17) + s0;
    else result = rotl64(s0 * 5, 7) * 9;
    s1 ^= s0;
    if (g->scr == SCR_PLUSPLUS) {
      s[0] = rotl64(s0, 49) ^ s1 ^ (s1 << 21);
      s[1] = rotl64(s1, 28);
    } else {
      s[0] = rotl64(s0, 24) ^ s1 ^ (s1 << 16);
      s[1] = rotl64(s1, 37);
    }
    return result;
  }
  default: {
    if (g->scr == SCR_PLUS) result = s[0] + s[2];
    else if (g->scr == SCR_PLUSPLUS) result = rotl64(s[0] + s[2], 17) + s[2];
    else result = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 11;
    s[2] ^= s[0]; s[5] ^= s[1]; s[1] ^= s[2]; s[7] ^= s[3];
    s[3] ^= s[4]; s[4] ^= s[5]; s[0] ^= s[6]; s[6] ^= s[7];
    s[6] ^= t;
    s[7] = rotl64(s[7], 21);
    return result;
  }
  }
}

/* hdr/generator.hpp:86-106: for each coefficient bit, low word / low bit
 * first: accumulate the state if the bit is set, then step. */
static void gen_jump(const gen_desc* g, gstate* st, const uint64_t* coef) {
  uint64_t work[8] = {0};
  for (int i = 0; i < g->n_words; ++i) {
    for (int b = 0; b < g->bits; ++b) {
      if (coef[i] & ((uint64_t)1 << b)) {
        for (int j = 0; j < g->n_words; ++j) work[j] ^= st->w[j];
      }
      gen_next(g, st);
    }
  }
  for (int j = 0; j < g->n_words; ++j) st->w[j] = work[j];
}

/* hdr/generator.hpp:114-122: state[i] = (int_type)(seed = splitmix64(seed)) */
static void gen_seed(const gen_desc* g, gstate* st, uint64_t seed) {
  memset(st, 0, sizeof *st);
  for (int i = 0; i < g->n_words; ++i) {
    seed = splitmix64(seed);
    st->w[i] = g->bits == 32 ? (uint32_t)seed : seed;
  }
}

/* the sampler generator, specialised: xoshiro256+ (hdr/xoshiro256.hpp:91-103) */
static inline uint64_t x256p_next(x256* st) {
  uint64_t* s = st->s;
  const uint64_t result = s[0] + s[3];
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl64(s[3], 45);
  return result;
}

static void x256_jump(uint64_t* s4, int is_long) {
  gstate g;
  memset(&g, 0, sizeof g);
  memcpy(g.w, s4, 32);
  gen_jump(&k_gens[0], &g, is_long ? k_gens[0].long_jump : k_gens[0].jump);
  memcpy(s4, g.w, 32);
}

/* ---- samplers and densities, generic over real_type ------------------------ */
#define ORC_CAT_(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT_(a, b)

#define R double
#define ORC_IS_FLOAT 0
#define FN(name) ORC_CAT(name, _d)
#define M(fn) fn
#define R_EPSILON DBL_EPSILON
#include "oracle_samplers.inc"
#undef R
#undef ORC_IS_FLOAT
#undef FN
#undef M
#undef R_EPSILON

#define R float
#define ORC_IS_FLOAT 1
#define FN(name) ORC_CAT(name, _f)
#define M(fn) ORC_CAT(fn, f)
#define R_EPSILON FLT_EPSILON
#include "oracle_samplers.inc"
#undef R
#undef ORC_IS_FLOAT
#undef FN
#undef M
#undef R_EPSILON

/* ---- multi-stream container and boundary loops ----------------------------- */
static const int k_n_params[MB200_DIST_COUNT] = {
  0, 2, 1, 1, 2, 2, 2, 2, 1, 2, 2, 2, 2, 2, 3, 4, 2, 2, 2, 3, 3, 2, 3, 3};

int orc_n_params(int dist) {
  return (dist >= 0 && dist < MB200_DIST_COUNT) ? k_n_params[dist] : -1;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* prng<T>(n, seed words): hdr/prng.hpp:38-53.  Copy full states while the
 * seed vector supplies them, then each further stream is jump() of the
 * previous one.  n_words = 0 means "integer seed": seed_data(seed),
 * hdr/generator.hpp:134-141. */
int orc_seed_states(uint64_t seed, const uint64_t* words, size_t n_words,
                    size_t n_streams, uint64_t* out) {
  uint64_t cur[4];
  uint64_t from_int[4];
  if (n_words == 0) {
    gstate g;
    gen_seed(&k_gens[0], &g, seed);
    memcpy(from_int, g.w, 32);
    words = from_int;
    n_words = 4;
  }
  const size_t n_seed = n_words / 4;
  memset(cur, 0, sizeof cur);  /* prng.hpp:41: `rng_state s;` -- only reachable with n_seed == 0 */
  for (size_t i = 0; i < n_streams; ++i) {
    if (i < n_seed) memcpy(cur, words + 4 * i, 32);
    else x256_jump(cur, 0);
    memcpy(out + 4 * i, cur, 32);
  }
  return 0;
}

/* src/random.cpp:382-396: every stream jumps n times */
int orc_jump(uint64_t* state, size_t n_streams, int n, int is_long) {
  for (int k = 0; k < n; ++k)
    for (size_t i = 0; i < n_streams; ++i) x256_jump(state + 4 * i, is_long);
  return 0;
}

/* src/random.cpp:211-338 (and :89-195 with n_samples = 1): stream-outer,
 * sample-inner; parameter k is broadcast when param_len[k] == 1
 * (src/random.cpp:21-42).  On a parameter error the loop stops: earlier
 * streams stay advanced, exactly as the reference's exception does. */
int orc_sample(int dist, int real_type, uint64_t* state, size_t n_streams,
               int deterministic, size_t n_samples,
               const double* const* params, const size_t* param_len,
               double* out, int n_threads) {
  const int np = orc_n_params(dist);
  if (np < 0) { snprintf(orc_last, sizeof orc_last, "unknown distribution id"); return -2; }
  const double* p[MB200_MAX_PARAMS] = {0, 0, 0, 0};
  size_t stride[MB200_MAX_PARAMS] = {0, 0, 0, 0};
  for (int k = 0; k < np; ++k) {
    if (param_len[k] != 1 && param_len[k] != n_streams) {
      snprintf(orc_last, sizeof orc_last, "bad parameter length"); return -2;
    }
    p[k] = params[k];
    stride[k] = param_len[k] == 1 ? 0 : 1;
  }
  int failed = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 1 ? n_threads : 1) if (n_threads > 1)
#endif
  for (long long ii = 0; ii < (long long)n_streams; ++ii) {
    const size_t i = (size_t)ii;
    if (failed) continue;
    x256 s;
    memcpy(s.s, state + 4 * i, 32);
    s.deterministic = deterministic;
    double a[MB200_MAX_PARAMS] = {0, 0, 0, 0};
    for (int k = 0; k < np; ++k) a[k] = p[k][stride[k] * i];
    orc_failed = 0;
    for (size_t j = 0; j < n_samples; ++j) {
      double y;
      if (real_type == MB200_REAL_F32) {
        y = draw_f(dist, &s, (float)a[0], (float)a[1], (float)a[2], (float)a[3]);
      } else {
        y = draw_d(dist, &s, a[0], a[1], a[2], a[3]);
      }
      if (orc_failed) break;
      out[n_samples * i + j] = y;
    }
    memcpy(state + 4 * i, s.s, 32);
    if (orc_failed) {
#ifdef _OPENMP
#pragma omp critical
#endif
      { if (!failed) { failed = 1; memcpy(orc_last, orc_msg, sizeof orc_last); } }
      orc_failed = 0;
    }
  }
  return failed ? -3 : 0;
}

/* src/random.cpp:884-934 */
int orc_multinomial(uint64_t* state, size_t n_streams, int deterministic,
                    size_t n_samples,
                    const double* size, size_t size_len,
                    const double* prob, size_t prob_len, size_t prob_cols,
                    double* out) {
  for (size_t i = 0; i < n_streams; ++i) {
    x256 s;
    memcpy(s.s, state + 4 * i, 32);
    s.deterministic = deterministic;
    const double* pr = prob + (prob_cols == 1 ? 0 : i * prob_len);
    const double sz = size[size_len == 1 ? 0 : i];
    orc_failed = 0;
    for (size_t j = 0; j < n_samples && !orc_failed; ++j) {
      multinomial_d(&s, sz, pr, (int)prob_len, out + prob_len * (n_samples * i + j));
    }
    memcpy(state + 4 * i, s.s, 32);
    if (orc_failed) { memcpy(orc_last, orc_msg, sizeof orc_last); orc_failed = 0; return -3; }
  }
  return 0;
}

/* next() with a chosen scrambler on 4 x u64 states (hdr/xoshiro256.hpp:65-103) */
int orc_raw(int gen, uint64_t* state, size_t n_streams, size_t n_samples, uint64_t* out) {
  if (gen < 0 || gen > MB200_GEN_XOSHIRO256STARSTAR) {
    snprintf(orc_last, sizeof orc_last, "raw: generator must be a xoshiro256 variant");
    return -2;
  }
  for (size_t i = 0; i < n_streams; ++i) {
    gstate g;
    memset(&g, 0, sizeof g);
    memcpy(g.w, state + 4 * i, 32);
    for (size_t j = 0; j < n_samples; ++j) out[n_samples * i + j] = gen_next(&k_gens[gen], &g);
    memcpy(state + 4 * i, g.w, 32);
  }
  return 0;
}

/* the golden protocol, src/test_rng.cpp:21-38: seed -> 3n values, a jump
 * before value index n-1 and a long_jump before index 2n-1. */
int orc_xoshiro_run(int gen, uint64_t seed, int n, uint64_t* values,
                    uint64_t* state_out, int* n_words) {
  if (gen < 0 || gen >= MB200_GEN_COUNT) {
    snprintf(orc_last, sizeof orc_last, "unknown generator id"); return -2;
  }
  const gen_desc* g = &k_gens[gen];
  gstate st;
  gen_seed(g, &st, seed);
  for (int i = 0; i < 3 * n; ++i) {
    if (i == n - 1) gen_jump(g, &st, g->jump);
    else if (i == 2 * n - 1) gen_jump(g, &st, g->long_jump);
    values[i] = gen_next(g, &st);
  }
  *n_words = g->n_words;
  for (int k = 0; k < g->n_words; ++k) state_out[k] = st.w[k];
  return 0;
}

/* src/density.cpp:6-254: every argument indexed by i, int32 or double */
static inline double arg_at(const void* p, int is_i32, size_t i) {
  return is_i32 ? (double)((const int32_t*)p)[i] : ((const double*)p)[i];
}
int orc_density(int dens, int real_type, size_t n, const void* x, int x_is_i32,
                const void* const* params, const int* param_is_i32, int lg,
                double* out, int n_threads) {
  if (dens < 0 || dens >= MB200_DENS_COUNT) {
    snprintf(orc_last, sizeof orc_last, "unknown density id"); return -2;
  }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 1 ? n_threads : 1) if (n_threads > 1)
#endif
  for (long long ii = 0; ii < (long long)n; ++ii) {
    const size_t i = (size_t)ii;
    double a[MB200_MAX_PARAMS] = {0, 0, 0, 0};
    for (int k = 0; k < MB200_MAX_PARAMS; ++k) {
      if (params && params[k]) a[k] = arg_at(params[k], param_is_i32 ? param_is_i32[k] : 0, i);
    }
    const double xi = arg_at(x, x_is_i32, i);
    if (real_type == MB200_REAL_F32) {
      out[i] = density_f(dens, (float)xi, (float)a[0], (float)a[1], (float)a[2], (float)a[3], lg);
    } else {
      out[i] = density_d(dens, xi, a[0], a[1], a[2], a[3], lg);
    }
  }
  return 0;
}
