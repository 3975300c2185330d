// rng_kernels.cu -- generator-level kernels: jump / long_jump of every
// stream, the doubling ladder that seeds n streams in O(log n) launches, raw
// next() output, and the 12-generator golden protocol.
#include "../../include/monty_b200.h"
#include "rng_dispatch.h"
#include "glibc_math.cuh"
#include "xoshiro.cuh"

namespace mb200 {

struct PolyArg { uint64_t c[4]; };

// state[first + i] <- poly(T) state[first + i]   (every stream jumps; ref:
// src/random.cpp:382-396 with the n-fold loop folded into one polynomial)
__global__ void __launch_bounds__(128)
apply_poly_kernel(uint64_t* __restrict__ state, unsigned long long first,
                  unsigned long long count, const PolyArg poly) {
  const unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  X256 s = load_state(state, first + i);
  x256_apply_poly(s, poly.c);
  store_state(state, first + i, s);
}

// One rung of the doubling ladder.  With streams [anchor, anchor + have)
// already in place and poly = c^have, stream anchor + have + r is `have`
// jumps ahead of stream anchor + r:
//     state[anchor + have + r] <- poly(T) state[anchor + r],  r < n_new <= have
// Every stream is produced by exactly one jump evaluation, so seeding n
// streams costs n jump evaluations in ceil(log2 n) launches, instead of the
// reference's n-1 sequential jumps (prng.hpp:45-51).
__global__ void __launch_bounds__(128)
ladder_kernel(uint64_t* __restrict__ state, unsigned long long anchor,
              unsigned long long have, unsigned long long n_new, const PolyArg poly) {
  const unsigned long long r = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_new) return;
  X256 s = load_state(state, anchor + r);
  x256_apply_poly(s, poly.c);
  store_state(state, anchor + have + r, s);
}

// n_samples successive next() per stream with scrambler SCR; out is
// stream-major like every sampler.  A diagnostic path (golden vectors), not
// tuned: each lane writes its own row.
template <int SCR>
__global__ void __launch_bounds__(128)
raw_kernel(uint64_t* __restrict__ state, unsigned long long n_streams,
           unsigned long long n_samples, uint64_t* __restrict__ out) {
  const unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_streams) return;
  X256 s = load_state(state, i);
  for (unsigned long long j = 0; j < n_samples; ++j) out[i * n_samples + j] = x256_next<SCR>(s);
  store_state(state, i, s);
}

// ---- all 12 generators of the reference, for the golden protocol ------------
// ref: hdr/xoshiro128.hpp:63-100, hdr/xoroshiro128.hpp:68-102,
//      hdr/xoshiro256.hpp:65-103, hdr/xoshiro512.hpp:74-123
struct GenDesc { int family, scr, n_words, bits; uint64_t jump[8], long_jump[8]; };

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

__device__ uint64_t gen_next(const GenDesc& g, uint64_t* s) {
  uint64_t result = 0;
  if (g.family == 0) {
    X256 st{s[0], s[1], s[2], s[3]};
    result = g.scr == kPlus ? x256_next<kPlus>(st)
           : g.scr == kPlusPlus ? x256_next<kPlusPlus>(st) : x256_next<kStarStar>(st);
    s[0] = st.s0; s[1] = st.s1; s[2] = st.s2; s[3] = st.s3;
  } else if (g.family == 1) {
    uint32_t a = s[0], b = s[1], c = s[2], d = s[3];
    uint32_t r;
    if (g.scr == kPlus) r = a + d;
    else if (g.scr == kPlusPlus) r = rotl32(a + d, 7) + a;
    else r = rotl32(b * 5, 7) * 9;
    const uint32_t t = b << 9;
    c ^= a; d ^= b; b ^= c; a ^= d;
    c ^= t;
    d = rotl32(d, 11);
    s[0] = a; s[1] = b; s[2] = c; s[3] = d;
    result = r;
  } else if (g.family == 2) {
    const uint64_t s0 = s[0];
    uint64_t s1 = s[1];
    if (g.scr == kPlus) result = s0 + s1;
    else if (g.scr == kPlusPlus) result = rotl64(s0 + s1, 17) + s0;
    else result = rotl64(s0 * 5, 7) * 9;
    s1 ^= s0;
    if (g.scr == kPlusPlus) {
      s[0] = rotl64(s0, 49) ^ s1 ^ (s1 << 21);
      s[1] = rotl64(s1, 28);
    } else {
      s[0] = rotl64(s0, 24) ^ s1 ^ (s1 << 16);
      s[1] = rotl64(s1, 37);
    }
  } else {
    if (g.scr == kPlus) result = s[0] + s[2];
    else if (g.scr == kPlusPlus) result = rotl64(s[0] + s[2], 17) + s[2];
    else result = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 11;
    s[2] ^= s[0]; s[5] ^= s[1]; s[1] ^= s[2]; s[7] ^= s[3];
    s[3] ^= s[4]; s[4] ^= s[5]; s[0] ^= s[6]; s[6] ^= s[7];
    s[6] ^= t;
    s[7] = rotl64(s[7], 21);
  }
  return result;
}

// hdr/generator.hpp:86-106
__device__ void gen_jump(const GenDesc& g, uint64_t* s, const uint64_t* coef) {
  uint64_t work[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < g.n_words; ++i) {
    for (int b = 0; b < g.bits; ++b) {
      if ((coef[i] >> b) & 1ull) {
        for (int j = 0; j < g.n_words; ++j) work[j] ^= s[j];
      }
      gen_next(g, s);
    }
  }
  for (int j = 0; j < g.n_words; ++j) s[j] = work[j];
}

// src/test_rng.cpp:21-38 on the device: seed (splitmix64, generator.hpp:114-122)
// -> 3n values with a jump before index n-1 and a long_jump before 2n-1.
__global__ void xoshiro_run_kernel(const GenDesc g, uint64_t seed, int n,
                                   uint64_t* __restrict__ values, uint64_t* __restrict__ state_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint64_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < g.n_words; ++i) {
    uint64_t z = (seed += 0x9e3779b97f4a7c15ull);      // hdr/utils.hpp:72-77
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    seed = z ^ (z >> 31);
    s[i] = g.bits == 32 ? (seed & 0xffffffffull) : seed;
  }
  for (int i = 0; i < 3 * n; ++i) {
    if (i == n - 1) gen_jump(g, s, g.jump);
    else if (i == 2 * n - 1) gen_jump(g, s, g.long_jump);
    values[i] = gen_next(g, s);
  }
  for (int k = 0; k < g.n_words; ++k) state_out[k] = s[k];
}

static const GenDesc k_gens[MB200_GEN_COUNT] = {
#define X256D(scr) {0, scr, 4, 64, \
  {0x180ec6d33cfd0abaull, 0xd5a61266f0c9392cull, 0xa9582618e03fc9aaull, 0x39abdc4529b1661cull}, \
  {0x76e15d3efefdcbbfull, 0xc5004e441c522fb3ull, 0x77710069854ee241ull, 0x39109bb02acbe635ull}}
  X256D(kPlus), X256D(kPlusPlus), X256D(kStarStar),
#define X128D(scr) {1, scr, 4, 32, \
  {0x8764000bull, 0xf542d2d3ull, 0x6fa035c3ull, 0x77f2db5bull}, \
  {0xb523952eull, 0x0b6f099full, 0xccf5a0efull, 0x1c580662ull}}
  X128D(kPlus), X128D(kPlusPlus), X128D(kStarStar),
  {2, kPlus, 2, 64, {0xdf900294d8f554a5ull, 0x170865df4b3201fcull},
                    {0xd2a98b26625eee7bull, 0xdddf9b1090aa7ac1ull}},
  {2, kPlusPlus, 2, 64, {0x2bd7a6a6e99c2ddcull, 0x0992ccaf6a6fca05ull},
                        {0x360fd5f2cf8d5d99ull, 0x9c6e6877736c46e3ull}},
  {2, kStarStar, 2, 64, {0xdf900294d8f554a5ull, 0x170865df4b3201fcull},
                        {0xd2a98b26625eee7bull, 0xdddf9b1090aa7ac1ull}},
#define X512D(scr) {3, scr, 8, 64, \
  {0x33ed89b6e7a353f9ull, 0x760083d7955323beull, 0x2837f2fbb5f22faeull, 0x4b8c5674d309511cull, \
   0xb11ac47a7ba28c25ull, 0xf1be7667092bcc1cull, 0x53851efdb6df0aafull, 0x1ebbc8b23eaf25dbull}, \
  {0x11467fef8f921d28ull, 0xa2a819f2e79c8ea8ull, 0xa8299fc284b3959aull, 0xb4d347340ca63ee1ull, \
   0x1cb0940bedbff6ceull, 0xd956c5c4fa1f8e17ull, 0x915e38fd4eda93bcull, 0x5b3ccdfa5d7daca5ull}}
  X512D(kPlus), X512D(kPlusPlus), X512D(kStarStar),
};

// ---- host launchers ----------------------------------------------------------
static inline unsigned blocks_for(unsigned long long n, int block) {
  return static_cast<unsigned>((n + block - 1) / block);
}

cudaError_t launch_apply_poly(uint64_t* state, unsigned long long first, unsigned long long count,
                              const uint64_t poly[4], cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  PolyArg p{{poly[0], poly[1], poly[2], poly[3]}};
  apply_poly_kernel<<<blocks_for(count, 128), 128, 0, st>>>(state, first, count, p);
  return cudaGetLastError();
}

cudaError_t launch_ladder(uint64_t* state, unsigned long long anchor, unsigned long long have,
                          unsigned long long n_new, const uint64_t poly[4], cudaStream_t st) {
  if (n_new == 0) return cudaSuccess;
  PolyArg p{{poly[0], poly[1], poly[2], poly[3]}};
  ladder_kernel<<<blocks_for(n_new, 128), 128, 0, st>>>(state, anchor, have, n_new, p);
  return cudaGetLastError();
}

cudaError_t launch_raw(int gen, uint64_t* state, unsigned long long n_streams,
                       unsigned long long n_samples, uint64_t* out, cudaStream_t st) {
  if (n_streams == 0) return cudaSuccess;
  const unsigned g = blocks_for(n_streams, 128);
  switch (gen) {
  case MB200_GEN_XOSHIRO256PLUS: raw_kernel<kPlus><<<g, 128, 0, st>>>(state, n_streams, n_samples, out); break;
  case MB200_GEN_XOSHIRO256PLUSPLUS: raw_kernel<kPlusPlus><<<g, 128, 0, st>>>(state, n_streams, n_samples, out); break;
  case MB200_GEN_XOSHIRO256STARSTAR: raw_kernel<kStarStar><<<g, 128, 0, st>>>(state, n_streams, n_samples, out); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_xoshiro_run(int gen, uint64_t seed, int n, uint64_t* values_dev,
                               uint64_t* state_dev, int* n_words, cudaStream_t st) {
  if (gen < 0 || gen >= MB200_GEN_COUNT) return cudaErrorInvalidValue;
  *n_words = k_gens[gen].n_words;
  xoshiro_run_kernel<<<1, 32, 0, st>>>(k_gens[gen], seed, n, values_dev, state_dev);
  return cudaGetLastError();
}

// The lgamma-of-small-integers table behind m_lgamma (samplers.cuh): entry i is
// what m_lgamma(i) computes without the table -- glibc's lgamma / lgammaf bit for
// bit (glibc_math.cuh) -- so a table hit is
// bit-identical to the call it replaces.
__global__ void lgt_build_kernel(double* __restrict__ d, float* __restrict__ f, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    d[i] = i >= 1 ? glibc::lgamma_positive(static_cast<double>(i)) : lgamma(static_cast<double>(i));
    f[i] = i >= 1 ? glibc::lgammaf_positive(static_cast<float>(i)) : lgammaf(static_cast<float>(i));
  }
}
cudaError_t launch_lgamma_table_build(double* d, float* f, int n, cudaStream_t st) {
  lgt_build_kernel<<<(n + 255) / 256, 256, 0, st>>>(d, f, n);
  return cudaGetLastError();
}

}  // namespace mb200
