// generators.cuh -- the reference's other nine generators (xoshiro128, xoroshiro128,
// xoshiro512, each with the +, ++ and ** scramblers) as register-resident states with the
// interface the samplers of samplers.cuh are templated on:
//
//     struct R { state_type s; int_type next(); }
//
// ref: inst/include/monty/random/xoshiro128.hpp:63-100, xoroshiro128.hpp:68-102,
//      xoshiro512.hpp:74-123 (next), :25-60 etc. (jump constants),
//      generator.hpp:86-106 (rng_jump_state), random.hpp:31-47 (generator<float> is
//      xoshiro128+).
// The batch kernels behind the C-ABI use xoshiro256+ only (src/random.cpp:10); these types
// serve the device-side API (include/monty_b200/random.cuh), where a model compiled for
// real_type = float gets xoshiro128+ exactly as with the reference.  Plain C++: usable on
// the host too (seeding and jumping a prng before upload).
#pragma once
#include <cstdint>

#include "xoshiro.cuh"

namespace mb200 {

struct X128 { uint32_t s0, s1, s2, s3; };       // xoshiro128: 4 x u32
struct XR128 { uint64_t s0, s1; };              // xoroshiro128: 2 x u64
struct X512 { uint64_t s[8]; };                 // xoshiro512: 8 x u64

__host__ __device__ __forceinline__ uint32_t g_rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
__host__ __device__ __forceinline__ uint64_t g_rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// hdr/xoshiro256.hpp:65-103 in plain C++ (the device's batch kernels use x256_next, which
// is written on 32-bit halves; same function)
template <int SCR>
__host__ __device__ __forceinline__ uint64_t gen_next(X256& s) {
  const uint64_t result = SCR == kPlus ? s.s0 + s.s3
                        : SCR == kPlusPlus ? g_rotl64(s.s0 + s.s3, 23) + s.s0 : g_rotl64(s.s1 * 5, 7) * 9;
  const uint64_t t = s.s1 << 17;
  s.s2 ^= s.s0; s.s3 ^= s.s1; s.s1 ^= s.s2; s.s0 ^= s.s3;
  s.s2 ^= t;
  s.s3 = g_rotl64(s.s3, 45);
  return result;
}
// hdr/xoshiro128.hpp:63-100
template <int SCR>
__host__ __device__ __forceinline__ uint32_t gen_next(X128& s) {
  const uint32_t result = SCR == kPlus ? s.s0 + s.s3
                        : SCR == kPlusPlus ? g_rotl32(s.s0 + s.s3, 7) + s.s0 : g_rotl32(s.s1 * 5, 7) * 9;
  const uint32_t t = s.s1 << 9;
  s.s2 ^= s.s0; s.s3 ^= s.s1; s.s1 ^= s.s2; s.s0 ^= s.s3;
  s.s2 ^= t;
  s.s3 = g_rotl32(s.s3, 11);
  return result;
}
// hdr/xoroshiro128.hpp:68-102 (the ++ variant has its own shifts)
template <int SCR>
__host__ __device__ __forceinline__ uint64_t gen_next(XR128& s) {
  const uint64_t s0 = s.s0;
  uint64_t s1 = s.s1;
  const uint64_t result = SCR == kPlus ? s0 + s1
                        : SCR == kPlusPlus ? g_rotl64(s0 + s1, 17) + s0 : g_rotl64(s0 * 5, 7) * 9;
  s1 ^= s0;
  if (SCR == kPlusPlus) {
    s.s0 = g_rotl64(s0, 49) ^ s1 ^ (s1 << 21);
    s.s1 = g_rotl64(s1, 28);
  } else {
    s.s0 = g_rotl64(s0, 24) ^ s1 ^ (s1 << 16);
    s.s1 = g_rotl64(s1, 37);
  }
  return result;
}
// hdr/xoshiro512.hpp:74-123
template <int SCR>
__host__ __device__ __forceinline__ uint64_t gen_next(X512& g) {
  uint64_t* s = g.s;
  const uint64_t result = SCR == kPlus ? s[0] + s[2]
                        : SCR == kPlusPlus ? g_rotl64(s[0] + s[2], 17) + s[2] : g_rotl64(s[1] * 5, 7) * 9;
  const uint64_t t = s[1] << 11;
  s[2] ^= s[0]; s[5] ^= s[1]; s[1] ^= s[2]; s[7] ^= s[3];
  s[3] ^= s[4]; s[4] ^= s[5]; s[0] ^= s[6]; s[6] ^= s[7];
  s[6] ^= t;
  s[7] = g_rotl64(s[7], 21);
  return result;
}

template <class State> struct GenTraits;
template <> struct GenTraits<X256> { using int_type = uint64_t; static constexpr int N = 4; };
template <> struct GenTraits<X128> { using int_type = uint32_t; static constexpr int N = 4; };
template <> struct GenTraits<XR128> { using int_type = uint64_t; static constexpr int N = 2; };
template <> struct GenTraits<X512> { using int_type = uint64_t; static constexpr int N = 8; };

// the samplers' view of a generator (samplers.cuh: every sampler is a template on this)
template <class State, int SCR> struct GenRng {
  using state_type = State;
  using int_type = typename GenTraits<State>::int_type;
  State s;
  __host__ __device__ __forceinline__ int_type next() { return gen_next<SCR>(s); }
};

// jump / long_jump polynomials, word 0 bit 0 first (hdr/xoshiro128.hpp:25-60,
// xoroshiro128.hpp:25-66, xoshiro256.hpp:25-63, xoshiro512.hpp:25-72)
template <class State, int SCR> struct JumpPoly;
template <int SCR> struct JumpPoly<X256, SCR> {
  __host__ __device__ static uint64_t jump(int i) {
    const uint64_t c[4] = {0x180ec6d33cfd0abaull, 0xd5a61266f0c9392cull, 0xa9582618e03fc9aaull, 0x39abdc4529b1661cull};
    return c[i];
  }
  __host__ __device__ static uint64_t long_jump(int i) {
    const uint64_t c[4] = {0x76e15d3efefdcbbfull, 0xc5004e441c522fb3ull, 0x77710069854ee241ull, 0x39109bb02acbe635ull};
    return c[i];
  }
};
template <int SCR> struct JumpPoly<X128, SCR> {
  __host__ __device__ static uint32_t jump(int i) {
    const uint32_t c[4] = {0x8764000bu, 0xf542d2d3u, 0x6fa035c3u, 0x77f2db5bu};
    return c[i];
  }
  __host__ __device__ static uint32_t long_jump(int i) {
    const uint32_t c[4] = {0xb523952eu, 0x0b6f099fu, 0xccf5a0efu, 0x1c580662u};
    return c[i];
  }
};
template <int SCR> struct JumpPoly<XR128, SCR> {
  __host__ __device__ static uint64_t jump(int i) {
    const uint64_t p[2] = {0xdf900294d8f554a5ull, 0x170865df4b3201fcull};
    const uint64_t pp[2] = {0x2bd7a6a6e99c2ddcull, 0x0992ccaf6a6fca05ull};
    return SCR == kPlusPlus ? pp[i] : p[i];
  }
  __host__ __device__ static uint64_t long_jump(int i) {
    const uint64_t p[2] = {0xd2a98b26625eee7bull, 0xdddf9b1090aa7ac1ull};
    const uint64_t pp[2] = {0x360fd5f2cf8d5d99ull, 0x9c6e6877736c46e3ull};
    return SCR == kPlusPlus ? pp[i] : p[i];
  }
};
template <int SCR> struct JumpPoly<X512, SCR> {
  __host__ __device__ static uint64_t jump(int i) {
    const uint64_t c[8] = {0x33ed89b6e7a353f9ull, 0x760083d7955323beull, 0x2837f2fbb5f22faeull, 0x4b8c5674d309511cull,
                           0xb11ac47a7ba28c25ull, 0xf1be7667092bcc1cull, 0x53851efdb6df0aafull, 0x1ebbc8b23eaf25dbull};
    return c[i];
  }
  __host__ __device__ static uint64_t long_jump(int i) {
    const uint64_t c[8] = {0x11467fef8f921d28ull, 0xa2a819f2e79c8ea8ull, 0xa8299fc284b3959aull, 0xb4d347340ca63ee1ull,
                           0x1cb0940bedbff6ceull, 0xd956c5c4fa1f8e17ull, 0x915e38fd4eda93bcull, 0x5b3ccdfa5d7daca5ull};
    return c[i];
  }
};

// the state as an array of its words (the exported layout, hdr/prng.hpp:92-105)
template <class State>
__host__ __device__ __forceinline__ typename GenTraits<State>::int_type* gen_words(State& s) {
  return reinterpret_cast<typename GenTraits<State>::int_type*>(&s);
}

// hdr/generator.hpp:86-106
template <class State, int SCR, bool LONG>
__host__ __device__ inline void gen_jump(State& s) {
  using W = typename GenTraits<State>::int_type;
  constexpr int N = GenTraits<State>::N;
  constexpr int BITS = static_cast<int>(sizeof(W)) * 8;
  W work[N];
  for (int j = 0; j < N; ++j) work[j] = 0;
  W* w = gen_words(s);
  for (int i = 0; i < N; ++i) {
    const W coef = LONG ? JumpPoly<State, SCR>::long_jump(i) : JumpPoly<State, SCR>::jump(i);
    for (int b = 0; b < BITS; ++b) {
      if ((coef >> b) & static_cast<W>(1)) {
        for (int j = 0; j < N; ++j) work[j] ^= w[j];
      }
      gen_next<SCR>(s);
    }
  }
  for (int j = 0; j < N; ++j) w[j] = work[j];
}

}  // namespace mb200
