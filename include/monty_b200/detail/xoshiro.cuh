// xoshiro.cuh -- device-side generator: xoshiro256 {+,++,**} state held in
// registers, next(), int->real conversion, and the generic polynomial jump.
//
// ref: inst/include/monty/random/xoshiro256.hpp:65-103 (next),
//      utils.hpp:14-44 (int_to_real), generator.hpp:86-106 (rng_jump_state).
// "hdr/" below abbreviates inst/include/monty/random/ of the reference.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mb200 {

enum Scrambler { kPlus = 0, kPlusPlus = 1, kStarStar = 2 };

// One stream's generator state, 4 x u64, exactly the reference's exported
// layout (hdr/prng.hpp:92-105).  Lives in registers while a kernel runs.
struct X256 {
  uint64_t s0, s1, s2, s3;
};

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) {
  return (x << k) | (x >> (64 - k));
}

// 128-bit vectorised state load / store: a stream's state is 32 contiguous,
// 32-byte aligned bytes (cudaMalloc'd base, 32 B stride).
__device__ __forceinline__ X256 load_state(const uint64_t* __restrict__ state, size_t stream) {
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(state + 4 * stream);
  const ulonglong2 a = p[0];
  const ulonglong2 b = p[1];
  return X256{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void store_state(uint64_t* __restrict__ state, size_t stream, const X256& s) {
  ulonglong2* p = reinterpret_cast<ulonglong2*>(state + 4 * stream);
  p[0] = make_ulonglong2(s.s0, s.s1);
  p[1] = make_ulonglong2(s.s2, s.s3);
}

// The linear state transition shared by the three scramblers
// (hdr/xoshiro256.hpp:68-75), on explicit 32-bit halves so that it costs the
// minimum the ALU pipe allows -- 8 three-input LOP3, 3 funnel shifts and one
// shift (which the compiler issues as IMAD.SHL on the FMA pipe):
//   s1' = s1 ^ s2 ^ s0,  s0' = s0 ^ s3 ^ s1,  s2' = s2 ^ s0 ^ (s1 << 17),
//   s3' = rotl(s3 ^ s1, 45) = swap halves, then rotate left by 13
// The 64-bit formulation made ptxas spend 5 more ALU ops per step on the
// rotate and on two-input xors; the uniform kernel is ALU-pipe bound, so this
// is worth 25% of its throughput (profiles/r01_baseline_*_details.csv).
__device__ __forceinline__ void x256_step(X256& s) {
  const uint32_t s0l = static_cast<uint32_t>(s.s0), s0h = static_cast<uint32_t>(s.s0 >> 32);
  const uint32_t s1l = static_cast<uint32_t>(s.s1), s1h = static_cast<uint32_t>(s.s1 >> 32);
  const uint32_t s2l = static_cast<uint32_t>(s.s2), s2h = static_cast<uint32_t>(s.s2 >> 32);
  const uint32_t s3l = static_cast<uint32_t>(s.s3), s3h = static_cast<uint32_t>(s.s3 >> 32);
  const uint32_t tl = s1l << 17;
  const uint32_t th = __funnelshift_l(s1l, s1h, 17);
  const uint32_t n3l = s3l ^ s1l, n3h = s3h ^ s1h;
  const uint32_t r1l = s1l ^ s2l ^ s0l, r1h = s1h ^ s2h ^ s0h;
  const uint32_t r0l = s0l ^ n3l, r0h = s0h ^ n3h;
  const uint32_t r2l = s2l ^ s0l ^ tl, r2h = s2h ^ s0h ^ th;
  // rotl64(n3, 45): (hi, lo) -> hi' = (lo << 13) | (hi >> 19), lo' = (hi << 13) | (lo >> 19)
  const uint32_t r3h = __funnelshift_l(n3h, n3l, 13);
  const uint32_t r3l = __funnelshift_l(n3l, n3h, 13);
  s.s0 = (static_cast<uint64_t>(r0h) << 32) | r0l;
  s.s1 = (static_cast<uint64_t>(r1h) << 32) | r1l;
  s.s2 = (static_cast<uint64_t>(r2h) << 32) | r2l;
  s.s3 = (static_cast<uint64_t>(r3h) << 32) | r3l;
}

template <int SCR>
__device__ __forceinline__ uint64_t x256_next(X256& s) {
  uint64_t result;
  if (SCR == kPlus) {
    result = s.s0 + s.s3;                      // hdr/xoshiro256.hpp:93
  } else if (SCR == kPlusPlus) {
    result = rotl64(s.s0 + s.s3, 23) + s.s0;   // hdr/xoshiro256.hpp:80
  } else {
    result = rotl64(s.s1 * 5, 7) * 9;          // hdr/xoshiro256.hpp:67
  }
  x256_step(s);
  return result;
}

// The sampler generator is xoshiro256+ (the reference's default_rng64,
// src/random.cpp:10).
// The samplers (samplers.cuh) are templates on the generator: anything with
// `state_type s` and `int_type next()` (generators.cuh has the other eleven).
struct Rng {
  using state_type = X256;
  using int_type = uint64_t;
  X256 s;
  __device__ __forceinline__ uint64_t next() { return x256_next<kPlus>(s); }
};

// hdr/utils.hpp:21-25: (x >> 11) * 2^-53 + 2^-54.  Clearing the low 11 bits
// leaves at most 53 significant bits, so the u64 -> f64 conversion is exact
// and equals (x >> 11) * 2^11; scaling by 2^-64 is exact too, hence one
// explicit FMA rounds exactly like the reference's multiply followed by add.
// (One LOP3 on the low word instead of a two-instruction 64-bit shift.)
__device__ __forceinline__ double u64_to_double(uint64_t x) {
  return __fma_rn(__ull2double_rn(x & ~0x7ffull), 5.421010862427522170e-20 /* 2^-64 */,
                  (1.1102230246251565e-16 / 2.0));
}
// hdr/utils.hpp:33-38: the top 32 bits, converted to float (one rounding),
// scaled by 2^-32 (exact; 2.3283064e-10f is 2^-32), plus 2^-33.
__device__ __forceinline__ float u64_to_float(uint64_t x) {
  const uint32_t t = static_cast<uint32_t>(x >> 32);
  return __fmaf_rn(static_cast<float>(t), 2.3283064e-10f, (2.3283064e-10f / 2.0f));
}

// hdr/utils.hpp:27-31, 40-44: a 32-bit generator's word.  x * 2^-32 is exact in double and
// so is adding 2^-33; in float the conversion rounds once, the scaling is exact and the
// addition rounds once, which one FMA reproduces.
__device__ __forceinline__ double u32_to_double(uint32_t x) {
  return __fma_rn(static_cast<double>(x), 2.3283064365386963e-10, (2.3283064365386963e-10 / 2.0));
}
__device__ __forceinline__ float u32_to_float(uint32_t x) {
  return __fmaf_rn(static_cast<float>(x), 2.3283064e-10f, (2.3283064e-10f / 2.0f));
}

template <typename T> struct IntToReal;
template <> struct IntToReal<double> {
  __device__ __forceinline__ static double from(uint64_t x) { return u64_to_double(x); }
  __device__ __forceinline__ static double from(uint32_t x) { return u32_to_double(x); }
};
template <> struct IntToReal<float> {
  __device__ __forceinline__ static float from(uint64_t x) { return u64_to_float(x); }
  __device__ __forceinline__ static float from(uint32_t x) { return u32_to_float(x); }
};
template <typename T, typename U> __device__ __forceinline__ T int_to_real(U x) { return IntToReal<T>::from(x); }

// hdr/generator.hpp:154-159
template <typename T, typename R>
__device__ __forceinline__ T random_real(R& r) {
  return int_to_real<T>(r.next());
}

// hdr/generator.hpp:86-106: advance by the polynomial `coef` (256 bits, word 0
// bit 0 first): acc ^= state when the bit is set, then step.  With coef = the
// reference's jump / long_jump constants this IS jump() / long_jump(); with
// coef = x^(k * 2^128) mod P it is k jumps at the cost of one (see
// jump_poly.h).
__device__ __forceinline__ void x256_apply_poly(X256& s, const uint64_t* __restrict__ coef) {
  uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 1
  for (int w = 0; w < 4; ++w) {
    const uint64_t c = coef[w];
#pragma unroll 4
    for (int b = 0; b < 64; ++b) {
      const uint64_t m = 0 - ((c >> b) & 1ull);   // all-ones when the bit is set
      a0 ^= s.s0 & m; a1 ^= s.s1 & m; a2 ^= s.s2 & m; a3 ^= s.s3 & m;
      x256_step(s);
    }
  }
  s.s0 = a0; s.s1 = a1; s.s2 = a2; s.s3 = a3;
}

}  // namespace mb200
