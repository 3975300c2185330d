// jump_poly.h -- host-side GF(2) polynomial arithmetic for xoshiro256 jumps.
//
// The reference builds an n-stream generator by n-1 SEQUENTIAL jump() calls
// (ref: inst/include/monty/random/prng.hpp:38-53; 1.25 us each, 21 s for 2^24
// streams) and jumps "n times" by looping (src/random.cpp:382-396).  jump() is
//     state <- c(T) state,   c(x) = x^(2^128) mod P(x)
// where T is the generator's linear transition over GF(2)^256, P its
// characteristic polynomial and c the published jump constants
// (generator.hpp:86-106 evaluates c(T) state bit by bit).  Hence k jumps are
// ONE evaluation of c(x)^k mod P, and stream i of a generator is
// (c^i mod P)(T) applied to stream 0.  This header computes those
// polynomials: P by Berlekamp-Massey on the generator's own output bits,
// products mod P by shift-and-xor.  It self-checks against the reference's
// jump and long-jump constants (x^(2^128), x^(2^192) mod P must reproduce
// them) so a mistake here cannot silently produce different streams.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace mb200 {

using Poly256 = std::array<uint64_t, 4>;   // bit b of word w = coefficient of x^(64 w + b)

// ref: hdr/xoshiro256.hpp:25-63
constexpr Poly256 kJumpPoly = {0x180ec6d33cfd0abaull, 0xd5a61266f0c9392cull,
                               0xa9582618e03fc9aaull, 0x39abdc4529b1661cull};
constexpr Poly256 kLongJumpPoly = {0x76e15d3efefdcbbfull, 0xc5004e441c522fb3ull,
                                   0x77710069854ee241ull, 0x39109bb02acbe635ull};

inline void host_x256_step(uint64_t s[4]) {   // hdr/xoshiro256.hpp:68-75
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
  s[2] ^= t;
  s[3] = (s[3] << 45) | (s[3] >> 19);
}

inline uint64_t host_splitmix64(uint64_t seed) {   // hdr/utils.hpp:72-77
  uint64_t z = (seed += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

class JumpAlgebra {
 public:
  JumpAlgebra() {
    find_characteristic_polynomial();
    // self-check against the reference's constants
    Poly256 p = {2, 0, 0, 0};   // x
    for (int i = 0; i < 128; ++i) p = mul(p, p);
    if (p != kJumpPoly) throw std::runtime_error("jump polynomial self-check failed");
    for (int i = 0; i < 64; ++i) p = mul(p, p);
    if (p != kLongJumpPoly) throw std::runtime_error("long-jump polynomial self-check failed");
  }

  // a * b mod P
  Poly256 mul(const Poly256& a, const Poly256& b) const {
    Poly256 r = {0, 0, 0, 0};
    for (int i = 255; i >= 0; --i) {
      const bool carry = (r[3] >> 63) & 1;
      r[3] = (r[3] << 1) | (r[2] >> 63);
      r[2] = (r[2] << 1) | (r[1] >> 63);
      r[1] = (r[1] << 1) | (r[0] >> 63);
      r[0] <<= 1;
      if (carry) for (int w = 0; w < 4; ++w) r[w] ^= p_low_[w];
      if ((b[i >> 6] >> (i & 63)) & 1) for (int w = 0; w < 4; ++w) r[w] ^= a[w];
    }
    return r;
  }

  // base^e mod P
  Poly256 pow(Poly256 base, uint64_t e) const {
    Poly256 r = {1, 0, 0, 0};
    while (e) {
      if (e & 1) r = mul(r, base);
      e >>= 1;
      if (e) base = mul(base, base);
    }
    return r;
  }

  Poly256 jumps(uint64_t n) const { return pow(kJumpPoly, n); }
  Poly256 long_jumps(uint64_t n) const { return pow(kLongJumpPoly, n); }
  // c^(2^k): the polynomial that jumps 2^k streams ahead (doubling ladder)
  Poly256 jump_pow2(int k) const {
    Poly256 p = kJumpPoly;
    for (int i = 0; i < k; ++i) p = mul(p, p);
    return p;
  }

  // host evaluation of c(T) state (generator.hpp:86-106), for the handful of
  // states the host needs before the device ladder starts
  static void apply(const Poly256& c, uint64_t s[4]) {
    uint64_t acc[4] = {0, 0, 0, 0};
    for (int w = 0; w < 4; ++w) {
      for (int b = 0; b < 64; ++b) {
        if ((c[w] >> b) & 1) for (int j = 0; j < 4; ++j) acc[j] ^= s[j];
        host_x256_step(s);
      }
    }
    for (int j = 0; j < 4; ++j) s[j] = acc[j];
  }

 private:
  Poly256 p_low_;   // P(x) - x^256

  void find_characteristic_polynomial() {
    // 512 output bits of an arbitrary non-zero state
    uint64_t s[4] = {0x9e3779b97f4a7c15ull, 0xbf58476d1ce4e5b9ull, 0x94d049bb133111ebull, 1};
    const int N = 512;
    std::vector<uint8_t> seq(N);
    for (int i = 0; i < N; ++i) { seq[i] = s[0] & 1; host_x256_step(s); }
    // Berlekamp-Massey over GF(2): connection polynomial C with
    // seq[n] = sum_{i=1..L} C[i] seq[n-i]
    std::vector<uint8_t> C(N + 1, 0), B(N + 1, 0);
    C[0] = B[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < N; ++n) {
      uint8_t d = seq[n];
      for (int i = 1; i <= L; ++i) d ^= C[i] & seq[n - i];
      if (!d) { ++m; continue; }
      if (2 * L <= n) {
        std::vector<uint8_t> T = C;
        for (int i = 0; i + m <= N; ++i) C[i + m] ^= B[i];
        L = n + 1 - L;
        B = T;
        m = 1;
      } else {
        for (int i = 0; i + m <= N; ++i) C[i + m] ^= B[i];
        ++m;
      }
    }
    if (L != 256) throw std::runtime_error("unexpected linear complexity of xoshiro256");
    // characteristic polynomial P(x) = x^L C(1/x): coefficient of x^(L-i) is C[i]
    p_low_ = {0, 0, 0, 0};
    for (int i = 1; i <= L; ++i) {
      if (C[i]) { const int e = L - i; p_low_[e >> 6] |= 1ull << (e & 63); }
    }
  }
};

}  // namespace mb200
