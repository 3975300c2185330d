// jump_constants.cuh -- the xoshiro256 jump / long-jump polynomials as device
// constants (ref: inst/include/monty/random/xoshiro256.hpp:25-63), for code
// that jumps a register-resident state inline (x256_apply_poly, xoshiro.cuh).
#pragma once
#include <cstdint>

namespace mb200 {
static __device__ const uint64_t kJump256[4] = {0x180ec6d33cfd0abaull, 0xd5a61266f0c9392cull,
                                                0xa9582618e03fc9aaull, 0x39abdc4529b1661cull};
static __device__ const uint64_t kLongJump256[4] = {0x76e15d3efefdcbbfull, 0xc5004e441c522fb3ull,
                                                    0x77710069854ee241ull, 0x39109bb02acbe635ull};
}  // namespace mb200
