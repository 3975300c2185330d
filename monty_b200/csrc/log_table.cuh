// log_table.cuh -- table and polynomial of the double-precision log used by the
// samplers (samplers.cuh: fast_log).  32 rows of 16 bytes over [0.6796875,
// 1.359375): {1/c_i as a float, the low part of log(c_i) as a float, the high
// part (42 bits)}; c_20 == 1 exactly so that arguments next to 1 need no special
// path.  The whole table is four cache lines, so the 32 look-ups of a warp cost
// at most four L1 wavefronts (a 128-row table cost 17 and made the exponential
// sampler L1-bound).  Generated with mpmath at 300 bits (tools/make_log_table.py);
// the values are data of this repository, not of the reference.
#pragma once
namespace mb200 {
// offset 0x3fe5c00000000000; P(r) = -1/2 + r c1 + ... + r^7 c7
constexpr unsigned long long kLogOff = 0x3fe5c00000000000ull;
static __constant__ double c_log_poly[7] = {0x1.5555555555555p-2, -0x1.0000000001b30p-2, 0x1.999999999cb08p-3, -0x1.5555544ad6d35p-3, 0x1.249248324d77ap-3, -0x1.001a1f3c514b0p-3, 0x1.c74bf070fdf8dp-4};
struct __align__(16) LogRow { float invc, lo; double hi; };
static __device__ const LogRow g_log_tab[32] = {
  {0x1.745d180000000p+0f, -0x1.44fb7a0000000p-44f, -0x1.7fafa5bd81000p-2},
  {0x1.6c16c20000000p+0f, -0x1.0769320000000p-45f, -0x1.68ac8589c6800p-2},
  {0x1.642c860000000p+0f, -0x1.ea57080000000p-45f, -0x1.522ae1b38a000p-2},
  {0x1.5c98820000000p+0f, -0x1.7aad4c0000000p-46f, -0x1.3c25255333000p-2},
  {0x1.5555560000000p+0f, -0x1.c53c1e0000000p-45f, -0x1.269623134d800p-2},
  {0x1.4e5e0a0000000p+0f, -0x1.1e058c0000000p-44f, -0x1.1178e6c27e000p-2},
  {0x1.47ae140000000p+0f, -0x1.b83ecc0000000p-46f, -0x1.f991c3cb3b000p-3},
  {0x1.4141420000000p+0f, -0x1.9932060000000p-45f, -0x1.d10383e655800p-3},
  {0x1.3b13b20000000p+0f, -0x1.ca6f2c0000000p-47f, -0x1.a93ed8c8ad800p-3},
  {0x1.3521d00000000p+0f, -0x1.deddba0000000p-46f, -0x1.823c18551a000p-3},
  {0x1.2f684c0000000p+0f, -0x1.6c3ee00000000p-45f, -0x1.5bf407b543800p-3},
  {0x1.29e4120000000p+0f, -0x1.07ea080000000p-53f, -0x1.365fc6c159000p-3},
  {0x1.24924a0000000p+0f, -0x1.15f78c0000000p-45f, -0x1.1178ee227e000p-3},
  {0x1.1f70480000000p+0f, -0x1.a814020000000p-46f, -0x1.da72783844000p-4},
  {0x1.1a7b960000000p+0f, -0x1.882e1e0000000p-48f, -0x1.9335e4d594800p-4},
  {0x1.15b1e60000000p+0f, -0x1.ab0f6a0000000p-46f, -0x1.4d31165207800p-4},
  {0x1.1111120000000p+0f, -0x1.a488a40000000p-48f, -0x1.08599959e3800p-4},
  {0x1.0c97140000000p+0f, -0x1.311a8c0000000p-48f, -0x1.894a8349fb000p-5},
  {0x1.0842100000000p+0f, -0x1.011c060000000p-47f, -0x1.0415c89e74000p-5},
  {0x1.0410420000000p+0f, -0x1.99b27c0000000p-48f, -0x1.0205a38935000p-6},
  {0x1.0000000000000p+0f, 0x0.0p+0f, 0x0.0p+0},
  {0x1.f07c200000000p-1f, 0x1.c0267c0000000p-49f, 0x1.f82990e783000p-6},
  {0x1.e1e1e20000000p-1f, 0x1.53b0be0000000p-48f, 0x1.f0a30a0116000p-5},
  {0x1.d41d420000000p-1f, 0x1.a65df20000000p-47f, 0x1.6f0d272e56800p-4},
  {0x1.c71c720000000p-1f, 0x1.73f4f60000000p-47f, 0x1.e27074e2af000p-4},
  {0x1.bacf920000000p-1f, 0x1.4b77020000000p-45f, 0x1.29552c41ff000p-3},
  {0x1.af286c0000000p-1f, 0x1.ea643a0000000p-46f, 0x1.5ff3060a79000p-3},
  {0x1.a41a420000000p-1f, 0x1.ade1900000000p-45f, 0x1.9525a80f45000p-3},
  {0x1.99999a0000000p-1f, 0x1.12d6120000000p-46f, 0x1.c8ff7a79a9800p-3},
  {0x1.8f9c180000000p-1f, 0x1.90e3560000000p-45f, 0x1.fb918bd5e3800p-3},
  {0x1.8618620000000p-1f, 0x1.8448e80000000p-44f, 0x1.1675c97aba000p-2},
  {0x1.7d05f40000000p-1f, 0x1.4c2f0c0000000p-44f, 0x1.2e8e2bee11800p-2},
};
}  // namespace mb200
