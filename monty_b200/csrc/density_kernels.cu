// density_kernels.cu -- elementwise (log-)density kernels with a fused,
// deterministic block/grid sum (the log-likelihood of SURVEY config C5).
//
// ref: src/density.cpp:6-254: `for i < n: ret[i] = density::<d>(x[i], p[i]...)`
// with every argument indexed by i (no recycling); x and the size-like
// arguments of the count distributions arrive as int32, the rest as double.
//
// One grid-stride pass: thread t evaluates elements t, t + stride, ... and
// keeps a running double sum; the block reduces with warp shuffles and writes
// one partial per block; reduce_partials_kernel adds the partials in a fixed
// order, so the sum is reproducible run to run for a given launch geometry.
#include "../../include/monty_b200.h"
#include "density.cuh"
#include "density_dispatch.h"
#include "launch_args.h"

#include <cstdlib>

namespace mb200 {

constexpr int kDensBlock = 256;

template <typename T>
__device__ __forceinline__ T ld_arg(const void* p, int is_i32, unsigned long long i) {
  return is_i32 ? static_cast<T>(static_cast<const int*>(p)[i])
                : static_cast<T>(static_cast<const double*>(p)[i]);
}

// number of parameters of density DENS (ref: src/density.cpp signatures)
template <int DENS> struct DensNP {
  static constexpr int value =
    (DENS == MB200_DENS_POISSON || DENS == MB200_DENS_EXPONENTIAL_RATE || DENS == MB200_DENS_EXPONENTIAL_MEAN) ? 1 :
    (DENS == MB200_DENS_BETA_BINOMIAL_PROB || DENS == MB200_DENS_BETA_BINOMIAL_AB || DENS == MB200_DENS_HYPERGEOMETRIC ||
     DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU || DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB) ? 3 :
    (DENS == MB200_DENS_TRUNCATED_NORMAL) ? 4 : 2;
};

template <int DENS, typename T>
__device__ __forceinline__ T density_compute(T x, const T* p, bool lg) {
  switch (DENS) {
  case MB200_DENS_BINOMIAL: return dens_binomial<T>(x, p[0], p[1], lg);
  case MB200_DENS_NORMAL: return dens_normal<T>(x, p[0], p[1], lg);
  case MB200_DENS_NEGATIVE_BINOMIAL_MU: return dens_negative_binomial_mu<T>(x, p[0], p[1], lg);
  case MB200_DENS_NEGATIVE_BINOMIAL_PROB: return dens_negative_binomial_prob<T>(x, p[0], p[1], lg);
  case MB200_DENS_BETA_BINOMIAL_PROB: return dens_beta_binomial_prob<T>(x, p[0], p[1], p[2], lg);
  case MB200_DENS_BETA_BINOMIAL_AB: return dens_beta_binomial_ab<T>(x, p[0], p[1], p[2], lg);
  case MB200_DENS_POISSON: return dens_poisson<T>(x, p[0], lg);
  case MB200_DENS_UNIFORM: return dens_uniform<T>(x, p[0], p[1], lg);
  case MB200_DENS_EXPONENTIAL_RATE: return dens_exponential_rate<T>(x, p[0], lg);
  case MB200_DENS_EXPONENTIAL_MEAN: return dens_exponential_mean<T>(x, p[0], lg);
  case MB200_DENS_GAMMA_RATE: return dens_gamma_rate<T>(x, p[0], p[1], lg);
  case MB200_DENS_GAMMA_SCALE: return dens_gamma_scale<T>(x, p[0], p[1], lg);
  case MB200_DENS_BETA: return dens_beta<T>(x, p[0], p[1], lg);
  case MB200_DENS_HYPERGEOMETRIC: return dens_hypergeometric<T>(x, p[0], p[1], p[2], lg);
  case MB200_DENS_LOG_NORMAL: return dens_log_normal<T>(x, p[0], p[1], lg);
  case MB200_DENS_TRUNCATED_NORMAL: return dens_truncated_normal<T>(x, p[0], p[1], p[2], p[3], lg);
  case MB200_DENS_CAUCHY: return dens_cauchy<T>(x, p[0], p[1], lg);
  case MB200_DENS_WEIBULL: return dens_weibull<T>(x, p[0], p[1], lg);
  case MB200_DENS_ZI_POISSON:
    return dens_zi_wrap<T>(x, p[0], dens_poisson<T>(x, p[1], true), lg);
  case MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU:
    return dens_zi_wrap<T>(x, p[0], dens_negative_binomial_mu<T>(x, p[1], p[2], true), lg);
  case MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB:
    return dens_zi_wrap<T>(x, p[0], dens_negative_binomial_prob<T>(x, p[1], p[2], true), lg);
  default: return 0;
  }
}

// Four observations per thread per trip, all their loads issued before the
// first use: the count densities became short once lgamma of integers is a
// table look-up, and with one element in flight per thread the kernel waited
// on HBM latency (ncu: long-scoreboard stall 11.9 per issue,
// profiles/r01_density_poisson_table_ncu_summary.txt).  Element order per
// thread and the reduction tree are fixed, so the sum stays reproducible.
// Only for the light densities: the lgamma-heavy ones (non-integer arguments)
// are FP64-pipe bound and lose occupancy to the extra registers.
template <int DENS> struct DensUnroll {
  static constexpr int value =
    (DENS == MB200_DENS_POISSON || DENS == MB200_DENS_NORMAL || DENS == MB200_DENS_UNIFORM ||
     DENS == MB200_DENS_EXPONENTIAL_RATE || DENS == MB200_DENS_EXPONENTIAL_MEAN || DENS == MB200_DENS_CAUCHY ||
     DENS == MB200_DENS_LOG_NORMAL || DENS == MB200_DENS_WEIBULL || DENS == MB200_DENS_ZI_POISSON) ? 4 : 1;
};

template <int DENS, typename T>
__global__ void __launch_bounds__(kDensBlock)
density_kernel(const DensityArgs a) {
  constexpr int NPD = DensNP<DENS>::value;
  constexpr int kDensUnroll = DensUnroll<DENS>::value;
  double acc = 0.0;
  const bool lg = a.log != 0;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kDensBlock;
  for (unsigned long long i0 = static_cast<unsigned long long>(blockIdx.x) * kDensBlock + threadIdx.x;
       i0 < a.n; i0 += stride * kDensUnroll) {
    T x[kDensUnroll], p[kDensUnroll][NPD];
#pragma unroll
    for (int u = 0; u < kDensUnroll; ++u) {
      const unsigned long long i = i0 + u * stride;
      if (i < a.n) {
        x[u] = ld_arg<T>(a.x, a.x_is_i32, i);
#pragma unroll
        for (int k = 0; k < NPD; ++k) p[u][k] = ld_arg<T>(a.par[k], a.par_is_i32[k], i);
      }
    }
#pragma unroll
    for (int u = 0; u < kDensUnroll; ++u) {
      const unsigned long long i = i0 + u * stride;
      if (i < a.n) {
        const double v = static_cast<double>(density_compute<DENS, T>(x[u], p[u], lg));
        if (a.out) a.out[i] = v;
        acc += v;
      }
    }
  }
  if (a.partial) {
    __shared__ double warp_sum[kDensBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kDensBlock / 32; ++w) s += warp_sum[w];
      a.partial[blockIdx.x] = s;
    }
  }
}

// ---------------------------------------------------------------------------
// The tiled kernel: the path the reference's own call shapes take.
//
// ref: src/density.cpp passes x and the size-like arguments of the count
// densities as cpp11::integers and everything else as doubles -- the "canonical"
// layout of each density (DensCanon).  With it the element types are compile-
// time facts, a CTA owns TILES of 1024 consecutive observations (thread t the
// observations 2t, 2t + 1 of each half of the tile, read with two 128-bit shared
// loads per double array and two 64-bit ones per int32 array, conflict free),
// and the tiles arrive by 1-D TMA bulk copies
// (cp.async.bulk ... mbarrier::complete_tx) into a ring of kDensStages shared
// buffers: one thread arms the stage's mbarrier with the byte count and issues
// one copy per argument array; the copy of tile k + kDensStages is in flight
// while tile k is evaluated.  No registers are spent on prefetching (the
// register double buffer of round 1 cost a CTA per SM) and the loads no longer
// stall the warps that issue them (round 1: long-scoreboard 5.6 per issue on
// scalar 4 / 8-byte loads of four separate sectors per thread).
//
// The last, partial tile is read with plain loads.  Per-thread order, the block
// tree and the order of the partials are fixed: the sum is reproducible.
// ---------------------------------------------------------------------------
constexpr int kDensTile = 4 * kDensBlock;     // observations per tile
// ring depth: a power of two (stage = it & 3, phase = it >> 2: no division in the loop);
// measured on the Poisson log-likelihood: 2 stages 311, 3 stages 292, 4 stages 316, 5 stages 299 Gobs/s
#ifndef MB200_DENS_STAGES
#define MB200_DENS_STAGES 4
#endif
constexpr int kDensStages = MB200_DENS_STAGES;

// bit 0: x is int32; bit k + 1: parameter k is int32 (ref: src/density.cpp signatures)
template <int DENS> struct DensCanon {
  static constexpr unsigned mask =
    (DENS == MB200_DENS_BINOMIAL || DENS == MB200_DENS_BETA_BINOMIAL_PROB || DENS == MB200_DENS_BETA_BINOMIAL_AB) ? 0x3u :
    (DENS == MB200_DENS_HYPERGEOMETRIC) ? 0xFu :
    (DENS == MB200_DENS_NEGATIVE_BINOMIAL_MU || DENS == MB200_DENS_NEGATIVE_BINOMIAL_PROB || DENS == MB200_DENS_POISSON ||
     DENS == MB200_DENS_ZI_POISSON || DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU ||
     DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB) ? 0x1u : 0x0u;
  static constexpr int n_arrays = 1 + DensNP<DENS>::value;
  __host__ __device__ static constexpr bool is_int(int j) { return (mask >> j) & 1u; }
  __host__ __device__ static constexpr unsigned tile_bytes(int j) { return kDensTile * (is_int(j) ? 4u : 8u); }
  __host__ __device__ static constexpr unsigned offset(int j) {       // of array j inside a stage
    unsigned o = 0;
    for (int i = 0; i < j; ++i) o += tile_bytes(i);
    return o;
  }
  static constexpr unsigned stage_bytes = offset(n_arrays);
};

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile("{\n\t"
               ".reg .pred p;\n\t"
               "MB200_WAIT:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra MB200_DONE;\n\t"
               "bra MB200_WAIT;\n\t"
               "MB200_DONE:\n\t"
               "}" :: "r"(mbar), "r"(parity) : "memory");
}

// This thread's four observations of one argument array, as T: elements 2t, 2t + 1
// of each half of the tile -- consecutive lanes read consecutive 16-byte (double) or
// 8-byte (int32) units, so the shared loads are bank-conflict free (four consecutive
// doubles per thread are a 32-byte stride: two wavefronts per load).
// (is_int is a compile-time fact after unrolling)
template <typename T>
__device__ __forceinline__ void lds_four(uint32_t array_saddr, int tid, bool is_int, T (&v)[4], int* raw = nullptr) {
  if (is_int) {
    int2 a, b;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "r"(array_saddr + tid * 8));
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(b.x), "=r"(b.y) : "r"(array_saddr + (kDensTile / 2) * 4 + tid * 8));
    v[0] = static_cast<T>(a.x); v[1] = static_cast<T>(a.y); v[2] = static_cast<T>(b.x); v[3] = static_cast<T>(b.y);
    if (raw) { raw[0] = a.x; raw[1] = a.y; raw[2] = b.x; raw[3] = b.y; }
  } else {
    const double2 a = lds_f64x2(array_saddr + tid * 16), b = lds_f64x2(array_saddr + (kDensTile / 2) * 8 + tid * 16);
    v[0] = static_cast<T>(a.x); v[1] = static_cast<T>(a.y); v[2] = static_cast<T>(b.x); v[3] = static_cast<T>(b.y);
  }
}

// Four observations at once.  Default: four calls.  Poisson in double (config C5's
// load-bound density) interleaves them: the four lgamma look-ups are issued
// together and the four logs run as one batch (independent chains hide the table
// latency; the window around 1 is worked off jointly) -- the same operations on
// the same values as four calls of dens_poisson (hdr/density.hpp:173-189).
// xi: the observations as the int32 they arrived as (Poisson's canonical layout), or null:
// lgamma(x + 1) is then table[xi + 1] without the round trip through double
template <int DENS, typename T, int NPD>
__device__ __forceinline__ void density_compute4(const T (&x)[4], const T (&p)[NPD > 0 ? NPD : 1][4], bool lg, double (&v)[4],
                                                 const int* xi = nullptr, const double* __restrict__ lgt = nullptr) {
  if constexpr (DENS == MB200_DENS_POISSON && sizeof(T) == 8) {
    double lgm[4], ll[4];
    bool plain = true;
    if (xi != nullptr && lgt != nullptr) {
      // branch-free look-ups (clamped index); the rare observation beyond the table is
      // patched afterwards
      bool beyond = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned i = static_cast<unsigned>(xi[u]);
        beyond = beyond || i >= static_cast<unsigned>(kLgtN - 1);
        lgm[u] = __ldg(lgt + (i < static_cast<unsigned>(kLgtN - 1) ? i : 0u) + 1);      // the bits m_lgamma(x + 1) returns (same table)
      }
      if (beyond) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (static_cast<unsigned>(xi[u]) >= static_cast<unsigned>(kLgtN - 1)) lgm[u] = m_lgamma(x[u] + 1);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) lgm[u] = m_lgamma(x[u] + 1);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      plain = plain && (glibc::f_hi(p[0][u]) - 0x00100000u < 0x7fe00000u);        // positive, finite, normal
    }
    if (plain) {
      double lam[4] = {p[0][0], p[0][1], p[0][2], p[0][3]};
      m_log_normal_batch<4>(lam, ll);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) ll[u] = m_log(p[0][u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double ret = x[u] * ll[u] - p[0][u] - lgm[u];
      if (x[u] == 0 && p[0][u] == 0) ret = 0;
      v[u] = maybe_log(ret, lg);
    }
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      T pu[NPD > 0 ? NPD : 1];
#pragma unroll
      for (int k = 0; k < NPD; ++k) pu[k] = p[k][u];
      v[u] = static_cast<double>(density_compute<DENS, T>(x[u], pu, lg));
    }
  }
}

template <int DENS, typename T>
#ifdef MB200_DENS_TILE_MINBLOCKS          /* occupancy experiments */
__global__ void __launch_bounds__(kDensBlock, MB200_DENS_TILE_MINBLOCKS)
#else
__global__ void __launch_bounds__(kDensBlock)
#endif
density_tile_kernel(const DensityArgs a, const unsigned long long n_full_tiles) {
  using Cn = DensCanon<DENS>;
  constexpr int NPD = DensNP<DENS>::value;
  constexpr int NA = Cn::n_arrays;
  extern __shared__ __align__(128) unsigned char dens_smem[];
  __shared__ __align__(8) unsigned long long mbar_store[2 * kDensStages];   // full[s], then empty[s]
  const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(dens_smem));
  const uint32_t mbar0 = static_cast<uint32_t>(__cvta_generic_to_shared(mbar_store));
  const int tid = threadIdx.x;
  const bool lg = a.log != 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kDensStages; ++s) {
      mbar_init(mbar0 + 8 * s, 1);
      mbar_init(mbar0 + 8 * (kDensStages + s), kDensBlock / 32);      // one arrival per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const void* src[NA];
  src[0] = a.x;
#pragma unroll
  for (int k = 0; k < NPD; ++k) src[k + 1] = a.par[k];

  auto issue = [&](unsigned long long tile, int s) {
    if (tid == 0 && tile < n_full_tiles) {
      const uint32_t mb = mbar0 + 8 * s;
      // the stage was read through the generic proxy; order those reads before the bulk writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(mb, Cn::stage_bytes);
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        const unsigned long long byte0 = tile * Cn::tile_bytes(j);
        bulk_copy_g2s(smem0 + s * Cn::stage_bytes + Cn::offset(j), static_cast<const char*>(src[j]) + byte0,
                      Cn::tile_bytes(j), mb);
      }
    }
  };

  const unsigned long long step = gridDim.x;
#pragma unroll
  for (int s = 0; s < kDensStages; ++s) issue(blockIdx.x + s * step, s);

  double acc = 0.0;
  unsigned it = 0;
  const double* __restrict__ lgt = g_lgt_d;           // lgamma of small integers (samplers.cuh), read once
  for (unsigned long long tile = blockIdx.x; tile < n_full_tiles; tile += step, ++it) {
    const int s = static_cast<int>(it % kDensStages);
    mbar_wait(mbar0 + 8 * s, (it / kDensStages) & 1u);
    const uint32_t base = smem0 + s * Cn::stage_bytes;
    T x[4], p[NPD > 0 ? NPD : 1][4];
    int xi[4];
    lds_four<T>(base + Cn::offset(0), tid, Cn::is_int(0), x, xi);
#pragma unroll
    for (int k = 0; k < NPD; ++k) lds_four<T>(base + Cn::offset(k + 1), tid, Cn::is_int(k + 1), p[k]);
    // This warp holds its observations: one arrival on the stage's "empty" barrier.  Only the
    // issuing thread waits for all eight (a block-wide barrier here made every warp wait for
    // the slowest one of each tile: ncu barrier stall 0.97 per issue, r02 summary).
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(mbar0 + 8 * (kDensStages + s));
    if (tid == 0 && tile + kDensStages * step < n_full_tiles) mbar_wait(mbar0 + 8 * (kDensStages + s), (it / kDensStages) & 1u);
    issue(tile + kDensStages * step, s);
    double v[4];
    density_compute4<DENS, T, NPD>(x, p, lg, v, Cn::is_int(0) ? xi : nullptr, lgt);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u];
    if (a.out) {
      double* o = a.out + tile * kDensTile + 2 * tid;
      *reinterpret_cast<double2*>(o) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(o + kDensTile / 2) = make_double2(v[2], v[3]);
    }
  }
  // the partial last tile: plain loads, by the block whose turn it would be
  const unsigned long long tail0 = n_full_tiles * kDensTile;
  if (tail0 < a.n && blockIdx.x == n_full_tiles % gridDim.x) {
    for (unsigned long long i = tail0 + tid; i < a.n; i += kDensBlock) {
      const T x = ld_arg<T>(a.x, Cn::is_int(0), i);
      T pu[NPD > 0 ? NPD : 1];
#pragma unroll
      for (int k = 0; k < NPD; ++k) pu[k] = ld_arg<T>(a.par[k], Cn::is_int(k + 1), i);
      const double v = static_cast<double>(density_compute<DENS, T>(x, pu, lg));
      if (a.out) a.out[i] = v;
      acc += v;
    }
  }
  if (a.partial) {
    __shared__ double warp_sum[kDensBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) warp_sum[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < kDensBlock / 32; ++w) sum += warp_sum[w];
      a.partial[blockIdx.x] = sum;
    }
  }
}

// ---------------------------------------------------------------------------
// The sorted kernel: the densities whose cost is lgamma of NON-INTEGER arguments
// (negative binomial, beta-binomial).  glibc's lgamma is a different program below 8
// (rational pieces) and above (Stirling); lgamma_positive_batch evaluates the Stirling
// range for every argument of an observation and then works the smaller arguments off
// one per lane per trip, so a warp makes as many trips as its WORST observation has
// small arguments while most lanes idle (ncu: 21 of 32 lanes active, nbinom; a warp of
// 32 random observations almost always contains one with 2+ small arguments).
// Here a CTA classifies a tile of 2048 observations by their number of small arguments
// (a few adds and compares), partitions the tile by that count -- a STABLE partition from
// warp votes and one prefix sum, so the order and with it the fused sum stay reproducible --
// and its warps evaluate 32-observation chunks of the sorted order: a chunk is (nearly)
// pure in its class and makes exactly that many trips.  Per observation nothing changes:
// the same density_compute() on the same values.
// ---------------------------------------------------------------------------
// observations per tile (measured, nbinom / beta-binomial Gobs/s: 1024: 44.7 / 12.8, 2048: 45.2 / 13.5,
// 4096: 33.5 / 7.9 -- two CTAs per SM; handing the chunks out dynamically with per-chunk sums: 41.1 / 13.2)
#ifndef MB200_DS_TILE
#define MB200_DS_TILE 2048
#endif
constexpr int kDsTile = MB200_DS_TILE;
constexpr int kDsPer = kDsTile / kDensBlock;      // observations per thread in the classification
constexpr int kDsClasses = 17;                     // 4 x 4 (arguments in [2, 8), arguments below 2 -- the two loops of
                                                   // lgamma_positive_batch), 16 = beyond the end of the array

template <int DENS> struct DensSorted {
  static constexpr bool value = DENS == MB200_DENS_NEGATIVE_BINOMIAL_MU || DENS == MB200_DENS_NEGATIVE_BINOMIAL_PROB ||
                                DENS == MB200_DENS_BETA_BINOMIAL_AB || DENS == MB200_DENS_BETA_BINOMIAL_PROB ||
                                DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU || DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB ||
                                DENS == MB200_DENS_GAMMA_RATE || DENS == MB200_DENS_GAMMA_SCALE || DENS == MB200_DENS_BETA;
};

// class of an observation: min(3, arguments in [2, 8)) * 4 + min(3, arguments below 2) --
// the trip counts of the two loops of lgamma_positive_batch (the class only steers which
// observations share a warp: it never changes a value, so it may be approximate)
__device__ __forceinline__ void dens_count_arg(double v, int& mid, int& low) {
  mid += (v >= 2.0 && v < 8.0);
  low += (v < 2.0);
}
template <int DENS>
__device__ __forceinline__ int dens_small_args(double x, const double* p) {
  int mid = 0, low = 0;
  if constexpr (DENS == MB200_DENS_NEGATIVE_BINOMIAL_MU || DENS == MB200_DENS_NEGATIVE_BINOMIAL_PROB) {
    dens_count_arg(p[0], mid, low);
    dens_count_arg(x + p[0], mid, low);
  } else if constexpr (DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU || DENS == MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB) {
    dens_count_arg(p[1], mid, low);              // (pi0, size, mu | prob)
    dens_count_arg(x + p[1], mid, low);
  } else if constexpr (DENS == MB200_DENS_GAMMA_RATE || DENS == MB200_DENS_GAMMA_SCALE) {
    dens_count_arg(p[0], mid, low);              // lgamma(shape)
  } else if constexpr (DENS == MB200_DENS_BETA) {
    dens_count_arg(p[0], mid, low);              // lbeta(a, b)
    dens_count_arg(p[1], mid, low);
    dens_count_arg(p[0] + p[1], mid, low);
  } else {
    double a = p[1], b = p[2];
    if constexpr (DENS == MB200_DENS_BETA_BINOMIAL_PROB) {
      const double k = 1.0 / p[2] - 1.0;
      a = p[1] * k;
      b = (1.0 - p[1]) * k;
    }
    const double pp = x + a, qq = p[0] - x + b;
    dens_count_arg(pp, mid, low);
    dens_count_arg(qq, mid, low);
    dens_count_arg(pp + qq, mid, low);
    dens_count_arg(a, mid, low);
    dens_count_arg(b, mid, low);
    dens_count_arg(a + b, mid, low);
  }
  return (mid < 3 ? mid : 3) * 4 + (low < 3 ? low : 3);
}

template <int DENS>
__global__ void __launch_bounds__(kDensBlock)
density_sorted_kernel(const DensityArgs a) {
  using Cn = DensCanon<DENS>;
  constexpr int NPD = DensNP<DENS>::value;
  extern __shared__ __align__(16) unsigned char ds_smem[];
  __shared__ unsigned short s_order[kDsTile];
  __shared__ unsigned short s_cnt[kDsClasses * kDsPer * (kDensBlock / 32)];      // [class][round][warp]
  // the tile's inputs in their own types: array j at ds_off(j)
  auto ds_off = [](int j) { unsigned o = 0; for (int i = 0; i < j; ++i) o += kDsTile * (Cn::is_int(i) ? 4u : 8u); return o; };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool lg = a.log != 0;
  const void* src[1 + NPD];
  src[0] = a.x;
#pragma unroll
  for (int k = 0; k < NPD; ++k) src[k + 1] = a.par[k];
  auto lds_arg = [&](int j, unsigned e) -> double {
    return Cn::is_int(j) ? static_cast<double>(reinterpret_cast<const int*>(ds_smem + ds_off(j))[e])
                         : reinterpret_cast<const double*>(ds_smem + ds_off(j))[e];
  };

  double acc = 0.0;
  const unsigned long long n_tiles = (a.n + kDsTile - 1) / kDsTile;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long base = tile * kDsTile;
    const unsigned count = (a.n - base) < kDsTile ? static_cast<unsigned>(a.n - base) : kDsTile;
    for (int i = tid; i < kDsClasses * kDsPer * (kDensBlock / 32); i += kDensBlock) s_cnt[i] = 0;
    __syncthreads();
    // 1. inputs -> shared memory (coalesced), class and rank inside (class, round, warp)
    unsigned cls[kDsPer], rank[kDsPer];
#pragma unroll
    for (int i = 0; i < kDsPer; ++i) {
      const unsigned e = static_cast<unsigned>(i * kDensBlock + tid);
      cls[i] = kDsClasses - 1;
      if (e < count) {
        double xv, pv[NPD];
#pragma unroll
        for (int j = 0; j <= NPD; ++j) {
          double v;
          if (Cn::is_int(j)) {
            const int iv = static_cast<const int*>(src[j])[base + e];
            reinterpret_cast<int*>(ds_smem + ds_off(j))[e] = iv;
            v = static_cast<double>(iv);
          } else {
            v = static_cast<const double*>(src[j])[base + e];
            reinterpret_cast<double*>(ds_smem + ds_off(j))[e] = v;
          }
          if (j == 0) xv = v; else pv[j - 1] = v;
        }
        const int c = dens_small_args<DENS>(xv, pv);
        cls[i] = static_cast<unsigned>(c < 0 ? 0 : (c > kDsClasses - 2 ? kDsClasses - 2 : c));
      }
      const unsigned same = __match_any_sync(0xffffffffu, cls[i]);
      rank[i] = static_cast<unsigned>(__popc(same & ((1u << lane) - 1u)));
      if (rank[i] == 0) s_cnt[(cls[i] * kDsPer + i) * (kDensBlock / 32) + warp] = static_cast<unsigned short>(__popc(same));
    }
    __syncthreads();
    // 2. exclusive prefix sum over (class, round, warp): 512 counts, one warp, 16 per lane
    if (warp == 0) {
      constexpr int kEach = kDsClasses * kDsPer * (kDensBlock / 32) / 32;
      unsigned local[kEach], sum = 0;
#pragma unroll
      for (int q = 0; q < kEach; ++q) { local[q] = s_cnt[lane * kEach + q]; sum += local[q]; }
      unsigned incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      unsigned run = incl - sum;
#pragma unroll
      for (int q = 0; q < kEach; ++q) { s_cnt[lane * kEach + q] = static_cast<unsigned short>(run); run += local[q]; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kDsPer; ++i) {
      const unsigned pos = s_cnt[(cls[i] * kDsPer + i) * (kDensBlock / 32) + warp] + rank[i];
      if (pos < kDsTile) s_order[pos] = static_cast<unsigned short>(i * kDensBlock + tid);
    }
    __syncthreads();
    // 3. warps evaluate 32-observation chunks of the sorted order (chunk c -> warp c % 8)
    for (unsigned k = static_cast<unsigned>(tid); k < count; k += kDensBlock) {
      const unsigned e = s_order[k];
      const double xv = lds_arg(0, e);
      double pv[NPD];
#pragma unroll
      for (int j = 0; j < NPD; ++j) pv[j] = lds_arg(j + 1, e);
      const double v = density_compute<DENS, double>(xv, pv, lg);
      if (a.out) a.out[base + e] = v;
      acc += v;
    }
    __syncthreads();                      // the tile buffers are reused
  }
  if (a.partial) {
    __shared__ double warp_sum[kDensBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) warp_sum[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < kDensBlock / 32; ++w) sum += warp_sum[w];
      a.partial[blockIdx.x] = sum;
    }
  }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ sum) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *sum = sh[0];
}

int density_grid(unsigned long long n, int sm_count) {
  const unsigned long long need = (n + kDensBlock - 1) / kDensBlock;
  const unsigned long long cap = static_cast<unsigned long long>(sm_count) * 8;
  return static_cast<int>(need < cap ? (need ? need : 1) : cap);
}

// the reference's own argument types, 16-byte aligned arrays, at least one full tile
template <int DENS>
static bool tiled_applies(const DensityArgs& a) {
  using Cn = DensCanon<DENS>;
  // the densities that evaluate lgamma of non-integer arguments are FP64-pipe
  // bound, not load bound: four observations per thread cost them registers and
  // a CTA per SM (negative binomial 39.3 -> 33.1, beta-binomial 14.2 -> 12.6 Gobs/s)
  if (DensUnroll<DENS>::value != 4) return false;
  if (a.n < static_cast<unsigned long long>(kDensTile)) return false;
  auto ok = [](const void* p, int is_i32, bool want_int) {
    return (is_i32 != 0) == want_int && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
  };
  if (!ok(a.x, a.x_is_i32, Cn::is_int(0))) return false;
  for (int k = 0; k < DensNP<DENS>::value; ++k) {
    if (!ok(a.par[k], a.par_is_i32[k], Cn::is_int(k + 1))) return false;
  }
  return a.out == nullptr || (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
}

template <int DENS, typename T>
static cudaError_t launch_tiled(const DensityArgs& a, int sm_count, int* grid_out, cudaStream_t st) {
  using Cn = DensCanon<DENS>;
  const size_t smem = static_cast<size_t>(Cn::stage_bytes) * kDensStages;
  const unsigned long long n_full = a.n / kDensTile;
  // CTAs per SM: as many as shared memory holds (200 KB budget), at most 6
  int per_sm = static_cast<int>((200u * 1024u) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
  unsigned long long grid = static_cast<unsigned long long>(sm_count) * per_sm;
  if (grid > n_full) grid = n_full;
  cudaError_t e = cudaFuncSetAttribute(density_tile_kernel<DENS, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  density_tile_kernel<DENS, T><<<static_cast<unsigned>(grid), kDensBlock, smem, st>>>(a, n_full);
  *grid_out = static_cast<int>(grid);
  return cudaGetLastError();
}

template <int DENS>
static cudaError_t launch_density_t(int real_type, const DensityArgs& a, int sm_count, int* grid_out, cudaStream_t st) {
  static const bool generic_only = std::getenv("MB200_DENSITY_GENERIC") != nullptr;      // A/B measurements
  if (!generic_only && tiled_applies<DENS>(a)) {
    return real_type == MB200_REAL_F32 ? launch_tiled<DENS, float>(a, sm_count, grid_out, st)
                                       : launch_tiled<DENS, double>(a, sm_count, grid_out, st);
  }
  if constexpr (DensSorted<DENS>::value) {
    // double, the reference's own argument types, enough observations to be worth sorting
    using Cn = DensCanon<DENS>;
    bool canon = real_type != MB200_REAL_F32 && a.n >= 4ull * kDsTile && (a.x_is_i32 != 0) == Cn::is_int(0);
    for (int k = 0; k < DensNP<DENS>::value; ++k) canon = canon && (a.par_is_i32[k] != 0) == Cn::is_int(k + 1);
    if (!generic_only && canon) {
      size_t smem = 0;
      for (int j = 0; j <= DensNP<DENS>::value; ++j) smem += kDsTile * (Cn::is_int(j) ? 4u : 8u);
      const unsigned long long tiles = (a.n + kDsTile - 1) / kDsTile;
      int per_sm = static_cast<int>((200u * 1024u) / (smem + 6 * 1024));
      per_sm = per_sm < 1 ? 1 : (per_sm > 5 ? 5 : per_sm);
      unsigned long long grid = static_cast<unsigned long long>(sm_count) * per_sm;
      if (grid > tiles) grid = tiles;
      cudaError_t e = cudaFuncSetAttribute(density_sorted_kernel<DENS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return e;
      density_sorted_kernel<DENS><<<static_cast<unsigned>(grid), kDensBlock, smem, st>>>(a);
      *grid_out = static_cast<int>(grid);
      return cudaGetLastError();
    }
  }
  const int grid = density_grid(a.n, sm_count);
  *grid_out = grid;
  if (real_type == MB200_REAL_F32) density_kernel<DENS, float><<<grid, kDensBlock, 0, st>>>(a);
  else density_kernel<DENS, double><<<grid, kDensBlock, 0, st>>>(a);
  return cudaGetLastError();
}

// Launches the density kernel; *grid_out = number of per-block partials written
// (a.partial must hold density_max_partials(sm_count) doubles).
cudaError_t launch_density(int dens, int real_type, const DensityArgs& a, int sm_count, int* grid_out, cudaStream_t st) {
  switch (dens) {
#define C(id) case id: return launch_density_t<id>(real_type, a, sm_count, grid_out, st)
    C(MB200_DENS_BINOMIAL); C(MB200_DENS_NORMAL); C(MB200_DENS_NEGATIVE_BINOMIAL_MU);
    C(MB200_DENS_NEGATIVE_BINOMIAL_PROB); C(MB200_DENS_BETA_BINOMIAL_PROB);
    C(MB200_DENS_BETA_BINOMIAL_AB); C(MB200_DENS_POISSON); C(MB200_DENS_UNIFORM);
    C(MB200_DENS_EXPONENTIAL_RATE); C(MB200_DENS_EXPONENTIAL_MEAN); C(MB200_DENS_GAMMA_RATE);
    C(MB200_DENS_GAMMA_SCALE); C(MB200_DENS_BETA); C(MB200_DENS_HYPERGEOMETRIC);
    C(MB200_DENS_LOG_NORMAL); C(MB200_DENS_TRUNCATED_NORMAL); C(MB200_DENS_CAUCHY);
    C(MB200_DENS_WEIBULL); C(MB200_DENS_ZI_POISSON); C(MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU);
    C(MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB);
#undef C
  default: return cudaErrorInvalidValue;
  }
}

int density_max_partials(int sm_count) { return sm_count * 8; }

cudaError_t install_lgamma_table_density(const double* d, const float* f) { return lgt_install(d, f); }

cudaError_t launch_reduce_partials(const double* partial, int n, double* sum, cudaStream_t st) {
  reduce_partials_kernel<<<1, 256, 0, st>>>(partial, n, sum);
  return cudaGetLastError();
}

}  // namespace mb200
