// sample_kernel.cuh -- the sampling kernels: one thread per stream, generator
// state in registers for the whole batch of draws, per-warp shared-memory
// staging so that the stream-major output y[n_samples * stream + sample]
// (ref: src/random.cpp:220-224) is written with fully coalesced, line-aligned
// 128-bit stores.
//
// Layout in HBM
//   state  uint64[n_streams][4]    32 B per stream, 2 x 128-bit load/store
//   par[k] double[n_streams] or double[1] (broadcast)
//   out    OutT[n_streams][n_samples]  (stream-major == R's column-major
//          [n_samples x n_streams])
//
// Work decomposition.  A warp owns 32 consecutive streams; lane l draws for
// stream 32w + l.  Think of `out` as one long array indexed by the global
// element g = stream * n_samples + sample.  Row r's draws occupy
// [B_r, B_r + n_samples), B_r = (32w + r) * n_samples.  The warp walks each
// row in WINDOWS of kTile elements aligned in g (window k of row r covers
// [A_r + kTile*k, A_r + kTile*(k+1)), A_r = B_r rounded down to kTile), so a
// window is a whole number of 128-byte lines whatever n_samples is.  (With
// windows aligned in `sample` instead, n_samples = 1000 put every other row
// 64 B off the line grid, and the half-written lines left L2 before their
// other half arrived: 4.6 TB/s instead of 6.0 TB/s on a pure-store
// microbenchmark, profiles/r01_write_pattern.txt.)
//
// Per window: lane l computes the draws of ITS row that fall into the window
// and parks them in row l of the warp's shared-memory tile [32][PITCH]; then
// the warp turns the tile round: each store instruction writes whole rows
// (2 rows x 256 B for double) with 16 bytes per lane.  The row pitch is an odd
// number of 16-byte units, so both the column-wise STS.128 and the row-wise
// LDS.128 are bank-conflict free.  First / last windows of a row are ragged
// and take a masked scalar path.  Warps never synchronise with each other;
// the grid is a multiple of the SM count and strides over warp-tiles.
#pragma once
#include <type_traits>

#include "samplers.cuh"
#include "launch_args.h"

namespace mb200 {

// lookup tables in global memory, one private copy per translation unit
// (7 KB); kernels stage what they need into shared memory
#include "monty_tables.inc"
static __device__ const double g_zig_x[257] = { MONTY_ZIG_X_VALUES };
static __device__ const double2 g_zig_xy[256] = { MONTY_ZIG_XY_VALUES };
static __device__ const double g_tail_d[10] = { MONTY_TAIL_D_VALUES };
static __device__ const float g_tail_f[128] = { MONTY_TAIL_F_VALUES };

constexpr int kBlock = 128;           // 4 warps
constexpr int kWarps = kBlock / 32;
constexpr int kTile = 32;             // window width in elements

template <typename OutT> struct TileCfg;
template <> struct TileCfg<double> { static constexpr int VEC = 2, PITCH = 34; };
template <> struct TileCfg<float> { static constexpr int VEC = 4, PITCH = 36; };

struct __align__(16) TableSmem {
  double2 zig_xy[256];
  double zig_x[264];
  double tail_d[16];
  float tail_f[128];
  float zig_x2s[260];
};

__device__ __forceinline__ void report_error(ErrRec* err, unsigned long long stream, int code,
                                             const double* vals) {
  const unsigned long long old = atomicMin(&err->stream, stream);
  if (stream < old) {
    // The smallest stream wins the index atomically; its code and values must
    // describe the SAME stream, so they are written under a lock and only by a
    // thread that still is the minimum (two failing streams used to be able to
    // leave one's index with the other's values).  Error path only.
    bool done = false;
    while (!done) {
      if (atomicCAS(&err->lock, ~0u, 0u) == ~0u) {
        if (*reinterpret_cast<volatile unsigned long long*>(&err->stream) == stream) {
          err->code = code;
          if (vals) { err->vals[0] = vals[0]; err->vals[1] = vals[1]; err->vals[2] = vals[2]; }
          err->owner = stream;
        }
        __threadfence();
        atomicExch(&err->lock, ~0u);
        done = true;
      }
    }
  }
}

template <typename OutT> __device__ __forceinline__ OutT out_nan();
template <> __device__ __forceinline__ double out_nan<double>() { return __longlong_as_double(0x7ff8000000000000ll); }
template <> __device__ __forceinline__ float out_nan<float>() { return __int_as_float(0x7fc00000); }

template <int TABLES>
__device__ __forceinline__ void stage_tables(TableSmem* ts, Ctx& ctx) {
  if (TABLES == kZigTables) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ts->zig_xy[i] = g_zig_xy[i];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) ts->zig_x[i] = g_zig_x[i];
    ctx.zig_xy = ts->zig_xy;
    for (int i = threadIdx.x; i < 257; i += blockDim.x) {
      ts->zig_x2s[i] = static_cast<float>(g_zig_x[i] * g_zig_x[i] * 0.72134752044448170368);
    }
    ctx.zig_x2s = ts->zig_x2s;
    ctx.zig_xy_saddr = smem_addr(ts->zig_xy);
    ctx.zig_x = ts->zig_x;
  }
  if (TABLES == kTailTables) {
    for (int i = threadIdx.x; i < 10; i += blockDim.x) ts->tail_d[i] = g_tail_d[i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) ts->tail_f[i] = g_tail_f[i];
    ctx.tail_d = ts->tail_d;
    ctx.tail_f = ts->tail_f;
  }
  if (TABLES != kNoTables) __syncthreads();
}

__device__ __forceinline__ double2 lds_vec(uint32_t addr, double) { return lds_f64x2(addr); }
__device__ __forceinline__ float4 lds_vec(uint32_t addr, float) { return lds_f32x4(addr); }

// One warp's description of its 32 rows.
struct WarpRows {
  unsigned long long s_first;     // first stream of the warp
  unsigned long long n_samples;
  int n_rows;                     // rows that exist (n_streams may end inside the warp)
};

// g-aligned start of row r's window 0
__device__ __forceinline__ unsigned long long row_window_base(const WarpRows& w, int r) {
  return ((w.s_first + r) * w.n_samples) & ~static_cast<unsigned long long>(kTile - 1);
}

// Vector write-out of a window in which every element of every row is valid.
// All loads first, then all stores: the LDS latency is paid once per window
// and no register is rewritten while a store still owns it.
template <typename OutT>
__device__ __forceinline__ void write_window_full(uint32_t tile_saddr, OutT* out, const WarpRows& w,
                                                  unsigned long long k, int lane) {
  constexpr int VEC = TileCfg<OutT>::VEC;
  constexpr int PITCH = TileCfg<OutT>::PITCH;
  constexpr int LPR = kTile / VEC;         // lanes per row: 16 (double), 8 (float)
  constexpr int RPI = 32 / LPR;            // rows per instruction: 2 or 4
  constexpr int STEPS = 32 / RPI;          // 16 or 8
  using V = typename std::conditional<sizeof(OutT) == 8, double2, float4>::type;
  const int sub = (lane % LPR) * VEC;
  const int rsub = lane / LPR;
  // two half-batches of 8 x 16 B: all loads of a batch first, then its stores
  constexpr int HALF = STEPS / 2;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    V v[HALF];
#pragma unroll
    for (int t = 0; t < HALF; ++t) {
      const int r = rsub + (h * HALF + t) * RPI;
      v[t] = lds_vec(tile_saddr + static_cast<uint32_t>((r * PITCH + sub) * sizeof(OutT)), OutT());
    }
#pragma unroll
    for (int t = 0; t < HALF; ++t) {
      const int r = rsub + (h * HALF + t) * RPI;
      OutT* dst = out + row_window_base(w, r) + k * kTile + sub;
      *reinterpret_cast<V*>(dst) = v[t];
    }
  }
}

// The same with the 16-byte destinations of this lane precomputed per warp-tile
// as 32-bit byte offsets from row 0's window 0 (roff, see sample_kernel): the
// 64-bit multiply-and-mask per row and window of the variant above is ~150
// instructions per window, a quarter of what a window of uniform draws costs.
template <typename OutT, int STEPS>
__device__ __forceinline__ void write_window_full_roff(uint32_t tile_saddr, char* wwin, const uint32_t (&roff)[STEPS], int lane) {
  constexpr int VEC = TileCfg<OutT>::VEC;
  constexpr int PITCH = TileCfg<OutT>::PITCH;
  constexpr int LPR = kTile / VEC;
  constexpr int RPI = 32 / LPR;
  static_assert(STEPS == 32 / RPI, "one offset per store instruction");
  using V = typename std::conditional<sizeof(OutT) == 8, double2, float4>::type;
  const int sub = (lane % LPR) * VEC;
  const int rsub = lane / LPR;
  constexpr int HALF = STEPS / 2;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    V v[HALF];
#pragma unroll
    for (int t = 0; t < HALF; ++t) {
      const int r = rsub + (h * HALF + t) * RPI;
      v[t] = lds_vec(tile_saddr + static_cast<uint32_t>((r * PITCH + sub) * sizeof(OutT)), OutT());
    }
#pragma unroll
    for (int t = 0; t < HALF; ++t) *reinterpret_cast<V*>(wwin + roff[h * HALF + t]) = v[t];
  }
}

// Masked write-out for ragged windows (first / last of a row, short rows,
// missing rows): lane e of step r writes element e of row r if it belongs to
// the row.
template <typename OutT>
__device__ __forceinline__ void write_window_edge(const OutT* tile, OutT* out, const WarpRows& w,
                                                  unsigned long long k, int lane) {
  constexpr int PITCH = TileCfg<OutT>::PITCH;
  for (int r = 0; r < w.n_rows; ++r) {
    const unsigned long long b = (w.s_first + r) * w.n_samples;
    const unsigned long long g = row_window_base(w, r) + k * kTile + lane;
    if (g >= b && g < b + w.n_samples) out[g] = tile[r * PITCH + lane];
  }
}

// D::draw_batch where the sampler has one (the call must not be instantiated otherwise)
template <class D, typename T, int N>
__device__ __forceinline__ void batch_draw(Rng& rng, const typename D::Prep& prep, Ctx& ctx, T (&b)[N]) {
  if constexpr (D::BATCH > 1) D::draw_batch(rng, prep, ctx, b);
}

// T: real_type of the computation; OutT: element type of `out`.
template <class D, typename T, typename OutT>
__global__ void __launch_bounds__(kBlock, D::MINBLOCKS)
sample_kernel(const SampleArgs a) {
  constexpr int VEC = TileCfg<OutT>::VEC;
  constexpr int PITCH = TileCfg<OutT>::PITCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  OutT* tile_all = reinterpret_cast<OutT*>(smem_raw);
  TableSmem* ts = reinterpret_cast<TableSmem*>(smem_raw + sizeof(OutT) * kWarps * 32 * PITCH);

  Ctx ctx;
  ctx.err = 0;
  stage_tables<D::TABLES>(ts, ctx);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  OutT* tile = tile_all + warp * 32 * PITCH;
  OutT* out = static_cast<OutT*>(a.out);

  const unsigned long long n_streams = a.n_streams;
  const unsigned long long n_samples = a.n_samples;
  const unsigned long long n_warp_tiles = (n_streams + 31) / 32;
  const unsigned long long warp_stride = static_cast<unsigned long long>(gridDim.x) * kWarps;
  // 128-bit stores need every row to start on a 16-byte boundary
  const bool vec_ok = (n_samples % VEC == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);

  for (unsigned long long wt = static_cast<unsigned long long>(blockIdx.x) * kWarps + warp;
       wt < n_warp_tiles; wt += warp_stride) {
    WarpRows w;
    w.s_first = wt * 32;
    w.n_samples = n_samples;
    w.n_rows = (n_streams - w.s_first) < 32 ? static_cast<int>(n_streams - w.s_first) : 32;
    const unsigned long long stream = w.s_first + lane;
    const bool exists = stream < n_streams;
    bool live = exists;

    Rng rng;
    typename D::Prep prep;
    if (live) {
      rng.s = load_state(a.state, stream);
      T par[D::NP > 0 ? D::NP : 1];
#pragma unroll
      for (int k = 0; k < D::NP; ++k) {
        par[k] = static_cast<T>(a.par[k][a.pstride[k] ? stream : 0]);
      }
      const int e = D::prepare(par, prep, ctx);
      if (e != kOk) {
        report_error(a.err, stream, e, nullptr);
        live = false;
      }
    }
    const bool touched = live;

    if (n_samples == 1) {
      // one draw per stream (monty_random_<dist>, the SIR-step shape):
      // consecutive lanes write consecutive elements, no staging needed
      if (exists) {
        OutT y = out_nan<OutT>();
        if (live) {
          y = static_cast<OutT>(D::draw(rng, prep, ctx));
          if (D::DYN_ERR && ctx.err) {
            report_error(a.err, stream, ctx.err, ctx.err_vals);
            ctx.err = 0;
            y = out_nan<OutT>();
          }
        }
        out[stream] = y;
      }
    } else {
      OutT* row = tile + lane * PITCH;
      const uint32_t tile_saddr = smem_addr(tile);
      const uint32_t row_saddr = tile_saddr + static_cast<uint32_t>(lane * PITCH * sizeof(OutT));
      // simple samplers (UNROLL != 0 and no attempt loop: registers to spare):
      // byte offsets of this lane's write-out destinations, relative to the
      // window-0 base of the warp-tile's first row, held in registers
      constexpr int kSteps = 32 / (32 / (kTile / VEC));
      constexpr bool kRoff = D::UNROLL != 0 && !D::DEFER;
      const bool roff_ok = kRoff && n_samples < (1ull << 24);
      uint32_t roff[kSteps];
      char* wbase0 = nullptr;
      if (kRoff && roff_ok) {
        const uint32_t nn = static_cast<uint32_t>(n_samples), nm = nn & 31u;
        const uint32_t sm0 = static_cast<uint32_t>(w.s_first * n_samples) & 31u;    // m of row 0
        wbase0 = reinterpret_cast<char*>(out + ((w.s_first * n_samples) & ~static_cast<unsigned long long>(kTile - 1)));
        constexpr int LPR = kTile / VEC, RPI = 32 / LPR;
#pragma unroll
        for (int t = 0; t < kSteps; ++t) {
          const uint32_t r = static_cast<uint32_t>(lane / LPR + t * RPI);
          // row r starts r * n elements after row 0; its window 0 starts at that minus its own m
          const uint32_t m_r = (sm0 + r * nm) & 31u;
          roff[t] = (sm0 + r * nn - m_r + static_cast<uint32_t>((lane % LPR) * VEC)) * static_cast<uint32_t>(sizeof(OutT));
          asm volatile("" : "+r"(roff[t]));     // keep them in registers across the windows
        }
      }
      // this lane's row: sample j sits in slot (m + j) % kTile of window
      // (m + j) / kTile, m = offset of the row start inside its window 0
      const int m = static_cast<int>((stream * n_samples) & (kTile - 1));
      // the warp walks as many windows as a row can span
      const unsigned long long n_windows = (n_samples + 2 * (kTile - 1)) / kTile;
      for (unsigned long long k = 0; k < n_windows; ++k) {
        // slots [e_lo, e_hi) of window k belong to this lane's row
        const long long j_first = static_cast<long long>(k * kTile) - m;   // sample index of slot 0
        const int e_lo = j_first < 0 ? static_cast<int>(-j_first) : 0;
        const long long rem = static_cast<long long>(n_samples) - j_first;
        const int e_hi = !exists ? 0 : (rem >= kTile ? kTile : (rem > 0 ? static_cast<int>(rem) : 0));
        if (!__any_sync(0xffffffffu, e_hi > e_lo)) continue;   // nobody has data in this window
        const bool row_full = e_lo == 0 && e_hi == kTile;
        const bool all_full = vec_ok && __all_sync(0xffffffffu, row_full);
        if constexpr (D::DEFER) {
          // Rejection samplers in resumable pieces (DistBase::DEFER): every lane
          // fills ITS slots of the window at its own pace -- one attempt per
          // trip of this loop -- instead of the warp waiting, draw by draw, for
          // the lane with the most rejections; candidates that need the
          // expensive accept test are parked and tested together (see
          // sorted_kernel.cuh, which runs the same loop over sorted streams).
          int e = e_lo;
          bool parked = false;
          typename D::Cand cand;
          for (;;) {
            const bool run = live && e < e_hi && !parked;
            const unsigned m_run = __ballot_sync(0xffffffffu, run);
            unsigned m_park = __ballot_sync(0xffffffffu, parked);
            if ((m_run | m_park) == 0) break;
            if (run) {
              T value;
              const int st = D::attempt(rng, prep, ctx, cand, value);
              if (D::DYN_ERR && ctx.err) {
                report_error(a.err, stream, ctx.err, ctx.err_vals);
                ctx.err = 0;
                live = false;
              } else if (st == 1) {
                row[e++] = static_cast<OutT>(value);
              } else if (st == 2) {
                parked = true;
              }
            }
            m_park = __ballot_sync(0xffffffffu, parked);
            const unsigned m_next = __ballot_sync(0xffffffffu, live && e < e_hi && !parked);
            if (m_park != 0 && (__popc(m_park) >= 16 || m_next == 0)) {
              if (parked) {
                T value;
                if (D::slow(prep, ctx, cand, value)) row[e++] = static_cast<OutT>(value);
                parked = false;
              }
            }
          }
          if (!live) {
            for (; e < e_hi; ++e) row[e] = out_nan<OutT>();
          }
        } else if (live) {
          if (all_full && D::UNROLL == 2) {
            // tiny draw(): the whole window straight-line, VEC draws per store
#pragma unroll
            for (int j = 0; j < kTile; j += VEC) {
              OutT v[VEC];
#pragma unroll
              for (int e = 0; e < VEC; ++e) v[e] = static_cast<OutT>(D::draw(rng, prep, ctx));
              sts_vec(row_saddr + static_cast<uint32_t>(j * sizeof(OutT)), v);
            }
          } else if (all_full && D::UNROLL == 1 && D::BATCH > 1) {
            // BATCH draws per trip (DistBase::BATCH), then BATCH / VEC vector stores
            static_assert(D::BATCH == 1 || (D::BATCH % VEC == 0 && kTile % D::BATCH == 0), "batch shape");
#pragma unroll 1
            for (int j = 0; j < kTile; j += D::BATCH) {
              T b[D::BATCH > 1 ? D::BATCH : 1];
              batch_draw<D, T>(rng, prep, ctx, b);
#pragma unroll
              for (int g = 0; g < D::BATCH; g += VEC) {
                OutT v[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = static_cast<OutT>(b[g + e]);
                sts_vec(row_saddr + static_cast<uint32_t>((j + g) * sizeof(OutT)), v);
              }
            }
          } else if (all_full && D::UNROLL == 1) {
#pragma unroll 1
            for (int j = 0; j < kTile; j += VEC) {
              OutT v[VEC];
#pragma unroll
              for (int e = 0; e < VEC; ++e) v[e] = static_cast<OutT>(D::draw(rng, prep, ctx));
              sts_vec(row_saddr + static_cast<uint32_t>(j * sizeof(OutT)), v);
            }
          } else {
#pragma unroll 1
            for (int e = e_lo; e < e_hi; ++e) row[e] = static_cast<OutT>(D::draw(rng, prep, ctx));
          }
          if (D::DYN_ERR && ctx.err) {
            // a data-dependent failure inside a composition: the stream is
            // abandoned and the rest of its output is NaN
            report_error(a.err, stream, ctx.err, ctx.err_vals);
            ctx.err = 0;
            live = false;
          }
        }
        if (!D::DEFER && !live) {
          for (int e = e_lo; e < e_hi; ++e) row[e] = out_nan<OutT>();
        }
        __syncwarp();
        if (all_full && kRoff && roff_ok) {
          write_window_full_roff<OutT, kSteps>(tile_saddr, wbase0 + k * (kTile * sizeof(OutT)), roff, lane);
        } else if (all_full) write_window_full<OutT>(tile_saddr, out, w, k, lane);
        else write_window_edge<OutT>(tile, out, w, k, lane);
        __syncwarp();
      }
    }
    if (touched) store_state(a.state, stream, rng.s);
  }
}

// validate-only pass: finds the first stream whose parameters the reference
// would reject, without touching any state (used by the host API to reproduce
// "streams before the offender are advanced, later ones are not").
template <class D, typename T>
__global__ void __launch_bounds__(kBlock)
validate_kernel(const SampleArgs a) {
  Ctx ctx;
  ctx.err = 0;
  ctx.zig_xy = g_zig_xy; ctx.zig_xy_saddr = 0; ctx.zig_x = g_zig_x; ctx.zig_x2s = nullptr; ctx.tail_d = g_tail_d; ctx.tail_f = g_tail_f;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long stream = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
       stream < a.n_streams; stream += stride) {
    T par[D::NP > 0 ? D::NP : 1];
#pragma unroll
    for (int k = 0; k < D::NP; ++k) {
      par[k] = static_cast<T>(a.par[k][a.pstride[k] ? stream : 0]);
    }
    typename D::Prep prep;
    const int e = D::prepare(par, prep, ctx);
    if (e != kOk) report_error(a.err, stream, e, nullptr);
  }
}

template <typename OutT>
inline size_t sample_smem_bytes(int tables) {
  size_t b = sizeof(OutT) * kWarps * 32 * TileCfg<OutT>::PITCH;
  b = (b + 15) & ~static_cast<size_t>(15);
  if (tables != kNoTables) b += sizeof(TableSmem);
  return b;
}

inline unsigned sample_grid(unsigned long long n_streams, int sm_count) {
  const unsigned long long blocks_needed = (n_streams + kBlock - 1) / kBlock;
  const unsigned long long cap = static_cast<unsigned long long>(sm_count) * 16;  // a multiple of the SM count
  return static_cast<unsigned>(blocks_needed < cap ? (blocks_needed ? blocks_needed : 1) : cap);
}

template <class D, typename T, typename OutT>
cudaError_t launch_sample_t(const SampleArgs& a, int sm_count, cudaStream_t st) {
  const size_t smem = sample_smem_bytes<OutT>(D::TABLES);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sample_kernel<D, T, OutT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  sample_kernel<D, T, OutT><<<sample_grid(a.n_streams, sm_count), kBlock, smem, st>>>(a);
  return cudaGetLastError();
}

template <class D, typename T>
cudaError_t launch_validate_t(const SampleArgs& a, int sm_count, cudaStream_t st) {
  validate_kernel<D, T><<<sample_grid(a.n_streams, sm_count), kBlock, 0, st>>>(a);
  return cudaGetLastError();
}

// Instantiates the three (real_type, out type) combinations of one sampler
// and the two validators.
template <template <typename> class D>
cudaError_t launch_sample_any(int real_type, int out_f32, bool validate_only,
                              const SampleArgs& a, int sm_count, cudaStream_t st) {
  if (validate_only) {
    return real_type == 1 ? launch_validate_t<D<float>, float>(a, sm_count, st)
                          : launch_validate_t<D<double>, double>(a, sm_count, st);
  }
  if (real_type == 1) {
    return out_f32 ? launch_sample_t<D<float>, float, float>(a, sm_count, st)
                   : launch_sample_t<D<float>, float, double>(a, sm_count, st);
  }
  return launch_sample_t<D<double>, double, double>(a, sm_count, st);
}

}  // namespace mb200
