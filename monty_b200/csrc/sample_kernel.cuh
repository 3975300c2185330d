// sample_kernel.cuh -- the sampling kernels: one thread per stream, generator
// state in registers for the whole batch of draws, per-warp shared-memory
// staging so that the stream-major output y[n_samples * stream + sample]
// (ref: src/random.cpp:220-224) is written with fully coalesced stores.
//
// Layout in HBM
//   state  uint64[n_streams][4]    32 B per stream, 2 x 128-bit load/store
//   par[k] double[n_streams] or double[1] (broadcast)
//   out    OutT[n_streams][n_samples]  (stream-major == R's column-major
//          [n_samples x n_streams])
//
// Work decomposition.  A warp owns 32 consecutive streams.  Lane l draws for
// stream 32w + l and parks its draws in row l of the warp's shared-memory
// tile [32][TILE+1]; after TILE draws the warp turns the tile round: in step
// r all 32 lanes write the TILE consecutive values of stream 32w + r, i.e.
// one contiguous TILE * sizeof(OutT) run per store instruction.  The +1
// padding makes both the column writes and the row reads bank-conflict free.
// Warps never synchronise with each other; the grid is sized as a multiple of
// the SM count and strides over warp-tiles.
#pragma once
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
constexpr int kTile = 32;             // draws staged per lane before a write-out
constexpr int kPitch = kTile + 1;

struct __align__(16) TableSmem {
  double2 zig_xy[256];
  double zig_x[264];
  double tail_d[16];
  float tail_f[128];
};

__device__ __forceinline__ void report_error(ErrRec* err, unsigned long long stream, int code,
                                             const double* vals) {
  const unsigned long long old = atomicMin(&err->stream, stream);
  if (stream < old) {
    err->code = code;
    if (vals) { err->vals[0] = vals[0]; err->vals[1] = vals[1]; err->vals[2] = vals[2]; }
    __threadfence();
    err->owner = stream;
  }
}

template <typename OutT> __device__ __forceinline__ OutT out_nan();
template <> __device__ __forceinline__ double out_nan<double>() { return __longlong_as_double(0x7ff8000000000000ll); }
template <> __device__ __forceinline__ float out_nan<float>() { return __int_as_float(0x7fc00000); }

template <class D, int TABLES>
__device__ __forceinline__ void stage_tables(TableSmem* ts, Ctx& ctx) {
  if (TABLES == kZigTables) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ts->zig_xy[i] = g_zig_xy[i];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) ts->zig_x[i] = g_zig_x[i];
    ctx.zig_xy = ts->zig_xy;
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

// T: real_type of the computation; OutT: element type of `out`.
template <class D, typename T, typename OutT>
__global__ void __launch_bounds__(kBlock)
sample_kernel(const SampleArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  OutT* tile_all = reinterpret_cast<OutT*>(smem_raw);
  TableSmem* ts = reinterpret_cast<TableSmem*>(smem_raw + sizeof(OutT) * kWarps * 32 * kPitch);

  Ctx ctx;
  ctx.err = 0;
  stage_tables<D, D::TABLES>(ts, ctx);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  OutT* tile = tile_all + warp * 32 * kPitch;
  OutT* out = static_cast<OutT*>(a.out);

  const unsigned long long n_streams = a.n_streams;
  const unsigned long long n_samples = a.n_samples;
  const unsigned long long n_warp_tiles = (n_streams + 31) / 32;
  const unsigned long long warp_stride = static_cast<unsigned long long>(gridDim.x) * kWarps;

  for (unsigned long long wt = static_cast<unsigned long long>(blockIdx.x) * kWarps + warp;
       wt < n_warp_tiles; wt += warp_stride) {
    const unsigned long long s_first = wt * 32;
    const unsigned long long stream = s_first + lane;
    bool live = stream < n_streams;

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
      if (stream < n_streams) {
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
      const int n_rows = (n_streams - s_first) < 32 ? static_cast<int>(n_streams - s_first) : 32;
      for (unsigned long long j0 = 0; j0 < n_samples; j0 += kTile) {
        const int len = (n_samples - j0) < kTile ? static_cast<int>(n_samples - j0) : kTile;
        OutT* row = tile + lane * kPitch;
        if (live) {
          for (int j = 0; j < len; ++j) {
            row[j] = static_cast<OutT>(D::draw(rng, prep, ctx));
            if (D::DYN_ERR && ctx.err) {
              report_error(a.err, stream, ctx.err, ctx.err_vals);
              ctx.err = 0;
              live = false;
              for (int jj = j; jj < len; ++jj) row[jj] = out_nan<OutT>();
              break;
            }
          }
        } else {
          for (int j = 0; j < len; ++j) row[j] = out_nan<OutT>();
        }
        __syncwarp();
        // write-out: step r stores stream (s_first + r)'s `len` consecutive draws
        OutT* dst = out + s_first * n_samples + j0 + lane;
        const OutT* src = tile + lane;
        if (lane < len) {
#pragma unroll 4
          for (int r = 0; r < n_rows; ++r) {
            dst[r * n_samples] = src[r * kPitch];
          }
        }
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
  ctx.zig_xy = g_zig_xy; ctx.zig_x = g_zig_x; ctx.tail_d = g_tail_d; ctx.tail_f = g_tail_f;
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
  size_t b = sizeof(OutT) * kWarps * 32 * kPitch;
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
