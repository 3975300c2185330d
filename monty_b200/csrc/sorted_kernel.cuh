// sorted_kernel.cuh -- samplers whose cost depends on the REGIME their
// per-stream parameters select (binomial: constant / inversion / BTRS;
// Poisson: Knuth / Hormann / Cauchy; gamma: small / exponential / large;
// hypergeometric: HIN / H2PE; ...), for short batches (n_samples <= 64: the
// dust2 SIR-step shape, BASELINE configs[2] and [3]).
//
// In the generic kernel a warp holds 32 consecutive streams, so with
// per-stream parameters spanning the regimes every warp executes every regime
// one after the other: ncu shows 4-10 active lanes per instruction
// (profiles/r01_binomial_d1_ncu_summary.txt, r01_hypergeometric_...).  Here a
// CTA takes a tile of 2048 consecutive streams and
//   1. classifies each stream from its parameters alone (D::regime(), a few
//      compares -- none of the set-up work),
//   2. counting-sorts the tile by regime in shared memory (ballot-free:
//      shared atomics hand out ranks; the order inside a regime is irrelevant
//      because a stream's draws depend only on its own state and parameters),
//   3. lets thread t process the t-th stream OF THE SORTED ORDER: warps are
//      regime-pure except at the <= R-1 boundaries of a tile, so prepare() and
//      draw() run without regime divergence (rejection-loop trip counts still
//      vary inside a warp),
//   4. writes each draw straight to its place out[stream * n + d]: these
//      kernels are compute-bound by an order of magnitude, and not staging
//      the output keeps shared memory for a 2048-stream tile (64 chunks of 32
//      for 8 warps: small tail at the tile barrier, few mixed-regime chunks).
// Generator states are read and written in place (one 32-byte sector per
// stream, whichever thread handles it).  Results are bit-identical to the
// generic kernel's: only the assignment of streams to lanes changes.
#pragma once
#include <cstdlib>

#include "sample_kernel.cuh"

namespace mb200 {

// resident CTAs per SM this kernel is compiled for: D::SORT_MINBLOCKS (samplers.cuh), 3 by
// default -- 24 warps per SM at 80 registers: these kernels wait on dependent FP64 chains,
// occupancy beats the few spills (2: binomial x16 29.3, 3: 33.9, 4: 33.9 but Poisson x16
// 45.3 -> 34.9 and hypergeometric 16.4 -> 15.0 Gdraws/s); the negative binomial and beta run
// at 4 (15.4 -> 16.5, 24.2 -> 24.6), the truncated normal at 2 (29.1 -> 29.8).
// -DMB200_SORT_MINBLOCKS=n overrides all of them (experiments).
constexpr int kSortBlock = 256;
constexpr int kSortTile = 2048;           // streams per tile, kSortTile / kSortBlock per thread
constexpr int kSortMaxSamples = 64;
constexpr int kSortKeys = 16;             // regime keys 0..14, 15 = stream does not exist
// free lanes at which a warp of the persistent variant takes new streams (measured 8 / 16 / 24:
// binomial x 1 14.6 / 14.5 / 14.4, Poisson x 1 20.4 / 20.4 / 20.1, truncated normal x 16 38.4 / 37.8 / 36.1)
#ifndef MB200_SORT_REFILL
#define MB200_SORT_REFILL 8
#endif
constexpr int kSortRefill = MB200_SORT_REFILL;

inline size_t sorted_smem_bytes(int n_params, int tables) {
  size_t b = sizeof(double) * kSortTile * (n_params > 0 ? n_params : 1);
  if (tables != kNoTables) b += sizeof(TableSmem);
  return b;
}

// PERSIST (DEFER samplers only): lanes keep working across 32-stream chunk boundaries, see
// step 3.  Chosen per launch (launch_sorted_t): it wins where streams finish at very
// different times (one draw per stream; the truncated normal's tail rejections) and loses
// ~10% on long regular batches, where each refill stalls the warp on the new streams' loads.
template <class D, typename T, typename OutT, bool PERSIST>
#ifdef MB200_SORT_MINBLOCKS
__global__ void __launch_bounds__(kSortBlock, MB200_SORT_MINBLOCKS)
#else
__global__ void __launch_bounds__(kSortBlock, PERSIST ? 3 : D::SORT_MINBLOCKS)
#endif
sorted_sample_kernel(const SampleArgs a) {
  constexpr int NP = D::NP;
  constexpr int PER = kSortTile / kSortBlock;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned s_count[kSortKeys];
  __shared__ unsigned s_offset[kSortKeys];
  __shared__ unsigned short s_order[kSortTile];
  __shared__ unsigned s_chunk;                       // next 32-stream chunk of the sorted order

  const unsigned n = static_cast<unsigned>(a.n_samples);
  double* par_s = reinterpret_cast<double*>(smem_raw);
  TableSmem* ts = reinterpret_cast<TableSmem*>(smem_raw + sizeof(double) * kSortTile * (NP > 0 ? NP : 1));

  Ctx ctx;
  ctx.err = 0;
  stage_tables<D::TABLES>(ts, ctx);

  const unsigned tid = threadIdx.x;
  OutT* const out = static_cast<OutT*>(a.out);
  const unsigned long long n_streams = a.n_streams;
  const unsigned long long n_tiles = (n_streams + kSortTile - 1) / kSortTile;

  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long base = tile * kSortTile;
    const unsigned count = (n_streams - base) < kSortTile ? static_cast<unsigned>(n_streams - base) : kSortTile;
    if (tid < kSortKeys) s_count[tid] = 0;
    if (tid == 0) s_chunk = 0;
    __syncthreads();
    // 1. parameters -> shared memory, regime key and rank inside the key
    unsigned key[PER], rank[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const unsigned j = tid + i * kSortBlock;
      key[i] = kSortKeys - 1;
      if (j < count) {
        T par[NP > 0 ? NP : 1];
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const double v = a.par[k][a.pstride[k] ? base + j : 0];
          par_s[k * kSortTile + j] = v;
          par[k] = static_cast<T>(v);
        }
        const int r = D::regime(par);
        key[i] = static_cast<unsigned>(r < 0 ? 0 : (r > kSortKeys - 2 ? kSortKeys - 2 : r));
      }
      rank[i] = atomicAdd(&s_count[key[i]], 1u);
    }
    __syncthreads();
    // 2. exclusive scan of the 16 counts (one warp), then the sorted order
    if (tid < 32) {
      const unsigned c = tid < kSortKeys ? s_count[tid] : 0;
      unsigned incl = c;
#pragma unroll
      for (int d = 1; d < kSortKeys; d <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
        if (tid >= static_cast<unsigned>(d)) incl += up;
      }
      if (tid < kSortKeys) s_offset[tid] = incl - c;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      s_order[s_offset[key[i]] + rank[i]] = static_cast<unsigned short>(tid + i * kSortBlock);
    }
    __syncthreads();
    // 3. warps take 32-stream chunks of the sorted order from a shared counter,
    //    the expensive regimes (high keys, at the end of the order) first: the
    //    cost of a chunk varies by an order of magnitude between regimes, and
    //    with a static assignment the cheap warps would wait at the barrier
    //    below for the expensive ones
    if constexpr (PERSIST) {
      // Persistent lanes.  A lane keeps its stream until the stream's n draws are made; when
      // at least half of the warp's lanes are free, the free lanes take the next streams of
      // the order TOGETHER (their prepare() runs with >= 16 lanes) -- so a warp does not
      // wait for the slowest of 32 streams before it starts the next 32 (n = 1: a third of
      // the lanes needed a second or third attempt while the others idled).
      // Rejection loop with the expensive accept test deferred: a lane whose candidate
      // fails the cheap squeeze parks it; the parked candidates of the warp are tested
      // together once half the warp waits (or nobody can run), so the ~250-instruction
      // test executes with many lanes active.  Per stream nothing changes: same words,
      // same operations.
      const unsigned lane = tid & 31;
      const unsigned lt_mask = (1u << lane) - 1u;
      bool have = false, parked = false, more = true;
      unsigned d = 0;
      unsigned long long stream = 0;
      OutT* row = out;
      typename D::Prep prep;
      typename D::Cand cand;
      Rng rng;
#pragma unroll 1
      for (;;) {
        const unsigned m_have = __ballot_sync(0xffffffffu, have);
        if (more && __popc(m_have) <= 32 - kSortRefill) {
          const unsigned m_free = ~m_have;
          const unsigned want = static_cast<unsigned>(__popc(m_free));
          unsigned first = 0;
          if (lane == 0) first = atomicAdd(&s_chunk, want);
          first = __shfl_sync(0xffffffffu, first, 0);
          if (first + want >= count) more = false;
          const unsigned k = first + static_cast<unsigned>(__popc(m_free & lt_mask));
          if (!have && k < count) {
            const unsigned j = s_order[count - 1 - k];
            stream = base + j;
            row = out + stream * n;
            T par[NP > 0 ? NP : 1];
#pragma unroll
            for (int q = 0; q < NP; ++q) par[q] = static_cast<T>(par_s[q * kSortTile + j]);
            const int e = D::prepare(par, prep, ctx);
            if (e != kOk) {
              report_error(a.err, stream, e, nullptr);
              for (unsigned q = 0; q < n; ++q) row[q] = out_nan<OutT>();
            } else {
              rng.s = load_state(a.state, stream);
              have = true;
              parked = false;
              d = 0;
            }
          }
        }
        const bool run = have && !parked;
        const unsigned m_run = __ballot_sync(0xffffffffu, run);
        unsigned m_park = __ballot_sync(0xffffffffu, parked);
        if ((m_run | m_park) == 0) {
          if (!more) break;
          continue;                                   // nothing in hand, streams left: refill
        }
        if (run) {
          T value;
          const int st = D::attempt(rng, prep, ctx, cand, value);
          if (D::DYN_ERR && ctx.err) {
            // a data-dependent failure inside a composition: the stream is abandoned
            // and the rest of its output is NaN (as in sample_kernel)
            report_error(a.err, stream, ctx.err, ctx.err_vals);
            ctx.err = 0;
            for (; d < n; ++d) row[d] = out_nan<OutT>();
            store_state(a.state, stream, rng.s);
            have = false;
          } else if (st == 1) {
            row[d++] = static_cast<OutT>(value);
            if (d == n) { store_state(a.state, stream, rng.s); have = false; }
          } else if (st == 2) {
            parked = true;
          }
        }
        m_park = __ballot_sync(0xffffffffu, parked);
        const unsigned m_next = __ballot_sync(0xffffffffu, have && !parked);
        if (m_park != 0 && (__popc(m_park) >= 16 || m_next == 0)) {
          if (parked) {
            T value;
            if (D::slow(prep, ctx, cand, value)) {
              row[d++] = static_cast<OutT>(value);
              if (d == n) { store_state(a.state, stream, rng.s); have = false; }
            }
            parked = false;
          }
        }
      }
    } else {
    const unsigned n_chunks = (count + 31) / 32;
#pragma unroll 1
    for (;;) {
      unsigned chunk = 0;
      if ((tid & 31) == 0) chunk = atomicAdd(&s_chunk, 1u);
      chunk = __shfl_sync(0xffffffffu, chunk, 0);
      if (chunk >= n_chunks) break;
      const unsigned slot = (n_chunks - 1 - chunk) * 32 + (tid & 31);
      const bool exists = slot < count;              // the non-existent streams sort last
      const unsigned j = exists ? s_order[slot] : 0;
      const unsigned long long stream = base + j;
      T par[NP > 0 ? NP : 1];
#pragma unroll
      for (int k = 0; k < NP; ++k) par[k] = static_cast<T>(par_s[k * kSortTile + j]);
      typename D::Prep prep;
      OutT* row = out + stream * n;                  // this stream's draws, written as they come
      bool live = exists;
      if (exists) {
        const int e = D::prepare(par, prep, ctx);
        if (e != kOk) {
          report_error(a.err, stream, e, nullptr);
          for (unsigned d = 0; d < n; ++d) row[d] = out_nan<OutT>();
          live = false;
        }
      }
      Rng rng;
      if (live) rng.s = load_state(a.state, stream);
      const bool touched = live;
      if constexpr (D::DEFER) {
        // Rejection loop with the expensive accept test deferred: a lane whose
        // candidate fails the cheap squeeze parks it; the parked candidates of
        // the warp are tested together once half the warp waits (or nobody can
        // run), so the ~250-instruction test executes with many lanes active
        // instead of once per warp-iteration for the one or two lanes that
        // need it.  Per stream nothing changes: same words, same operations.
        unsigned d = 0;
        bool parked = false;
        typename D::Cand cand;
        for (;;) {
          const bool run = live && d < n && !parked;
          const unsigned m_run = __ballot_sync(0xffffffffu, run);
          unsigned m_park = __ballot_sync(0xffffffffu, parked);
          if ((m_run | m_park) == 0) break;
          if (run) {
            T value;
            const int st = D::attempt(rng, prep, ctx, cand, value);
            if (D::DYN_ERR && ctx.err) {
              // a data-dependent failure inside a composition: the stream is
              // abandoned and the rest of its output is NaN (as in sample_kernel)
              report_error(a.err, stream, ctx.err, ctx.err_vals);
              ctx.err = 0;
              live = false;
              for (; d < n; ++d) row[d] = out_nan<OutT>();
            } else if (st == 1) {
              row[d++] = static_cast<OutT>(value);
            } else if (st == 2) {
              parked = true;
            }
          }
          m_park = __ballot_sync(0xffffffffu, parked);
          const unsigned m_next = __ballot_sync(0xffffffffu, live && d < n && !parked);
          if (m_park != 0 && (__popc(m_park) >= 16 || m_next == 0)) {
            if (parked) {
              T value;
              if (D::slow(prep, ctx, cand, value)) row[d++] = static_cast<OutT>(value);
              parked = false;
            }
          }
        }
      } else {
#pragma unroll 1
        for (unsigned d = 0; d < n; ++d) {
          if (!exists) break;
          OutT y = out_nan<OutT>();
          if (live) {
            y = static_cast<OutT>(D::draw(rng, prep, ctx));
            if (D::DYN_ERR && ctx.err) {
              // a data-dependent failure inside a composition: the stream is
              // abandoned and the rest of its output is NaN (as in sample_kernel)
              report_error(a.err, stream, ctx.err, ctx.err_vals);
              ctx.err = 0;
              live = false;
              y = out_nan<OutT>();
            }
          }
          row[d] = y;
        }
      }
      if (touched) store_state(a.state, stream, rng.s);
    }
    }
    __syncthreads();                                 // par_s / s_order are reused by the next tile
  }
}

template <class D, typename T, typename OutT, bool PERSIST>
cudaError_t launch_sorted_v(const SampleArgs& a, int sm_count, cudaStream_t st) {
  const size_t smem = sorted_smem_bytes(D::NP, D::TABLES);
  cudaError_t e = cudaFuncSetAttribute(sorted_sample_kernel<D, T, OutT, PERSIST>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  const unsigned long long tiles = (a.n_streams + kSortTile - 1) / kSortTile;
  const unsigned long long cap = static_cast<unsigned long long>(sm_count) * 8;
  const unsigned grid = static_cast<unsigned>(tiles < cap ? (tiles ? tiles : 1) : cap);
  sorted_sample_kernel<D, T, OutT, PERSIST><<<grid, kSortBlock, smem, st>>>(a);
  return cudaGetLastError();
}
template <class D, typename T, typename OutT>
cudaError_t launch_sorted_t(const SampleArgs& a, int sm_count, cudaStream_t st) {
  if constexpr (D::DEFER) {
    // persistent lanes up to D::SORT_PERSIST_MAX_N draws per stream (samplers.cuh)
    if (a.n_samples <= static_cast<unsigned long long>(D::SORT_PERSIST_MAX_N)) return launch_sorted_v<D, T, OutT, true>(a, sm_count, st);
  }
  return launch_sorted_v<D, T, OutT, false>(a, sm_count, st);
}

// Is the sorted kernel the right one for this launch?  Only when the
// parameters vary per stream (with scalars every stream is in the same regime
// already), the batch is short enough to stage, and the tile is worth sorting.
template <class D>
inline bool sorted_kernel_applies(const SampleArgs& a) {
  // the samplers with deferred accept tests gain at any batch length (n = 1000,
  // per-stream parameters: Poisson 17.9 -> 27.1, binomial 15.5 -> 18.3 Gdraws/s);
  // for the others the generic kernel's coalesced write-out wins on long batches
  const unsigned long long max_samples = D::DEFER ? (1ull << 31) : kSortMaxSamples;
  if (D::REGIMES <= 1 || a.n_samples < 1 || a.n_samples > max_samples) return false;
  if (a.n_streams < 2048) return false;
  bool any_vector = false;
  for (int k = 0; k < D::NP; ++k) any_vector = any_vector || a.pstride[k] != 0;
  return any_vector;
}

// Front end used by the sampler groups: the sorted kernel where it applies
// (MB200_NO_SORT=1 keeps the generic kernel, for A/B measurements), else the
// generic one.
template <template <typename> class D>
cudaError_t launch_sample_regimes(int real_type, int out_f32, bool validate_only,
                                  const SampleArgs& a, int sm_count, cudaStream_t st) {
  if constexpr (D<double>::REGIMES > 1) {
    static const bool no_sort = std::getenv("MB200_NO_SORT") != nullptr;
    if (!validate_only && !no_sort && sorted_kernel_applies<D<double>>(a)) {
      if (real_type == 1) {
        return out_f32 ? launch_sorted_t<D<float>, float, float>(a, sm_count, st)
                       : launch_sorted_t<D<float>, float, double>(a, sm_count, st);
      }
      return launch_sorted_t<D<double>, double, double>(a, sm_count, st);
    }
  }
  return launch_sample_any<D>(real_type, out_f32, validate_only, a, sm_count, st);
}

}  // namespace mb200
