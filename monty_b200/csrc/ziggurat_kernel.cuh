// ziggurat_kernel.cuh -- monty_random_n_normal(algorithm = "ziggurat"), the
// headline workload (BASELINE configs[1]).
//
// ref: inst/include/monty/random/normal_ziggurat.hpp:70-114 (the draw),
//      :14-30 (tail), :61-65 (layer from bits 3..10 of the same word).
//
// Why not the generic sample_kernel.  98.5% of ziggurat draws need one word,
// one table look-up, one compare; 1.5% need the wedge test (one more word),
// 3e-4 the tail.  In lock-step a single wedge lane stalls its warp (38% of
// warp-iterations), the random {x, y} look-up is an 8-byte-per-bank-group
// gather that costs ~10 shared-memory wavefronts per warp-draw, and the budget
// at the HBM roofline is only 48 issue cycles and 12 shared-memory wavefronts
// per warp-draw (profiles/r01_ziggurat_*.txt).  This kernel is built around those
// three numbers:
//
//  * One word per lane per iteration, no data-dependent branch.  A lane is in
//    mode FAST (the word is a new attempt) or WEDGE (the word is the u1 of the
//    candidate it parked in the previous iteration).  Both modes execute the
//    same straight-line code; which results are kept is decided by predicates.
//    Each stream still consumes its words in exactly the reference's order.
//  * The wedge decision `f1 + u1 (f0 - f1) < 1` is only a boolean.  It is
//    taken in single precision in the log domain, lg2(f(x1) + u1 (f(xi) -
//    f(x1))) + z^2 / (2 ln 2) < 0, from a 23-bit image of the word (one IMAD.HI,
//    no conversion instruction); total error < 1e-5, and the reference's own
//    double expression decides inside the |t| <= 1e-4 band (rare branch).  The
//    accepted value never depends on the test, so draws are bit-identical.
//  * The {x, y} table is replicated 8 times, copy j in bank group j, lane l
//    reads copy l % 8: the 16-byte gather is conflict free (4 wavefronts).
//    32 KB of tables are affordable because ONE CTA of 20 warps owns an SM.
//  * Warps take 32-stream tiles from an atomic counter (11 tiles per warp at
//    2^20 streams: static striding would leave an 8% tail).
//
// Output staging (line-aligned windows, 128-bit row-wise write-out) is the
// scheme of sample_kernel.cuh; the tile is filled with scalar stores through a
// per-lane pointer that advances on every emitted draw.
#pragma once
#include <cstdlib>

#include "sample_kernel.cuh"

#ifndef MB200_ZIG_UNROLL
#define MB200_ZIG_UNROLL 4
#endif
// MEASUREMENT ONLY (never the shipped build): 1 = every attempt is accepted and the
// wedge machinery is compiled out -- the bare "next + convert + gather + multiply +
// compare + store + write-out" loop, whose throughput bounds what any scheduling of
// the wedge events could reach with this tile / table design (WRONG draws)
#ifndef MB200_ZIG_BARE
#define MB200_ZIG_BARE 0
#endif
// MEASUREMENT ONLY (never the shipped build): 1 = the full per-word instruction stream
// (wedge machinery included), but every lane stores and advances its slot pointer on
// every word, so lanes never drift apart: no drift-induced shared-memory store conflicts
// and no clean-up iterations.  Bounds what a tile layout immune to drift could gain
// (WRONG draws).
#ifndef MB200_ZIG_NODRIFT
#define MB200_ZIG_NODRIFT 0
#endif
// MEASUREMENT ONLY: 1 = pointers drift and clean-up iterations run as shipped, but the
// tile store goes to a column that depends on the iteration only (conflict free whatever
// the drift): isolates the cost of the drift-induced store conflicts (WRONG draws)
#ifndef MB200_ZIG_NOCONF
#define MB200_ZIG_NOCONF 0
#endif
// 1 = the wedge evaluation and the park block run only in the words in which some lane of the
// warp needs them (one warp vote each; 38% of the words) instead of predicated in every word
#ifndef MB200_ZIG_GATE
#define MB200_ZIG_GATE 0
#endif

namespace mb200 {

constexpr int kZigUnroll = MB200_ZIG_UNROLL;     // hot-loop unrolling (experiments: -DMB200_ZIG_UNROLL=n)
#ifndef MB200_ZIG_MAXWARPS
#define MB200_ZIG_MAXWARPS 20
#endif
#ifndef MB200_ZIG_PITCH
#define MB200_ZIG_PITCH 35
#endif
constexpr int kZigMaxWarps = MB200_ZIG_MAXWARPS;
// Row pitch of the output tile in elements: 32 slots of the window + pad slots
// that hold draws banked for the next window (see ziggurat_body).  The pitch is
// odd for the double tile: the scalar 64-bit stores of a half-warp then fall in
// 16 different bank pairs (an even pitch leaves lanes l and l + 8 on the same
// pair: two wavefronts per half-warp), at the price of odd rows that are only
// 8-byte aligned (see the write-out).  Measured: 36 doubles (4-way store
// conflicts) 450 Gdraws/s, 38 doubles (19 warps fit) 460, 34 doubles 457.7, 35 and
// 37 doubles 465.3 (same box, same day).
template <typename OutT> struct ZigTile;
template <> struct ZigTile<double> { static constexpr int PITCH = MB200_ZIG_PITCH; };
template <> struct ZigTile<float> { static constexpr int PITCH = 36; };
constexpr int kZigCopies = 8;                       // replicas of the {x, y} table
constexpr float kZigBand = 1.0e-4f;                 // |t| below this: decide in double
constexpr size_t kZigTableBytes = 256 * kZigCopies * sizeof(double2) + 256 * sizeof(float4);

__device__ __forceinline__ void sts_out(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_out(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}

// shared-memory tables: zig8[i][j] = {x[i], y[i]} (j = copy), wtab[i] =
// {fd, fa - fd, 2 s x[i], -3 s x[i]} with f(x) = exp(-x^2 / 2), fa = f(x[i+1]),
// fd = f(x[i]) - f(x[i+1]), s = sqrt(log2(e) / 2).
//
// In the double path the look-up works on U = u0 * 2^63 (see zig_word), so the
// table holds {x * 2^-63, y * 2^63}: power-of-two scalings, exact, and
// fl(U * (x 2^-63)) == fl(u0 * x) bit for bit.
template <typename T>
__device__ __forceinline__ void zig_stage_tables(unsigned char* base) {
  double2* zig8 = reinterpret_cast<double2*>(base);
  float4* wtab = reinterpret_cast<float4*>(base + 256 * kZigCopies * sizeof(double2));
  const double up = sizeof(T) == 8 ? 9223372036854775808.0 : 1.0;             // 2^63
  const double dn = sizeof(T) == 8 ? 1.0842021724855044340e-19 : 1.0;         // 2^-63
  for (int i = threadIdx.x; i < 256 * kZigCopies; i += blockDim.x) {
    const double2 xy = g_zig_xy[i / kZigCopies];
    zig8[i] = make_double2(xy.x * dn, xy.y * up);
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const double xi = g_zig_x[i], x1 = g_zig_x[i + 1];
    const double fi = exp(-0.5 * xi * xi), f1 = exp(-0.5 * x1 * x1);
    const double s = 0.84932180028801904272;      // sqrt(0.5 * log2(e))
    wtab[i] = make_float4(static_cast<float>(fi - f1), static_cast<float>(f1 - (fi - f1)),
                          static_cast<float>(2.0 * s * xi), static_cast<float>(-3.0 * s * xi));
    // layer 0 (tail): yr = 1, z^2 term = 0, so t == 0 and the rare branch is taken
    if (i == 0) wtab[i] = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
  }
  __syncthreads();
}

// hdr/normal_ziggurat.hpp:14-30 with the first uniform already drawn (the
// word that follows a parked layer-0 candidate).  Out of line, state by value
// (see ziggurat_tail in samplers.cuh).
template <typename T>
__device__ __noinline__ TailResult<T> ziggurat_tail_from(T u1, X256 state, T x1, bool negative) {
  Rng r{state};
  for (;;) {
    const T u2 = random_real<T>(r);
    // m_log, not m_log_normal: the result is the same and this routine is cold,
    // but the shorter variant shifts the register allocation of the whole
    // kernel and measures 1% slower (453 vs 458 Gdraws/s)
    const T x = m_log(u1) / x1;
    const T y = m_log(u2);
    if (-2 * y > x * x) return TailResult<T>{negative ? x - x1 : x1 - x, r.s};
    u1 = random_real<T>(r);
  }
}

__device__ __forceinline__ uint32_t bump(uint32_t p, uint32_t by) {
  return __float_as_uint(__fadd_rn(__uint_as_float(p), __uint_as_float(by)));
}

// elements e and e + 16 of a row (odd rows of the odd-pitch double tile)
__device__ __forceinline__ double2 zig_lds_pair(uint32_t addr, double) {
  double2 v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v.x) : "r"(addr));
  asm volatile("ld.shared.f64 %0, [%1+128];" : "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 zig_lds_pair(uint32_t, float) { return float4{0, 0, 0, 0}; }   // never used
__device__ __forceinline__ void zig_stg_pair(char* dst, const double2& v) {
  *reinterpret_cast<double*>(dst) = v.x;
  *reinterpret_cast<double*>(dst + 128) = v.y;
}
__device__ __forceinline__ void zig_stg_pair(char*, const float4&) {}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Per-lane sampler state that lives across iterations.
template <typename T, typename OutT>
struct ZigLane {
  Rng rng;
  bool wedge;        // mode: the next word is the u1 of the parked candidate
  double wz;         // parked z = u0 * x[i] (not kept when the tile slot holds it, see SAME)
  float wfd, wfa;    // wedge constants of the parked layer
  float wzs;         // z^2 * log2(e) / 2 of the parked candidate (single precision)
  uint32_t widx;     // parked layer * 128 + copy offset
  uint32_t p;        // shared address of the next output slot
#if MB200_ZIG_NOCONF
  uint32_t q, qbase;
#endif
  T mean, sd;
};

// One generator word of one lane.
//
//   FAST  (!wedge): the word is an attempt.  Its value z * sd + mean is stored
//          at the lane's next output slot straight away; the slot pointer moves
//          on only if |u0| < y[i].  Otherwise the candidate stays parked in that
//          slot, the wedge constants of its layer are loaded and the lane
//          switches to WEDGE.
//   WEDGE : the word is u1.  Accept -> the pointer moves past the parked value,
//          reject -> the next attempt overwrites it.  Back to FAST either way.
//
// Layer 0 has no wedge: its wtab entry makes t == 0, so the candidate always
// reaches the rare branch, which runs the tail algorithm with this word as its
// first uniform -- exactly the word the reference's tail would draw first.
template <typename T, typename OutT, bool STANDARD>
__device__ __forceinline__ void zig_word(ZigLane<T, OutT>& L, uint32_t zig_saddr, uint32_t lane_off,
                                         uint32_t wtab_lane_saddr) {
  // standard normal in double: the tile slot holds the parked z itself
  constexpr bool SAME = STANDARD && sizeof(T) == 8 && sizeof(OutT) == 8;
  const uint64_t v = L.rng.next();
  const uint32_t vlo = static_cast<uint32_t>(v), vhi = static_cast<uint32_t>(v >> 32);
  double u0, td = 0;
  float rf = 0;
  if (sizeof(T) == 8) {
    // hdr/utils.hpp:21-25 and normal_ziggurat.hpp:94 in two roundings that
    // coincide with the reference's: t = (x & ~0x7ff) + 2^10 rounds exactly like
    // (x >> 11) * 2^-53 + 2^-54 (same significand, scaled by 2^64), and
    // 2 r - 1 is exact in both
    // U = u0 * 2^63 = t - 2^63 is exact too and needs no register constant;
    // the table is scaled to match (zig_stage_tables)
    td = __dadd_rn(__ull2double_rn(v & ~0x7ffull), 1024.0);
    u0 = __dadd_rn(td, -9223372036854775808.0);
  } else {
    rf = u64_to_float(v);
    u0 = static_cast<double>(fma_exact2m1(rf));       // the tables are double: promoted like the reference
  }
  const uint32_t idx = ((vlo << 4) & 0x7f80u) | lane_off;    // ((v >> 3) & 255) * 128 + copy * 16
  const double2 xy = lds_f64x2(zig_saddr + idx);
#if MB200_ZIG_BARE
  const bool inside = fabs(u0) < xy.y * 4.0;        // always true: the compare is kept, its outcome is not
  const bool fast = inside;
  const bool park = false;
  L.wedge = false;
#else
  const bool inside = fabs(u0) < xy.y;
  const bool fast = !L.wedge && inside;
  const bool park = !L.wedge && !inside;
#endif
  const double zx = u0 * xy.x;
  const T zt = static_cast<T>(zx);
  const OutT fy = static_cast<OutT>(STANDARD ? zt : zt * L.sd + L.mean);
#if MB200_ZIG_NODRIFT
  sts_out(L.p, fy);
#elif MB200_ZIG_NOCONF
  if (!L.wedge) sts_out(L.qbase + L.q, fy);
  L.q = (L.q + sizeof(OutT)) & (32 * sizeof(OutT) - 1);      // 32 columns round and round, the same for every lane
#else
  if (!L.wedge) sts_out(L.p, fy);
#endif

#if MB200_ZIG_BARE
  if (fast) L.p = bump(L.p, sizeof(OutT));
  return;
#endif
#if MB200_ZIG_GATE
  bool wemit = false;
  if (__any_sync(__activemask(), L.wedge)) {
  const float m = __uint_as_float(__umulhi(vhi, 1u << 23) + 0x3f800000u);
  const float yr = __fmaf_rn(m, L.wfd, L.wfa);
  const float t = lg2_approx(yr) + L.wzs;
  wemit = L.wedge && t < 0.0f;
  if (__builtin_expect(L.wedge && fabsf(t) <= kZigBand, 0)) {
#else
  // 23-bit image of the word: m = 1 + floor(hi / 2^9) / 2^23 in [1, 2)
  const float m = __uint_as_float(__umulhi(vhi, 1u << 23) + 0x3f800000u);
  // the parked candidate against this word (used only when L.wedge)
  const float yr = __fmaf_rn(m, L.wfd, L.wfa);
  const float t = lg2_approx(yr) + L.wzs;
  bool wemit = L.wedge && t < 0.0f;
  if (__builtin_expect(L.wedge && fabsf(t) <= kZigBand, 0)) {
#endif
    const T r = sizeof(T) == 8 ? static_cast<T>(td * 5.421010862427522170e-20 /* 2^-64 */) : static_cast<T>(rf);
    const double z = SAME ? lds_f64(L.p) : L.wz;
    const int layer = static_cast<int>(L.widx >> 7);
    if (layer == 0) {
      const TailResult<T> tr = ziggurat_tail_from<T>(r, L.rng.s, static_cast<T>(g_zig_x[1]), z < 0);
      L.rng.s = tr.state;
      sts_out(L.p, static_cast<OutT>(STANDARD ? tr.value : tr.value * L.sd + L.mean));
      wemit = true;
    } else {
      // the reference's own evaluation (hdr/normal_ziggurat.hpp:104-107)
      const double xi = g_zig_x[layer], x1 = g_zig_x[layer + 1];
      const double g0 = m_exp(-0.5 * (xi * xi - z * z));
      const double g1 = m_exp(-0.5 * (x1 * x1 - z * z));
      wemit = g1 + r * (g0 - g1) < 1.0;
    }
  }
#if MB200_ZIG_GATE
  }
#endif
  // the slot pointer moves on the FMA pipe (the loop is ALU-pipe bound): a shared
  // address is below 2^23, so as a float bit pattern it is a subnormal, and
  // adding the subnormal whose pattern is sizeof(OutT) is exact integer addition
  // (no flush-to-zero in this translation unit)
#if MB200_ZIG_NODRIFT
  // two pointer updates per word as in the shipped build, net one slot whatever happened
  L.p = bump(L.p, fast ? 2 * sizeof(OutT) : sizeof(OutT));
  if (fast) L.p = L.p - sizeof(OutT);
#else
  if (fast) L.p = bump(L.p, sizeof(OutT));
  if (wemit) L.p = bump(L.p, sizeof(OutT));
#endif
#if MB200_ZIG_GATE
  if (__any_sync(__activemask(), park) && park) {
    const float m = __uint_as_float(__umulhi(vhi, 1u << 23) + 0x3f800000u);
#else
  if (park) {
#endif
    const float4 w = lds_f32x4(wtab_lane_saddr + (idx >> 3));
    const float zr = __fmaf_rn(m, w.z, w.w);
    L.wfd = w.x;
    L.wfa = w.y;
    L.wzs = zr * zr;
    L.widx = idx;
    if (!SAME) L.wz = zx;
  }
  L.wedge = park;
}

template <typename T, typename OutT, bool STANDARD>
__device__ __forceinline__ void ziggurat_body(const SampleArgs& a) {
  constexpr int VEC = TileCfg<OutT>::VEC;
  constexpr int PITCH = ZigTile<OutT>::PITCH;
  constexpr int LPR = kTile / VEC;         // lanes per row in the write-out
  constexpr int RPI = 32 / LPR;            // rows per store instruction
  constexpr int STEPS = 32 / RPI;
  using V = typename std::conditional<sizeof(OutT) == 8, double2, float4>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // tables first: their shared addresses are link-time constants
  zig_stage_tables<T>(smem_raw);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  OutT* tile = reinterpret_cast<OutT*>(smem_raw + kZigTableBytes) + warp * 32 * PITCH;
  const uint32_t tile_saddr = smem_addr(tile);
  const uint32_t row_saddr = tile_saddr + static_cast<uint32_t>(lane * PITCH * sizeof(OutT));
  const uint32_t zig_saddr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  const uint32_t lane_off = static_cast<uint32_t>((lane & (kZigCopies - 1)) * sizeof(double2));
  // wtab entry of a look-up index idx = layer * 128 + lane_off is at (idx >> 3) from here
  const uint32_t wtab_lane_saddr = zig_saddr + static_cast<uint32_t>(256 * kZigCopies * sizeof(double2)) - (lane_off >> 3);
  OutT* const out = static_cast<OutT*>(a.out);
  const unsigned long long n_streams = a.n_streams;
  const uint32_t n = static_cast<uint32_t>(a.n_samples);        // host guarantees 2 <= n <= 2^24
  const unsigned long long n_warp_tiles = (n_streams + 31) / 32;
  const bool vec_ok = (n % VEC == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  // write-out geometry of this lane: 16-byte unit `sub` of rows rsub + t * RPI.
  // With an odd pitch (SPLIT, the double tile) the odd rows are only 8-byte
  // aligned: step t then covers rows 4 (t / 2) + (t % 2) + 2 rsub, so that a
  // step is all-even (one 128-bit access per lane as before) or all-odd (two
  // 64-bit accesses per lane: elements sub1 and sub1 + 16 of the row, each
  // half-warp moving 128 contiguous bytes).
  constexpr bool SPLIT = (PITCH & 1) != 0;
  static_assert(!SPLIT || (sizeof(OutT) == 8 && RPI == 2), "odd pitch: double tile only");
  const int sub = (lane % LPR) * VEC;
  const int sub1 = lane % LPR;
  const int rsub = lane / LPR;
  const uint32_t rd_saddr = tile_saddr + static_cast<uint32_t>(((SPLIT ? 2 * rsub : rsub) * PITCH + sub) * sizeof(OutT));
  const uint32_t rd1_saddr = tile_saddr + static_cast<uint32_t>((2 * rsub * PITCH + sub1) * sizeof(OutT));

  for (;;) {
    // next 32-stream tile (the counter starts at ~0u: reset_err's 0xFF fill)
    unsigned wt32 = 0;
    if (lane == 0) wt32 = atomicAdd(&a.err->work, 1u) + 1u;
    wt32 = __shfl_sync(0xffffffffu, wt32, 0);
    if (wt32 >= n_warp_tiles) break;
    const unsigned long long s_first = static_cast<unsigned long long>(wt32) * 32;
    const int n_rows = (n_streams - s_first) < 32 ? static_cast<int>(n_streams - s_first) : 32;
    const unsigned long long stream = s_first + lane;
    const bool exists = stream < n_streams;
    OutT* const wbase = out + s_first * n;
    // row r starts m_r elements into its first window
    const uint32_t sm = static_cast<uint32_t>(s_first & 31u), nm = n & 31u;
    uint32_t roff[STEPS];                       // element offset of (row, sub) from wbase, window 0
#pragma unroll
    for (int t = 0; t < STEPS; ++t) {
      const uint32_t r = SPLIT ? 4 * (t >> 1) + (t & 1) + 2 * rsub : rsub + t * RPI;
      const uint32_t col = (SPLIT && (t & 1)) ? sub1 : sub;
      roff[t] = (r * n - (((sm + r) * nm) & 31u) + col) * static_cast<uint32_t>(sizeof(OutT));   // bytes
      // opaque: keeps the 16 offsets in registers instead of re-deriving them
      // (6 ALU instructions each) at every window
      asm volatile("" : "+r"(roff[t]));
    }

    ZigLane<T, OutT> L;
    L.wedge = false;
    L.wz = 0; L.wfd = 0; L.wfa = 1.0f; L.wzs = 0; L.widx = 128;
    L.mean = 0; L.sd = 1;
    if (exists) {
      L.rng.s = load_state(a.state, stream);
      L.mean = static_cast<T>(a.par[0][a.pstride[0] ? stream : 0]);
      L.sd = static_cast<T>(a.par[1][a.pstride[1] ? stream : 0]);
    } else {
      L.rng.s = X256{1, 2, 3, 4};
    }
    uint32_t remaining = exists ? n : 0;
    uint32_t e_lo = ((sm + lane) * nm) & 31u;   // m of this lane's row
    uint32_t wpos = 0;                          // window index * 32
    // Draws of the NEXT window already sitting in this lane's row.  A lane that
    // completes its window keeps drawing into the kLead pad slots behind it for
    // as long as the warp runs clean-up iterations for the lanes that parked
    // candidates, so those iterations do useful work on (nearly) every lane and
    // a lane that is ahead absorbs its own next wedge events: 3.7 -> 2.1
    // clean-up iterations per window with the two pad slots of the double tile.
    // The banked draws are the stream's next draws in order, so nothing changes
    // per stream; they move to the front of the row after the write-out.
    constexpr uint32_t kLead = PITCH - kTile;
    constexpr uint32_t kSz = sizeof(OutT);
    uint32_t lead = 0;
    for (;;) {
      const uint32_t room = kTile - e_lo;
      const uint32_t cnt = remaining < room ? remaining : room;
      if (__reduce_max_sync(0xffffffffu, cnt) == 0) break;
      remaining -= cnt;
      const uint32_t extra = remaining < kLead ? remaining : kLead;     // never past the end of the stream's batch
      L.p = row_saddr + (e_lo + lead) * kSz;
#if MB200_ZIG_NOCONF
      L.q = 0;
      L.qbase = row_saddr;
#endif
      const uint32_t p_end = row_saddr + (e_lo + cnt) * kSz;
      const uint32_t p_cap = p_end + extra * kSz;
      const uint32_t n_hot = __reduce_min_sync(0xffffffffu, (p_cap - L.p) / kSz);
      const bool all_full = __reduce_min_sync(0xffffffffu, cnt) == kTile;
      // hot loop: every lane has room for n_hot more outputs and an iteration
      // emits at most one
#pragma unroll kZigUnroll
      for (uint32_t i = 0; i < n_hot; ++i) zig_word<T, OutT, STANDARD>(L, zig_saddr, lane_off, wtab_lane_saddr);
      // lanes that parked candidates (or had more room) finish here; the others
      // run ahead into their pad slots meanwhile
      while (__any_sync(0xffffffffu, L.p < p_end)) {
        if (L.p < p_cap) zig_word<T, OutT, STANDARD>(L, zig_saddr, lane_off, wtab_lane_saddr);
      }
      lead = L.p > p_end ? (L.p - p_end) / kSz : 0;
      __syncwarp();
      if (vec_ok && all_full) {
        // every row is full: STEPS x (RPI rows x 256 B / 128 B), loads batched before stores
        constexpr int HALF = STEPS / 2;
        char* const wwin = reinterpret_cast<char*>(wbase + wpos);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          V v[HALF];
#pragma unroll
          for (int t = 0; t < HALF; ++t) {
            const int step = h * HALF + t;
            if (!SPLIT) {
              v[t] = lds_vec(rd_saddr + static_cast<uint32_t>(step * RPI * PITCH * sizeof(OutT)), OutT());
            } else {
              const uint32_t row0 = static_cast<uint32_t>((4 * (step >> 1) + (step & 1)) * PITCH * sizeof(OutT));
              if (step & 1) v[t] = zig_lds_pair(rd1_saddr + row0, OutT());
              else v[t] = lds_vec(rd_saddr + row0, OutT());
            }
          }
#pragma unroll
          for (int t = 0; t < HALF; ++t) {
            const int step = h * HALF + t;
            char* const dst = wwin + roff[step];          // 64-bit base + 32-bit offset
            if (SPLIT && (step & 1)) zig_stg_pair(dst, v[t]);
            else *reinterpret_cast<V*>(dst) = v[t];
          }
        }
      } else {
        // ragged window: lane e of step r writes element e of row r if it belongs to the row
        for (int r = 0; r < n_rows; ++r) {
          const uint32_t m_r = ((sm + r) * nm) & 31u;
          const uint32_t j = wpos + lane - m_r;            // sample index (wraps below zero)
          if (wpos + lane >= m_r && j < n) wbase[static_cast<size_t>(r) * n + j] = tile[r * PITCH + lane];
        }
      }
      __syncwarp();
      // banked draws (and a candidate parked behind them) move to the front
      {
        OutT carry[kLead];
#pragma unroll
        for (uint32_t c = 0; c < kLead; ++c) carry[c] = tile[lane * PITCH + kTile + c];
#pragma unroll
        for (uint32_t c = 0; c < kLead; ++c) tile[lane * PITCH + c] = carry[c];
      }
      e_lo = 0;
      wpos += kTile;
    }
    if (exists) store_state(a.state, stream, L.rng.s);
  }
}

// mean == 0 and sd == 1 given as scalars: z * 1 + 0 == z bit for bit (z = u0 * x
// is never a zero: |u0| >= 2^-53), so the standard normal skips the scaling.
// The test is uniform over the grid; both bodies live in one kernel.
template <typename T, typename OutT>
__global__ void __launch_bounds__(kZigMaxWarps * 32, 1)
ziggurat_kernel(const SampleArgs a) {
  const bool standard = a.pstride[0] == 0 && a.pstride[1] == 0 && a.par[0][0] == 0.0 && a.par[1][0] == 1.0;
  if (standard) ziggurat_body<T, OutT, true>(a);
  else ziggurat_body<T, OutT, false>(a);
}

// The batched kernel handles 2 <= n_samples <= 2^24 and fewer than 2^36 streams
// (32-bit tile counter).  The write-out offsets are 32-bit BYTE offsets from the
// first row of a warp-tile: (31 n + 62) * 8 must stay below 2^32.
inline bool ziggurat_kernel_applies(const SampleArgs& a) {
  return a.n_samples >= 2 && a.n_samples <= (1ull << 24) && a.n_streams < (1ull << 36);
}

template <typename T, typename OutT>
cudaError_t launch_ziggurat_t(const SampleArgs& a, int sm_count, cudaStream_t st) {
  const unsigned long long n_tiles = (a.n_streams + 31) / 32;
  unsigned long long wpb = (n_tiles + sm_count - 1) / sm_count;      // spread small problems over the SMs
  unsigned long long max_warps = kZigMaxWarps;
  if (const char* w = std::getenv("MB200_ZIG_WARPS")) {       // occupancy experiments
    const long v = std::atol(w);
    if (v >= 1 && v <= kZigMaxWarps) max_warps = static_cast<unsigned long long>(v);
  }
  if (wpb > max_warps) wpb = max_warps;
  if (wpb < 1) wpb = 1;
  unsigned long long grid = (n_tiles + wpb - 1) / wpb;
  if (grid > static_cast<unsigned long long>(sm_count)) grid = sm_count;
  const size_t smem = sizeof(OutT) * wpb * 32 * ZigTile<OutT>::PITCH + kZigTableBytes;
  cudaError_t e = cudaFuncSetAttribute(ziggurat_kernel<T, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  ziggurat_kernel<T, OutT><<<static_cast<unsigned>(grid), static_cast<unsigned>(wpb * 32), smem, st>>>(a);
  return cudaGetLastError();
}

inline cudaError_t launch_ziggurat(int real_type, int out_f32, const SampleArgs& a, int sm_count, cudaStream_t st) {
  if (real_type == 1) {
    return out_f32 ? launch_ziggurat_t<float, float>(a, sm_count, st)
                   : launch_ziggurat_t<float, double>(a, sm_count, st);
  }
  return launch_ziggurat_t<double, double>(a, sm_count, st);
}

}  // namespace mb200
