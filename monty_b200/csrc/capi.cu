// capi.cu -- host side of the C-ABI declared in include/monty_b200.h.
//
// This is the B200 replacement of the reference's cpp11 layer for the RNG /
// sampler / density path: it owns the device-resident stream state, recycles
// and validates parameters like src/random.cpp:21-82, launches the kernels,
// and reports errors with the reference's exact text.  Each entry point cites
// the routine it replaces.  There is no CPU fallback anywhere in this file:
// every sample is produced by a CUDA kernel or the call fails.
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/monty_b200.h"
#include "density_dispatch.h"
#include "deterministic_dispatch.h"
#include "errors.h"
#include "jump_poly.h"
#include "launch_args.h"
#include "multinomial_dispatch.h"
#include "rng_dispatch.h"
#include "sample_dispatch.h"

using namespace mb200;

namespace {

thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};

int fail(int code, mb200_rng* rng, const char* fmt, ...);

const JumpAlgebra& algebra() {
  static JumpAlgebra a;   // throws if the self-check against the reference constants fails
  return a;
}

}  // namespace

struct mb200_rng {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  uint64_t* d_state = nullptr;
  size_t n_streams = 0;
  bool deterministic = false;
  ErrRec* d_err = nullptr;
  ErrRec* h_err = nullptr;       // pinned
  double* d_par[MB200_MAX_PARAMS] = {nullptr, nullptr, nullptr, nullptr};
  size_t par_cap[MB200_MAX_PARAMS] = {0, 0, 0, 0};
  void* d_out = nullptr;
  size_t out_cap = 0;
  // pinned staging ring for transfers to / from PAGEABLE host vectors (what R
  // hands us), allocated on first use: see copy_to_host / copy_from_host
  char* stage[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[3] = {nullptr, nullptr, nullptr};
  std::string last_error;
  // deferred error context of the last device-resident sample call
  bool pending = false;
  int pending_dist = 0, pending_rt = 0;
  const double* pending_par[MB200_MAX_PARAMS] = {nullptr, nullptr, nullptr, nullptr};
  size_t pending_len[MB200_MAX_PARAMS] = {0, 0, 0, 0};
};

namespace {

int fail(int code, mb200_rng* rng, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_last_error = buf;
  if (rng) rng->last_error = buf;
  return code;
}

// Entry points switch to the handle's device; the caller's current device is
// restored on return (mb200_*_dev are called from inside torch processes).
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CU(call)                                                                   \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess)                                                        \
      return fail(MB200_ERR_CUDA, rng, "CUDA error: %s (%s) at %s:%d",             \
                  cudaGetErrorString(e__), #call, __FILE__, __LINE__);             \
  } while (0)

int ensure(mb200_rng* rng, void** p, size_t* cap, size_t bytes) {
  if (*cap >= bytes) return MB200_OK;
  if (*p) { CU(cudaStreamSynchronize(rng->stream)); CU(cudaFree(*p)); *p = nullptr; *cap = 0; }
  CU(cudaMalloc(p, bytes));
  *cap = bytes;
  return MB200_OK;
}

// ---------------------------------------------------------------------------
// Transfers between device memory and PAGEABLE host memory.
//
// R's vectors (the parameters it passes, the cpp11::writable::doubles the
// result lands in) are ordinary malloc'ed memory.  cudaMemcpy on such memory
// goes through the driver's internal staging buffer with a single-threaded
// host copy: 10-25 GB/s, and the call blocks.  Here large transfers are cut
// into 32 MB chunks that travel by DMA between the device and a ring of three
// pinned staging buffers (full PCIe rate, asynchronous), while a small pool of
// host threads copies the previous chunk between the ring and the caller's
// vector -- the two overlap, and the host side runs at several threads' memory
// bandwidth.  Pinned callers (the bench's e2e leg, anything cudaHostRegister'ed)
// and small transfers keep the plain asynchronous copy.
// ---------------------------------------------------------------------------
constexpr size_t kStageChunk = 32u << 20;
constexpr size_t kStageMin = 8u << 20;        // below this one plain copy is cheaper

class CopyPool {
 public:
  // never destroyed: its (detached) workers outlive static destruction
  static CopyPool& get() { static CopyPool* p = new CopyPool(); return *p; }
  // memcpy(dst, src, n) split over the pool's threads; returns when done
  void copy(char* dst, const char* src, size_t n) {
    const size_t nt = workers_.size();
    if (nt == 0 || n < (4u << 20)) { std::memcpy(dst, src, n); return; }
    std::unique_lock<std::mutex> lock(m_);
    dst_ = dst; src_ = src; n_ = n; pending_ = nt; ++generation_;
    cv_.notify_all();
    done_.wait(lock, [&] { return pending_ == 0; });
  }
 private:
  CopyPool() {
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = hw >= 16 ? 8 : (hw >= 4 ? hw / 2 : 0);
    for (unsigned t = 0; t < nt; ++t) workers_.emplace_back([this, t, nt] { run(t, nt); });
    for (auto& w : workers_) w.detach();
  }
  void run(unsigned t, unsigned nt) {
    unsigned long long seen = 0;
    for (;;) {
      char* d; const char* s; size_t n;
      {
        std::unique_lock<std::mutex> lock(m_);
        cv_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_; d = dst_; s = src_; n = n_;
      }
      const size_t per = ((n / nt) + 4095) & ~static_cast<size_t>(4095);
      const size_t lo = per * t < n ? per * t : n, hi = per * (t + 1) < n ? per * (t + 1) : n;
      if (t == nt - 1 && hi < n) std::memcpy(d + hi, s + hi, n - hi);
      if (hi > lo) std::memcpy(d + lo, s + lo, hi - lo);
      {
        std::lock_guard<std::mutex> lock(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  char* dst_ = nullptr; const char* src_ = nullptr; size_t n_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
};

bool host_is_pageable(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}

int ensure_stage(mb200_rng* rng) {
  for (int s = 0; s < 3; ++s) {
    if (!rng->stage[s]) CU(cudaMallocHost(&rng->stage[s], kStageChunk));
    if (!rng->stage_ev[s]) CU(cudaEventCreateWithFlags(&rng->stage_ev[s], cudaEventDisableTiming));
  }
  return MB200_OK;
}

// device -> host vector, ordered on the handle's stream.  Large pageable
// destinations are complete on return; otherwise the copy is queued and the
// caller's next synchronisation (fetch_err) completes it.
int copy_to_host(mb200_rng* rng, void* dst_host, const void* src_dev, size_t bytes) {
  if (bytes < kStageMin || !host_is_pageable(dst_host)) {
    CU(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, rng->stream));
    return MB200_OK;
  }
  int rc = ensure_stage(rng);
  if (rc != MB200_OK) return rc;
  const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
  auto len = [&](size_t i) { return i + 1 < n_chunks ? kStageChunk : bytes - i * kStageChunk; };
  auto issue = [&](size_t i) -> cudaError_t {
    cudaError_t e = cudaMemcpyAsync(rng->stage[i % 3], static_cast<const char*>(src_dev) + i * kStageChunk, len(i),
                                    cudaMemcpyDeviceToHost, rng->stream);
    return e == cudaSuccess ? cudaEventRecord(rng->stage_ev[i % 3], rng->stream) : e;
  };
  for (size_t i = 0; i < n_chunks && i < 3; ++i) CU(issue(i));
  for (size_t i = 0; i < n_chunks; ++i) {
    CU(cudaEventSynchronize(rng->stage_ev[i % 3]));
    CopyPool::get().copy(static_cast<char*>(dst_host) + i * kStageChunk, rng->stage[i % 3], len(i));
    if (i + 3 < n_chunks) CU(issue(i + 3));
  }
  return MB200_OK;
}

// host vector -> device, ordered on the handle's stream (the source may be
// reused as soon as the call returns only in the staged case; the plain
// asynchronous copy of a pageable source is staged by the driver before it returns)
int copy_from_host(mb200_rng* rng, void* dst_dev, const void* src_host, size_t bytes) {
  if (bytes < kStageMin || !host_is_pageable(src_host)) {
    CU(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, rng->stream));
    return MB200_OK;
  }
  int rc = ensure_stage(rng);
  if (rc != MB200_OK) return rc;
  const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
  for (size_t i = 0; i < n_chunks; ++i) {
    const size_t n = i + 1 < n_chunks ? kStageChunk : bytes - i * kStageChunk;
    CU(cudaEventSynchronize(rng->stage_ev[i % 3]));        // the slot's previous DMA (this or an earlier call) has read it
    CopyPool::get().copy(rng->stage[i % 3], static_cast<const char*>(src_host) + i * kStageChunk, n);
    CU(cudaMemcpyAsync(static_cast<char*>(dst_dev) + i * kStageChunk, rng->stage[i % 3], n, cudaMemcpyHostToDevice,
                       rng->stream));
    CU(cudaEventRecord(rng->stage_ev[i % 3], rng->stream));
  }
  return MB200_OK;
}

// Per-device scratch of the host-buffer density call (mb200_density).
struct DensityScratch {
  std::mutex m;
  mb200_rng* io = nullptr;       // stream + staging ring (no generator state)
  void* x = nullptr; size_t x_cap = 0;
  void* p[MB200_MAX_PARAMS] = {nullptr, nullptr, nullptr, nullptr};
  size_t p_cap[MB200_MAX_PARAMS] = {0, 0, 0, 0};
  void* out = nullptr; size_t out_cap = 0;
  double* sum = nullptr;
  double* h_sum = nullptr;       // pinned
};
DensityScratch* g_density_scratch[64];
std::mutex g_density_scratch_mutex;

int density_scratch(int device, DensityScratch** out) {
  mb200_rng* rng = nullptr;
  if (device < 0 || device >= 64) return fail(MB200_ERR_ARG, nullptr, "device ordinal %d out of range", device);
  std::lock_guard<std::mutex> lock(g_density_scratch_mutex);
  if (!g_density_scratch[device]) {
    std::unique_ptr<DensityScratch> sc(new DensityScratch());
    std::unique_ptr<mb200_rng> io(new mb200_rng());
    io->device = device;
    CU(cudaStreamCreateWithFlags(&io->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&sc->sum, sizeof(double)));
    CU(cudaMallocHost(&sc->h_sum, sizeof(double)));
    sc->io = io.release();
    g_density_scratch[device] = sc.release();
  }
  *out = g_density_scratch[device];
  return MB200_OK;
}

// One lgamma-of-integers table per device (samplers.cuh: m_lgamma), built by
// the device itself the first time a handle or a density call touches it.
constexpr int kLgammaTableN = 16384;        // == kLgtN
struct LgammaTable { double* d = nullptr; float* f = nullptr; bool ready = false; };
// Per-device scratch for the per-block partial sums of the fused density
// reduction: a ring of slots handed out round-robin, so a call never allocates
// (cudaMallocAsync after a synchronize costs up to milliseconds: the pool gives
// its memory back) and up to kPartialSlots calls can be in flight per device.
constexpr int kPartialSlots = 64;
constexpr int kPartialSlotDoubles = 2048;     // >= density_grid(): 8 blocks per SM
struct PartialRing {
  double* base = nullptr;
  std::atomic<unsigned> next{0};
  // recorded after the reduction that reads a slot; the next user of the slot --
  // possibly on another stream -- waits for it before its kernel writes there
  cudaEvent_t used[kPartialSlots] = {};
};
PartialRing g_partials[64];
LgammaTable g_lgamma_table[64];
std::mutex g_lgamma_mutex;

cudaError_t ensure_lgamma_table(int device, cudaStream_t st) {
  if (device < 0 || device >= 64) return cudaSuccess;         // no table: every call computes lgamma
  std::lock_guard<std::mutex> lock(g_lgamma_mutex);
  LgammaTable& t = g_lgamma_table[device];
  if (t.ready) return cudaSuccess;
  cudaError_t e = cudaMalloc(&t.d, kLgammaTableN * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&t.f, kLgammaTableN * sizeof(float));
  if (e == cudaSuccess) e = launch_lgamma_table_build(t.d, t.f, kLgammaTableN, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = install_lgamma_table_group_b(t.d, t.f);
  if (e == cudaSuccess) e = install_lgamma_table_group_c(t.d, t.f);
  if (e == cudaSuccess) e = install_lgamma_table_density(t.d, t.f);
  if (e == cudaSuccess) {
    e = cudaMalloc(&g_partials[device].base, sizeof(double) * kPartialSlots * kPartialSlotDoubles);
  }
  t.ready = e == cudaSuccess;
  if (!t.ready) {                      // nothing half-built survives a failure: the next call starts over
    cudaFree(t.d); cudaFree(t.f); cudaFree(g_partials[device].base);
    t.d = nullptr; t.f = nullptr; g_partials[device].base = nullptr;
  }
  return e;
}

int reset_err(mb200_rng* rng) {
  CU(cudaMemsetAsync(rng->d_err, 0xFF, sizeof(ErrRec), rng->stream));
  return MB200_OK;
}

// copies the error record back and waits for the stream
int fetch_err(mb200_rng* rng) {
  CU(cudaMemcpyAsync(rng->h_err, rng->d_err, sizeof(ErrRec), cudaMemcpyDeviceToHost, rng->stream));
  CU(cudaStreamSynchronize(rng->stream));
  return MB200_OK;
}

cudaError_t launch_sample(int dist, int real_type, int out_f32, bool validate_only,
                          const SampleArgs& a, int sm_count, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (dist_in_group_a(dist)) return launch_sample_group_a(dist, real_type, out_f32, validate_only, a, sm_count, st);
  if (dist_in_group_b(dist)) return launch_sample_group_b(dist, real_type, out_f32, validate_only, a, sm_count, st);
  return launch_sample_group_c(dist, real_type, out_f32, validate_only, a, sm_count, st);
}

// ref: src/random.cpp:21-42 (struct input): length 1 (broadcast) or n_streams
int check_param_lengths(mb200_rng* rng, int dist, const size_t* param_len) {
  const int np = dist_n_params(dist);
  const char* const* names = dist_param_names(dist);
  for (int k = 0; k < np; ++k) {
    const size_t len = param_len[k];
    if (len != 1 && len != rng->n_streams) {
      if (rng->n_streams == 1) {
        return fail(MB200_ERR_ARG, rng, "Expected '%s' to have length %d, not %d",
                    names[k], (int)rng->n_streams, (int)len);
      }
      return fail(MB200_ERR_ARG, rng, "Expected '%s' to have length %d or 1, not %d",
                  names[k], (int)rng->n_streams, (int)len);
    }
  }
  return MB200_OK;
}

int param_error(mb200_rng* rng, int dist, int real_type, const double* const* params,
                const size_t* param_len, bool params_on_device, unsigned long long stream) {
  double v[MB200_MAX_PARAMS] = {0, 0, 0, 0};
  const int np = dist_n_params(dist);
  for (int k = 0; k < np; ++k) {
    const double* src = params[k] + (param_len[k] == 1 ? 0 : stream);
    if (params_on_device) {
      if (cudaMemcpy(&v[k], src, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) v[k] = NAN;
    } else {
      v[k] = *src;
    }
  }
  const std::string msg = format_param_error(dist, real_type, v);
  return fail(MB200_ERR_PARAM, rng, "%s", msg.c_str());
}

// deterministic = TRUE: every sampler returns its expectation and leaves the
// stream alone (deterministic_kernel.cu).  The values of the streams before
// the first invalid one are produced, then the reference's error is raised.
int sample_deterministic(mb200_rng* rng, int dist, int real_type, size_t n_samples,
                         const double* const* params_host, const size_t* param_len, double* out_host) {
  const int np = dist_n_params(dist);
  SampleArgs a{};
  a.state = rng->d_state;
  int rc;
  for (int k = 0; k < np; ++k) {
    const size_t bytes = param_len[k] * sizeof(double);
    rc = ensure(rng, reinterpret_cast<void**>(&rng->d_par[k]), &rng->par_cap[k], bytes);
    if (rc != MB200_OK) return rc;
    CU(cudaMemcpyAsync(rng->d_par[k], params_host[k], bytes, cudaMemcpyHostToDevice, rng->stream));
    a.par[k] = rng->d_par[k];
    a.pstride[k] = param_len[k] == 1 ? 0u : 1u;
  }
  const size_t out_bytes = rng->n_streams * n_samples * sizeof(double);
  rc = ensure(rng, &rng->d_out, &rng->out_cap, out_bytes);
  if (rc != MB200_OK) return rc;
  rc = reset_err(rng);
  if (rc != MB200_OK) return rc;
  a.out = rng->d_out;
  a.n_streams = rng->n_streams;
  a.n_samples = n_samples;
  a.err = rng->d_err;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CU(launch_deterministic(dist, np, real_type, a, rng->stream));
  CU(cudaMemcpyAsync(out_host, rng->d_out, out_bytes, cudaMemcpyDeviceToHost, rng->stream));
  rc = fetch_err(rng);
  if (rc != MB200_OK) return rc;
  if (rng->h_err->stream != ~0ull) {
    if (rng->h_err->code == kErrParam && dist != MB200_DIST_BINOMIAL) {
      return param_error(rng, dist, real_type, params_host, param_len, false, rng->h_err->stream);
    }
    if (rng->h_err->code == kErrHypergeometricDet) {
      // hypergeometric_deterministic validates and prints the UNROUNDED inputs (%.0f)
      double v[3];
      for (int k = 0; k < 3; ++k) v[k] = params_host[k][param_len[k] == 1 ? 0 : rng->h_err->stream];
      if (real_type == MB200_REAL_F32) for (double& x : v) x = static_cast<float>(x);
      return fail(MB200_ERR_PARAM, rng, "Invalid call to hypergeometric with n1 = %.0f, n2 = %.0f, k = %.0f",
                  v[0], v[1], v[2]);
    }
    const std::string msg = format_dynamic_error(rng->h_err->code == kErrParam ? kErrInnerBinomial : rng->h_err->code,
                                                 rng->h_err->vals);
    return fail(MB200_ERR_PARAM, rng, "%s", msg.c_str());
  }
  return MB200_OK;
}

}  // namespace

extern "C" {

const char* mb200_version(void) { return "monty_b200 0.1.0 (reference monty 0.4.14)"; }

int mb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* mb200_last_error(void) { return t_last_error.c_str(); }
uint64_t mb200_launch_count(void) { return g_launches.load(); }

// ---------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------
int mb200_rng_create(uint64_t seed, const uint8_t* seed_state_host, size_t seed_state_len,
                     size_t n_streams, int deterministic, uint64_t stream_offset,
                     uint64_t long_jumps, int device, mb200_rng** out) {
  mb200_rng* rng = nullptr;   // for the macros
  if (!out) return fail(MB200_ERR_ARG, nullptr, "out must not be NULL");
  *out = nullptr;
  if (seed_state_len % 32 != 0) {
    // ref: inst/include/monty/r/random.hpp:19-30
    return fail(MB200_ERR_STATE, nullptr, "Expected raw vector of length as multiple of 32 for 'seed'");
  }
  if (n_streams == 0) return fail(MB200_ERR_ARG, nullptr, "n_streams must be at least 1");
  if (mb200_device_count() <= device || device < 0) {
    return fail(MB200_ERR_CUDA, nullptr,
                "no usable CUDA device %d (monty_b200 has no CPU fallback)", device);
  }
  const JumpAlgebra* alg = nullptr;
  try { alg = &algebra(); } catch (const std::exception& e) {
    return fail(MB200_ERR_STATE, nullptr, "%s", e.what());
  }
  DeviceGuard device_guard;
  CU(cudaSetDevice(device));
  // mb200_rng_free releases whatever has been allocated when a step below fails
  std::unique_ptr<mb200_rng, void (*)(mb200_rng*)> h(new mb200_rng(), mb200_rng_free);
  rng = h.get();
  rng->device = device;
  rng->n_streams = n_streams;
  rng->deterministic = deterministic != 0;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  rng->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&rng->stream, cudaStreamNonBlocking));
  CU(ensure_lgamma_table(device, rng->stream));
  CU(cudaMalloc(&rng->d_state, n_streams * 32));
  CU(cudaMalloc(&rng->d_err, sizeof(ErrRec)));
  CU(cudaMallocHost(&rng->h_err, sizeof(ErrRec)));

  // The seed states, as the reference's seed vector (prng.hpp:38-53): k full
  // states, or the splitmix64 expansion of the integer seed
  // (generator.hpp:114-141).
  std::vector<uint64_t> seeds;
  if (seed_state_len > 0) {
    seeds.resize(seed_state_len / 8);
    std::memcpy(seeds.data(), seed_state_host, seed_state_len);
  } else {
    seeds.resize(4);
    uint64_t s = seed;
    for (int i = 0; i < 4; ++i) seeds[i] = s = host_splitmix64(s);   // generator.hpp:118-121
  }
  const size_t n_seed = seeds.size() / 4;
  if (long_jumps > 0) {
    const Poly256 lp = alg->long_jumps(long_jumps);
    for (size_t i = 0; i < n_seed; ++i) JumpAlgebra::apply(lp, &seeds[4 * i]);
  }
  // local stream i is global stream g = stream_offset + i:
  //   g <  n_seed: the supplied state g
  //   g >= n_seed: (g - (n_seed - 1)) jumps from the last supplied state
  size_t n_copied = 0;
  if (stream_offset < n_seed) {
    n_copied = std::min<size_t>(n_seed - stream_offset, n_streams);
    CU(cudaMemcpyAsync(rng->d_state, &seeds[4 * stream_offset], n_copied * 32,
                       cudaMemcpyHostToDevice, rng->stream));
  }
  if (n_copied < n_streams) {
    size_t anchor;   // local index of the stream the ladder grows from
    if (n_copied > 0) {
      anchor = n_copied - 1;   // the last supplied state
    } else {
      uint64_t a[4];
      std::memcpy(a, &seeds[4 * (n_seed - 1)], 32);
      JumpAlgebra::apply(alg->jumps(stream_offset - (n_seed - 1)), a);
      CU(cudaMemcpyAsync(rng->d_state, a, 32, cudaMemcpyHostToDevice, rng->stream));
      CU(cudaStreamSynchronize(rng->stream));   // `a` is a stack buffer
      anchor = 0;
    }
    // doubling ladder: have = streams in place from the anchor on
    unsigned long long have = 1;
    const unsigned long long total = n_streams - anchor;
    Poly256 p = kJumpPoly;   // c^have
    while (have < total) {
      const unsigned long long n_new = std::min<unsigned long long>(have, total - have);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      CU(launch_ladder(rng->d_state, anchor, have, n_new, p.data(), rng->stream));
      have += n_new;
      if (have < total) p = alg->mul(p, p);
    }
  }
  CU(cudaStreamSynchronize(rng->stream));
  *out = h.release();
  return MB200_OK;
}

void mb200_rng_free(mb200_rng* rng) {
  if (!rng) return;
  DeviceGuard device_guard;
  cudaSetDevice(rng->device);
  if (rng->stream) cudaStreamSynchronize(rng->stream);
  cudaFree(rng->d_state);
  cudaFree(rng->d_err);
  if (rng->h_err) cudaFreeHost(rng->h_err);
  for (int k = 0; k < MB200_MAX_PARAMS; ++k) cudaFree(rng->d_par[k]);
  cudaFree(rng->d_out);
  for (int s = 0; s < 3; ++s) {
    if (rng->stage[s]) cudaFreeHost(rng->stage[s]);
    if (rng->stage_ev[s]) cudaEventDestroy(rng->stage_ev[s]);
  }
  if (rng->stream) cudaStreamDestroy(rng->stream);
  delete rng;
}

size_t mb200_rng_n_streams(const mb200_rng* rng) { return rng ? rng->n_streams : 0; }
int mb200_rng_deterministic(const mb200_rng* rng) { return rng && rng->deterministic; }
int mb200_rng_device(const mb200_rng* rng) { return rng ? rng->device : -1; }
const char* mb200_rng_last_error(const mb200_rng* rng) { return rng ? rng->last_error.c_str() : ""; }
uint64_t* mb200_rng_state_dev(mb200_rng* rng) { return rng ? rng->d_state : nullptr; }
void* mb200_rng_cuda_stream(mb200_rng* rng) { return rng ? static_cast<void*>(rng->stream) : nullptr; }

int mb200_rng_state(mb200_rng* rng, uint8_t* out_host, size_t len) {
  if (!rng || !out_host) return fail(MB200_ERR_ARG, rng, "NULL argument");
  if (len != rng->n_streams * 32) {
    return fail(MB200_ERR_STATE, rng, "state buffer must hold %d bytes (but was %d)",
                (int)(rng->n_streams * 32), (int)len);
  }
  DeviceGuard device_guard;
  CU(cudaSetDevice(rng->device));
  CU(cudaMemcpyAsync(out_host, rng->d_state, len, cudaMemcpyDeviceToHost, rng->stream));
  CU(cudaStreamSynchronize(rng->stream));
  return MB200_OK;
}

int mb200_rng_set_state(mb200_rng* rng, const uint8_t* value_host, size_t len) {
  if (!rng || !value_host) return fail(MB200_ERR_ARG, rng, "NULL argument");
  if (len != rng->n_streams * 32) {
    // ref: src/random.cpp:372-375
    return fail(MB200_ERR_STATE, rng, "'value' must be a raw vector of length %d (but was %d)",
                (int)(rng->n_streams * 32), (int)len);
  }
  DeviceGuard device_guard;
  CU(cudaSetDevice(rng->device));
  CU(cudaMemcpyAsync(rng->d_state, value_host, len, cudaMemcpyHostToDevice, rng->stream));
  CU(cudaStreamSynchronize(rng->stream));
  return MB200_OK;
}

static int jump_all(mb200_rng* rng, int n, bool is_long) {
  if (!rng) return fail(MB200_ERR_ARG, rng, "NULL argument");
  if (n < 0) return fail(MB200_ERR_ARG, rng, "'n' must be non-negative");
  if (n == 0) return MB200_OK;
  DeviceGuard device_guard;
  CU(cudaSetDevice(rng->device));
  const Poly256 p = is_long ? algebra().long_jumps(n) : algebra().jumps(n);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CU(launch_apply_poly(rng->d_state, 0, rng->n_streams, p.data(), rng->stream));
  return MB200_OK;
}
int mb200_rng_jump(mb200_rng* rng, int n) { return jump_all(rng, n, false); }
int mb200_rng_long_jump(mb200_rng* rng, int n) { return jump_all(rng, n, true); }

int mb200_rng_sync(mb200_rng* rng) {
  if (!rng) return fail(MB200_ERR_ARG, rng, "NULL argument");
  DeviceGuard device_guard;
  CU(cudaSetDevice(rng->device));
  if (!rng->pending) { CU(cudaStreamSynchronize(rng->stream)); return MB200_OK; }
  rng->pending = false;
  int rc = fetch_err(rng);
  if (rc != MB200_OK) return rc;
  if (rng->h_err->stream == ~0ull) return MB200_OK;
  if (rng->h_err->code == kErrParam) {
    return param_error(rng, rng->pending_dist, rng->pending_rt, rng->pending_par,
                       rng->pending_len, true, rng->h_err->stream);
  }
  const std::string msg = format_dynamic_error(rng->h_err->code, rng->h_err->vals);
  return fail(MB200_ERR_PARAM, rng, "%s", msg.c_str());
}

// ---------------------------------------------------------------------------
// sampling
// ---------------------------------------------------------------------------
int mb200_rng_sample_dev(mb200_rng* rng, int dist, int real_type, size_t n_samples,
                         const double* const* params_dev, const size_t* param_len,
                         void* out_dev, int out_f32) {
  if (!rng) return fail(MB200_ERR_ARG, rng, "NULL argument");
  const int np = dist_n_params(dist);
  if (np < 0) return fail(MB200_ERR_ARG, rng, "unknown distribution id %d", dist);
  if (real_type != MB200_REAL_F64 && real_type != MB200_REAL_F32) {
    return fail(MB200_ERR_ARG, rng, "unknown real_type %d", real_type);
  }
  if (out_f32 && real_type != MB200_REAL_F32) {
    return fail(MB200_ERR_ARG, rng, "float output needs real_type float");
  }
  if (rng->deterministic) {
    return fail(MB200_ERR_UNSUPPORTED, rng, "deterministic mode is not available on the device-resident path");
  }
  if (n_samples == 0) return MB200_OK;
  if (!out_dev) return fail(MB200_ERR_ARG, rng, "out_dev must not be NULL");
  int rc = check_param_lengths(rng, dist, param_len);
  if (rc != MB200_OK) return rc;
  DeviceGuard device_guard;
  CU(cudaSetDevice(rng->device));
  if (rng->pending) { rc = mb200_rng_sync(rng); if (rc != MB200_OK) return rc; }
  SampleArgs a{};
  a.state = rng->d_state;
  for (int k = 0; k < np; ++k) {
    a.par[k] = params_dev[k];
    a.pstride[k] = param_len[k] == 1 ? 0u : 1u;
    rng->pending_par[k] = params_dev[k];
    rng->pending_len[k] = param_len[k];
  }
  a.out = out_dev;
  a.n_streams = rng->n_streams;
  a.n_samples = n_samples;
  a.err = rng->d_err;
  rc = reset_err(rng);
  if (rc != MB200_OK) return rc;
  CU(launch_sample(dist, real_type, out_f32, false, a, rng->sm_count, rng->stream));
  rng->pending = true;
  rng->pending_dist = dist;
  rng->pending_rt = real_type;
  return MB200_OK;
}

int mb200_rng_sample(mb200_rng* rng, int dist, int real_type, size_t n_samples,
                     const double* const* params_host, const size_t* param_len,
                     double* out_host) {
  if (!rng) return fail(MB200_ERR_ARG, rng, "NULL argument");
  const int np = dist_n_params(dist);
  if (np < 0) return fail(MB200_ERR_ARG, rng, "unknown distribution id %d", dist);
  if (real_type != MB200_REAL_F64 && real_type != MB200_REAL_F32) {
    return fail(MB200_ERR_ARG, rng, "unknown real_type %d", real_type);
  }
  int rc = check_param_lengths(rng, dist, param_len);
  if (rc != MB200_OK) return rc;
  if (n_samples == 0) return MB200_OK;
  if (!out_host) return fail(MB200_ERR_ARG, rng, "out_host must not be NULL");
  DeviceGuard device_guard;
  CU(cudaSetDevice(rng->device));
  if (rng->pending) { rc = mb200_rng_sync(rng); if (rc != MB200_OK) return rc; }
  if (rng->deterministic) return sample_deterministic(rng, dist, real_type, n_samples, params_host, param_len, out_host);

  // stage parameters (R vectors live in host memory)
  SampleArgs a{};
  a.state = rng->d_state;
  bool all_scalar = true;
  for (int k = 0; k < np; ++k) {
    const size_t bytes = param_len[k] * sizeof(double);
    rc = ensure(rng, reinterpret_cast<void**>(&rng->d_par[k]), &rng->par_cap[k], bytes);
    if (rc != MB200_OK) return rc;
    rc = copy_from_host(rng, rng->d_par[k], params_host[k], bytes);
    if (rc != MB200_OK) return rc;
    a.par[k] = rng->d_par[k];
    a.pstride[k] = param_len[k] == 1 ? 0u : 1u;
    all_scalar = all_scalar && param_len[k] == 1;
  }
  a.err = rng->d_err;

  // The reference throws at the first stream whose parameters are invalid,
  // after advancing every earlier stream (src/random.cpp:211-228).  A cheap
  // validate-only pass finds that stream so the sampling launch can stop
  // there.
  unsigned long long n_active = rng->n_streams;
  bool have_offender = false;
  if (dist_validates(dist)) {
    rc = reset_err(rng);
    if (rc != MB200_OK) return rc;
    a.n_streams = all_scalar ? 1 : rng->n_streams;
    a.n_samples = 0;
    CU(launch_sample(dist, real_type, 0, true, a, rng->sm_count, rng->stream));
    rc = fetch_err(rng);
    if (rc != MB200_OK) return rc;
    if (rng->h_err->stream != ~0ull) {
      have_offender = true;
      n_active = rng->h_err->stream;   // 0 when all parameters are scalars
    }
  }

  int result = MB200_OK;
  if (n_active > 0) {
    const size_t out_bytes = static_cast<size_t>(n_active) * n_samples * sizeof(double);
    rc = ensure(rng, &rng->d_out, &rng->out_cap, out_bytes);
    if (rc != MB200_OK) return rc;
    rc = reset_err(rng);
    if (rc != MB200_OK) return rc;
    a.out = rng->d_out;
    a.n_streams = n_active;
    a.n_samples = n_samples;
    CU(launch_sample(dist, real_type, 0, false, a, rng->sm_count, rng->stream));
    rc = copy_to_host(rng, out_host, rng->d_out, out_bytes);
    if (rc != MB200_OK) return rc;
    rc = fetch_err(rng);
    if (rc != MB200_OK) return rc;
    if (rng->h_err->stream != ~0ull && rng->h_err->stream < n_active) {
      // a data-dependent failure inside a composition precedes the offender
      const std::string msg = format_dynamic_error(rng->h_err->code, rng->h_err->vals);
      return fail(MB200_ERR_PARAM, rng, "%s", msg.c_str());
    }
  }
  if (have_offender) {
    result = param_error(rng, dist, real_type, params_host, param_len, false, n_active);
  }
  return result;
}

int mb200_rng_multinomial(mb200_rng* rng, size_t n_samples,
                          const double* size_host, size_t size_len,
                          const double* prob_host, size_t prob_len, size_t prob_cols,
                          double* out_host) {
  if (!rng) return fail(MB200_ERR_ARG, rng, "NULL argument");
  if (size_len != 1 && size_len != rng->n_streams) {
    if (rng->n_streams == 1) {
      return fail(MB200_ERR_ARG, rng, "Expected 'size' to have length %d, not %d", 1, (int)size_len);
    }
    return fail(MB200_ERR_ARG, rng, "Expected 'size' to have length %d or 1, not %d",
                (int)rng->n_streams, (int)size_len);
  }
  if (prob_cols != 1 && prob_cols != rng->n_streams) {
    // ref: src/random.cpp:61-69 (input_array)
    if (rng->n_streams == 1) {
      return fail(MB200_ERR_ARG, rng, "Expected 'prob' to have 1 column, not %d", (int)prob_cols);
    }
    return fail(MB200_ERR_ARG, rng, "Expected 'prob' to have %d or 1 columns, not %d",
                (int)rng->n_streams, (int)prob_cols);
  }
  if (n_samples == 0 || prob_len == 0) return MB200_OK;
  DeviceGuard device_guard;
  CU(cudaSetDevice(rng->device));
  int rc;
  if (rng->pending) { rc = mb200_rng_sync(rng); if (rc != MB200_OK) return rc; }
  rc = ensure(rng, reinterpret_cast<void**>(&rng->d_par[0]), &rng->par_cap[0], size_len * 8);
  if (rc != MB200_OK) return rc;
  rc = ensure(rng, reinterpret_cast<void**>(&rng->d_par[1]), &rng->par_cap[1], prob_len * prob_cols * 8);
  if (rc != MB200_OK) return rc;
  const size_t out_bytes = prob_len * n_samples * rng->n_streams * sizeof(double);
  rc = ensure(rng, &rng->d_out, &rng->out_cap, out_bytes);
  if (rc != MB200_OK) return rc;
  CU(cudaMemcpyAsync(rng->d_par[0], size_host, size_len * 8, cudaMemcpyHostToDevice, rng->stream));
  CU(cudaMemcpyAsync(rng->d_par[1], prob_host, prob_len * prob_cols * 8, cudaMemcpyHostToDevice, rng->stream));
  rc = reset_err(rng);
  if (rc != MB200_OK) return rc;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (rng->deterministic) {
    CU(launch_multinomial_deterministic(rng->n_streams, n_samples, rng->d_par[0],
                                        size_len == 1 ? 0u : 1u, rng->d_par[1], static_cast<int>(prob_len),
                                        prob_cols == 1 ? 0u : 1u, static_cast<double*>(rng->d_out),
                                        rng->d_err, rng->stream));
  } else {
    CU(launch_multinomial(rng->d_state, rng->n_streams, n_samples, rng->d_par[0],
                          size_len == 1 ? 0u : 1u, rng->d_par[1], static_cast<int>(prob_len),
                          prob_cols == 1 ? 0u : 1u, static_cast<double*>(rng->d_out),
                          rng->d_err, rng->stream));
  }
  CU(cudaMemcpyAsync(out_host, rng->d_out, out_bytes, cudaMemcpyDeviceToHost, rng->stream));
  rc = fetch_err(rng);
  if (rc != MB200_OK) return rc;
  if (rng->h_err->stream != ~0ull) {
    // ref: multinomial.hpp:60-66
    if (rng->h_err->code == 5) return fail(MB200_ERR_PARAM, rng, "Negative prob passed to multinomial");
    if (rng->h_err->code == 6) return fail(MB200_ERR_PARAM, rng, "No positive prob in call to multinomial");
    const std::string msg = format_dynamic_error(rng->h_err->code, rng->h_err->vals);
    return fail(MB200_ERR_PARAM, rng, "%s", msg.c_str());
  }
  return MB200_OK;
}

int mb200_rng_raw(mb200_rng* rng, int gen, size_t n_samples, uint64_t* out_host) {
  if (!rng || !out_host) return fail(MB200_ERR_ARG, rng, "NULL argument");
  if (gen < MB200_GEN_XOSHIRO256PLUS || gen > MB200_GEN_XOSHIRO256STARSTAR) {
    return fail(MB200_ERR_ARG, rng, "raw: generator must be a xoshiro256 variant");
  }
  if (n_samples == 0) return MB200_OK;
  DeviceGuard device_guard;
  CU(cudaSetDevice(rng->device));
  const size_t bytes = rng->n_streams * n_samples * sizeof(uint64_t);
  int rc = ensure(rng, &rng->d_out, &rng->out_cap, bytes);
  if (rc != MB200_OK) return rc;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CU(launch_raw(gen, rng->d_state, rng->n_streams, n_samples, static_cast<uint64_t*>(rng->d_out), rng->stream));
  CU(cudaMemcpyAsync(out_host, rng->d_out, bytes, cudaMemcpyDeviceToHost, rng->stream));
  CU(cudaStreamSynchronize(rng->stream));
  return MB200_OK;
}

int mb200_xoshiro_run(int gen, uint64_t seed, int n, int device,
                      uint64_t* out_values_host, uint64_t* out_state_host, int* out_state_words) {
  mb200_rng* rng = nullptr;
  if (!out_values_host || !out_state_host || !out_state_words || n <= 0) {
    return fail(MB200_ERR_ARG, nullptr, "bad argument");
  }
  if (gen < 0 || gen >= MB200_GEN_COUNT) return fail(MB200_ERR_ARG, nullptr, "unknown generator id %d", gen);
  if (mb200_device_count() <= device || device < 0) {
    return fail(MB200_ERR_CUDA, nullptr, "no usable CUDA device %d (monty_b200 has no CPU fallback)", device);
  }
  DeviceGuard device_guard;
  CU(cudaSetDevice(device));
  uint64_t* d = nullptr;
  CU(cudaMalloc(&d, (3 * n + 8) * sizeof(uint64_t)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = launch_xoshiro_run(gen, seed, n, d, d + 3 * n, out_state_words, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(out_values_host, d, 3 * n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_state_host, d + 3 * n, *out_state_words * sizeof(uint64_t), cudaMemcpyDeviceToHost);
  cudaFree(d);
  CU(e);
  return MB200_OK;
}

// ---------------------------------------------------------------------------
// densities
// ---------------------------------------------------------------------------
static int density_n_params(int dens) {
  static const int np[MB200_DENS_COUNT] = {2, 2, 2, 2, 3, 3, 1, 2, 1, 1, 2, 2, 2, 3, 2, 4, 2, 2, 2, 3, 3};
  return (dens >= 0 && dens < MB200_DENS_COUNT) ? np[dens] : -1;
}

int mb200_density_dev(int dens, int real_type, size_t n,
                      const void* x_dev, int x_is_i32,
                      const void* const* params_dev, const int* param_is_i32,
                      int log, double* out_dev, double* sum_dev,
                      int device, void* cuda_stream) {
  mb200_rng* rng = nullptr;
  const int np = density_n_params(dens);
  if (np < 0) return fail(MB200_ERR_ARG, nullptr, "unknown density id %d", dens);
  if (real_type != MB200_REAL_F64 && real_type != MB200_REAL_F32) {
    return fail(MB200_ERR_ARG, nullptr, "unknown real_type %d", real_type);
  }
  DeviceGuard device_guard;
  CU(cudaSetDevice(device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  CU(ensure_lgamma_table(device, st));
  if (n == 0) {
    if (sum_dev) CU(cudaMemsetAsync(sum_dev, 0, sizeof(double), st));
    return MB200_OK;
  }
  cudaDeviceProp prop;
  static int sm_cache[64] = {0};
  if (device < 64 && sm_cache[device] == 0) {
    CU(cudaGetDeviceProperties(&prop, device));
    sm_cache[device] = prop.multiProcessorCount;
  }
  const int sm = device < 64 ? sm_cache[device] : 148;
  DensityArgs a{};
  a.x = x_dev; a.x_is_i32 = x_is_i32;
  for (int k = 0; k < np; ++k) { a.par[k] = params_dev[k]; a.par_is_i32[k] = param_is_i32 ? param_is_i32[k] : 0; }
  a.out = out_dev;
  a.n = n;
  a.log = log;
  int grid = density_max_partials(sm);
  double* partial = nullptr;
  bool pooled = false;
  unsigned slot = 0;
  if (sum_dev) {
    if (device < 64 && g_partials[device].base && grid <= kPartialSlotDoubles) {
      slot = g_partials[device].next.fetch_add(1, std::memory_order_relaxed) % kPartialSlots;
      partial = g_partials[device].base + static_cast<size_t>(slot) * kPartialSlotDoubles;
      pooled = true;
      if (!g_partials[device].used[slot]) {
        CU(cudaEventCreateWithFlags(&g_partials[device].used[slot], cudaEventDisableTiming));
      } else {
        CU(cudaStreamWaitEvent(st, g_partials[device].used[slot], 0));
      }
    } else {
      CU(cudaMallocAsync(&partial, grid * sizeof(double), st));
    }
    a.partial = partial;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CU(launch_density(dens, real_type, a, sm, &grid, st));
  if (sum_dev) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CU(launch_reduce_partials(partial, grid, sum_dev, st));
    if (pooled) CU(cudaEventRecord(g_partials[device].used[slot], st));
    else CU(cudaFreeAsync(partial, st));
  }
  return MB200_OK;
}

int mb200_density(int dens, int real_type, size_t n,
                  const void* x_host, int x_is_i32,
                  const void* const* params_host, const int* param_is_i32,
                  int log, double* out_host, double* sum_out, int device) {
  mb200_rng* rng = nullptr;
  const int np = density_n_params(dens);
  if (np < 0) return fail(MB200_ERR_ARG, nullptr, "unknown density id %d", dens);
  if (mb200_device_count() <= device || device < 0) {
    return fail(MB200_ERR_CUDA, nullptr, "no usable CUDA device %d (monty_b200 has no CPU fallback)", device);
  }
  DeviceGuard device_guard;
  CU(cudaSetDevice(device));
  CU(ensure_lgamma_table(device, nullptr));
  if (n == 0) { if (sum_out) *sum_out = 0.0; return MB200_OK; }
  // Device buffers, stream and pinned staging ring are kept per device and grow
  // on demand (a likelihood evaluated in an MCMC loop calls this thousands of
  // times with the same lengths: no cudaMalloc / cudaFree per call); the calls of
  // one device are serialised by the scratch's mutex.
  DensityScratch* sc = nullptr;
  int rc = density_scratch(device, &sc);
  if (rc != MB200_OK) return rc;
  std::lock_guard<std::mutex> lock(sc->m);
  rng = sc->io;                               // transfers and error text go through this pseudo-handle
  void* d_p[MB200_MAX_PARAMS] = {nullptr, nullptr, nullptr, nullptr};
  const size_t xb = n * (x_is_i32 ? sizeof(int32_t) : sizeof(double));
  rc = ensure(rng, &sc->x, &sc->x_cap, xb);
  if (rc == MB200_OK) rc = copy_from_host(rng, sc->x, x_host, xb);
  for (int k = 0; k < np && rc == MB200_OK; ++k) {
    const size_t pb = n * ((param_is_i32 && param_is_i32[k]) ? sizeof(int32_t) : sizeof(double));
    rc = ensure(rng, &sc->p[k], &sc->p_cap[k], pb);
    if (rc == MB200_OK) rc = copy_from_host(rng, sc->p[k], params_host[k], pb);
    d_p[k] = sc->p[k];
  }
  if (rc == MB200_OK && out_host) rc = ensure(rng, &sc->out, &sc->out_cap, n * sizeof(double));
  if (rc != MB200_OK) return rc;
  rc = mb200_density_dev(dens, real_type, n, sc->x, x_is_i32, d_p, param_is_i32, log,
                         out_host ? static_cast<double*>(sc->out) : nullptr, sum_out ? sc->sum : nullptr, device,
                         rng->stream);
  if (rc != MB200_OK) return rc;
  if (out_host) {
    rc = copy_to_host(rng, out_host, sc->out, n * sizeof(double));
    if (rc != MB200_OK) return rc;
  }
  if (sum_out) CU(cudaMemcpyAsync(sc->h_sum, sc->sum, sizeof(double), cudaMemcpyDeviceToHost, rng->stream));
  CU(cudaStreamSynchronize(rng->stream));
  if (sum_out) *sum_out = *sc->h_sum;
  return MB200_OK;
}

// ---------------------------------------------------------------------------
// named entry points, 1:1 with the reference's registered routines
// ---------------------------------------------------------------------------
#define MB200_DEF_SAMPLER0(name, id) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, double* out_host) { \
    return mb200_rng_sample(rng, id, MB200_REAL_F64, n_samples, nullptr, nullptr, out_host); }
#define MB200_DEF_SAMPLER1(name, id, a) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, const double* a, size_t a##_len, \
                          double* out_host) { \
    const double* p[1] = {a}; const size_t l[1] = {a##_len}; \
    return mb200_rng_sample(rng, id, MB200_REAL_F64, n_samples, p, l, out_host); }
#define MB200_DEF_SAMPLER2(name, id, a, b) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, const double* a, size_t a##_len, \
                          const double* b, size_t b##_len, double* out_host) { \
    const double* p[2] = {a, b}; const size_t l[2] = {a##_len, b##_len}; \
    return mb200_rng_sample(rng, id, MB200_REAL_F64, n_samples, p, l, out_host); }
#define MB200_DEF_SAMPLER3(name, id, a, b, c) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, const double* a, size_t a##_len, \
                          const double* b, size_t b##_len, const double* c, size_t c##_len, \
                          double* out_host) { \
    const double* p[3] = {a, b, c}; const size_t l[3] = {a##_len, b##_len, c##_len}; \
    return mb200_rng_sample(rng, id, MB200_REAL_F64, n_samples, p, l, out_host); }
#define MB200_DEF_SAMPLER4(name, id, a, b, c, d) \
  int mb200_random_##name(mb200_rng* rng, size_t n_samples, const double* a, size_t a##_len, \
                          const double* b, size_t b##_len, const double* c, size_t c##_len, \
                          const double* d, size_t d##_len, double* out_host) { \
    const double* p[4] = {a, b, c, d}; const size_t l[4] = {a##_len, b##_len, c##_len, d##_len}; \
    return mb200_rng_sample(rng, id, MB200_REAL_F64, n_samples, p, l, out_host); }

MB200_DEF_SAMPLER0(real, MB200_DIST_REAL)
MB200_DEF_SAMPLER1(exponential_rate, MB200_DIST_EXPONENTIAL_RATE, rate)
MB200_DEF_SAMPLER1(exponential_mean, MB200_DIST_EXPONENTIAL_MEAN, mean)
MB200_DEF_SAMPLER1(poisson, MB200_DIST_POISSON, lambda)
MB200_DEF_SAMPLER2(beta, MB200_DIST_BETA, a, b)
MB200_DEF_SAMPLER2(binomial, MB200_DIST_BINOMIAL, size, prob)
MB200_DEF_SAMPLER2(cauchy, MB200_DIST_CAUCHY, location, scale)
MB200_DEF_SAMPLER2(gamma_scale, MB200_DIST_GAMMA_SCALE, shape, scale)
MB200_DEF_SAMPLER2(gamma_rate, MB200_DIST_GAMMA_RATE, shape, rate)
MB200_DEF_SAMPLER2(negative_binomial_prob, MB200_DIST_NEGATIVE_BINOMIAL_PROB, size, prob)
MB200_DEF_SAMPLER2(negative_binomial_mu, MB200_DIST_NEGATIVE_BINOMIAL_MU, size, mu)
MB200_DEF_SAMPLER2(normal_box_muller, MB200_DIST_NORMAL_BOX_MULLER, mean, sd)
MB200_DEF_SAMPLER2(normal_polar, MB200_DIST_NORMAL_POLAR, mean, sd)
MB200_DEF_SAMPLER2(normal_ziggurat, MB200_DIST_NORMAL_ZIGGURAT, mean, sd)
MB200_DEF_SAMPLER2(uniform, MB200_DIST_UNIFORM, min, max)
MB200_DEF_SAMPLER2(weibull, MB200_DIST_WEIBULL, shape, scale)
MB200_DEF_SAMPLER2(log_normal, MB200_DIST_LOG_NORMAL, meanlog, sdlog)
MB200_DEF_SAMPLER2(zi_poisson, MB200_DIST_ZI_POISSON, pi0, lambda)
MB200_DEF_SAMPLER3(beta_binomial_prob, MB200_DIST_BETA_BINOMIAL_PROB, size, prob, rho)
MB200_DEF_SAMPLER3(beta_binomial_ab, MB200_DIST_BETA_BINOMIAL_AB, size, a, b)
MB200_DEF_SAMPLER3(hypergeometric, MB200_DIST_HYPERGEOMETRIC, n1, n2, k)
MB200_DEF_SAMPLER3(zi_negative_binomial_prob, MB200_DIST_ZI_NEGATIVE_BINOMIAL_PROB, pi0, size, prob)
MB200_DEF_SAMPLER3(zi_negative_binomial_mu, MB200_DIST_ZI_NEGATIVE_BINOMIAL_MU, pi0, size, mu)
MB200_DEF_SAMPLER4(truncated_normal, MB200_DIST_TRUNCATED_NORMAL, mean, sd, min, max)

}  // extern "C"
