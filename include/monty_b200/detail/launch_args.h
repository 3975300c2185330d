// launch_args.h -- plain structs shared by host code and kernels.
#pragma once
#include <cstdint>

namespace mb200 {

// ---- error codes a sampler can raise (formatted on the host with the
//      reference's message text, see errors.h) --------------------------------
enum SampleError {
  kOk = 0,
  kErrParam = 1,        // the distribution's own validate() failed
  kErrInnerBinomial = 2,// beta_binomial: binomial_validate on the beta draw
  kErrInnerPoisson = 3, // negative_binomial: poisson_validate on the gamma draw
  kErrCauchyDeterministic = 4,
  kErrMultinomialNegative = 5,
  kErrMultinomialNoPositive = 6,
  kErrBinomialDetN = 7,      // binomial_deterministic: n < 0 beyond round-off
  kErrHypergeometricDet = 8  // hypergeometric_deterministic validates the unrounded inputs
};

// First-offender record written by kernels when a sampler rejects its
// parameters (device code cannot throw; the host turns this into the
// reference's exception text, see errors.h).
struct ErrRec {
  unsigned long long stream;   // smallest offending stream index, ~0ull if none
  unsigned long long owner;    // stream that wrote code/vals (== stream when consistent)
  int code;                    // SampleError
  unsigned work;               // tile counter of the self-scheduling kernels (starts at ~0u)
  double vals[3];              // values for data-dependent messages
  unsigned lock;               // ~0u: free (the record is reset with a 0xFF fill); guards code / vals / owner
  unsigned pad;
};

struct SampleArgs {
  uint64_t* state;             // [n_streams][4]
  const double* par[4];        // device pointers
  unsigned pstride[4];         // 0: broadcast scalar, 1: one value per stream
  void* out;                   // OutT[n_streams][n_samples]
  unsigned long long n_streams;  // streams processed by this launch
  unsigned long long n_samples;
  ErrRec* err;
};

struct DensityArgs {
  const void* x; int x_is_i32;
  const void* par[4]; int par_is_i32[4];
  double* out;                 // may be null (reduce only)
  double* partial;             // per-block partial sums (null: no reduction)
  unsigned long long n;
  int log;
};

}  // namespace mb200
