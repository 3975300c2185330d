// multinomial_kernel.cu -- multinomial draws as chained conditional binomials.
//
// ref: inst/include/monty/random/multinomial.hpp:53-82 (algorithm),
//      src/random.cpp:884-934 (loops, layout [len x n_samples x n_streams]).
// One thread per stream; the probabilities are a shared vector (prob_cols == 1)
// or one column per stream.  Always real_type = double, like the R interface.
#include "../../include/monty_b200.h"
#include "multinomial_dispatch.h"
#include "sample_kernel.cuh"

namespace mb200 {

__global__ void __launch_bounds__(kBlock)
multinomial_kernel(uint64_t* __restrict__ state, unsigned long long n_streams,
                   unsigned long long n_samples,
                   const double* __restrict__ size, unsigned size_stride,
                   const double* __restrict__ prob, int prob_len, unsigned prob_stride,
                   double* __restrict__ out, ErrRec* err) {
  __shared__ TableSmem ts;
  Ctx ctx;
  ctx.err = 0;
  stage_tables<kTailTables>(&ts, ctx);
  const unsigned long long stream = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (stream >= n_streams) return;
  const double* pr = prob + static_cast<unsigned long long>(prob_stride) * stream * prob_len;
  // multinomial.hpp:58-67: validate and total once per call (the values do
  // not depend on the draw, so once per stream is the same arithmetic)
  double p_tot0 = 0;
  for (int i = 0; i < prob_len; ++i) {
    if (pr[i] < 0) { report_error(err, stream, 5, nullptr); return; }
    p_tot0 += pr[i];
  }
  if (p_tot0 == 0) { report_error(err, stream, 6, nullptr); return; }
  Rng rng;
  rng.s = load_state(state, stream);
  const double size0 = size[size_stride ? stream : 0];
  bool ok = true;
  for (unsigned long long j = 0; j < n_samples && ok; ++j) {
    double* y = out + static_cast<unsigned long long>(prob_len) * (n_samples * stream + j);
    double sz = size0;
    double p_tot = p_tot0;
    for (int i = 0; i < prob_len - 1; ++i) {
      const double p = pr[i];
      if (p > 0) {
        const double q = p / p_tot;
        const double pi = q < 1.0 ? q : 1.0;          // utils::min(a, b) = a < b ? a : b
        DistBinomial<double>::Prep bq;
        if (DistBinomial<double>::prepare2(sz, pi, bq, ctx) != kOk) {
          const double vals[3] = {round(sz), pi, 1 - pi};
          report_error(err, stream, kErrInnerBinomial, vals);
          ok = false;
          break;
        }
        const double draw = DistBinomial<double>::draw(rng, bq, ctx);
        y[i] = draw;
        sz -= draw;
        p_tot -= p;
      } else {
        y[i] = 0;
      }
    }
    if (ok) y[prob_len - 1] = sz;
  }
  store_state(state, stream, rng.s);
}

cudaError_t launch_multinomial(uint64_t* state, unsigned long long n_streams,
                               unsigned long long n_samples,
                               const double* size, unsigned size_stride,
                               const double* prob, int prob_len, unsigned prob_stride,
                               double* out, ErrRec* err, cudaStream_t st) {
  if (n_streams == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((n_streams + kBlock - 1) / kBlock);
  multinomial_kernel<<<grid, kBlock, 0, st>>>(state, n_streams, n_samples, size, size_stride,
                                             prob, prob_len, prob_stride, out, err);
  return cudaGetLastError();
}

}  // namespace mb200
