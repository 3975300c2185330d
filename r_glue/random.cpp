// r_glue/random.cpp -- replaces src/random.cpp of mrc-ide/monty: the same
// [[cpp11::register]] routines (ref: src/random.cpp:341-934, registered in
// src/cpp11.cpp:552-608), each one call into the C-ABI of libmonty_b200.so.
//
// What stays the reference's: the routine names and argument order (R/cpp11.R
// calls them positionally), the `preserve_stream_dimension` attribute and the
// dim() rules of the results (src/random.cpp:85-87, 197-209), the error
// messages (formatted inside the library with the reference's text:
// parameter-length errors src/random.cpp:21-42, distribution errors, partial
// advance of the streams before an offender), and the external-pointer
// life cycle (R's GC frees the device state through the finaliser).
#include <cpp11/doubles.hpp>
#include <cpp11/external_pointer.hpp>
#include <cpp11/integers.hpp>
#include <cpp11/raws.hpp>
#include <cpp11/sexp.hpp>
#include <cpp11/protect.hpp>

#include <cmath>
#include <cstdint>
#include <limits>

#include <monty_b200.h>

namespace {

// The generator: a handle on the device-resident stream states.
struct b200_rng {
  mb200_rng* h = nullptr;
  ~b200_rng() { mb200_rng_free(h); }
};

void check(int rc, const b200_rng* rng = nullptr) {
  if (rc != MB200_OK) cpp11::stop("%s", rng && rng->h ? mb200_rng_last_error(rng->h) : mb200_last_error());
}

// ref: src/random.cpp:12-19
b200_rng* read_ptr(cpp11::sexp ptr, const char* context) {
  if (!R_ExternalPtrAddr(ptr)) {
    cpp11::stop("Pointer has been serialised, cannot continue safely (%s)", context);
  }
  return cpp11::as_cpp<cpp11::external_pointer<b200_rng>>(ptr).get();
}

// ref: src/random.cpp:85-87
bool preserve_stream_dimension(size_t n_streams, cpp11::sexp ptr) {
  return n_streams > 1 || INTEGER(ptr.attr("preserve_stream_dimension"))[0] == 1;
}

// One draw per stream (monty_random_<dist>: n_samples == 0 in the C-ABI, a plain
// vector of length n_streams) or n_samples draws per stream
// (monty_random_n_<dist>: [n_samples x n_streams] when the stream dimension is
// kept).  `call` receives the handle, the C-ABI's n_samples and the output.
template <typename Call>
cpp11::doubles sample(cpp11::sexp ptr, const char* what, bool is_n, size_t n_samples, Call call) {
  b200_rng* rng = read_ptr(ptr, what);
  const size_t n_streams = mb200_rng_n_streams(rng->h);
  cpp11::writable::doubles r_y(static_cast<R_xlen_t>((is_n ? n_samples : 1) * n_streams));
  if (is_n && n_samples == 0) {
    // an empty result: nothing to draw (the C-ABI reads n_samples == 0 as "one draw")
  } else {
    check(call(rng->h, is_n ? n_samples : 0, REAL(r_y)), rng);
  }
  if (is_n && preserve_stream_dimension(n_streams, ptr)) {
    r_y.attr("dim") = cpp11::writable::integers{static_cast<int>(n_samples), static_cast<int>(n_streams)};
  }
  return r_y;
}

}  // namespace

// ---- life cycle, state, jumps (ref: src/random.cpp:341-396) ------------------

// Seed forms of as_rng_seed (inst/include/monty/r/random.hpp:47-67): a number
// (expanded by splitmix64 inside the library), a raw vector of whole generator
// states, or NULL (drawn from R's own generator).
[[cpp11::register]]
SEXP monty_rng_alloc(cpp11::sexp r_seed, int n_streams, bool deterministic) {
  uint64_t seed = 0;
  const uint8_t* raw = nullptr;
  size_t raw_len = 0;
  const auto type = TYPEOF(r_seed);
  if (type == INTSXP || type == REALSXP) {
    seed = cpp11::as_cpp<size_t>(r_seed);
  } else if (type == RAWSXP) {
    raw = reinterpret_cast<const uint8_t*>(RAW(r_seed));
    raw_len = static_cast<size_t>(Rf_xlength(r_seed));
    if (raw_len == 0 || raw_len % 32 != 0) {
      cpp11::stop("Expected raw vector of length as multiple of 32 for 'seed'");
    }
  } else if (type == NILSXP) {
    GetRNGstate();
    seed = static_cast<uint64_t>(std::ceil(std::abs(::unif_rand()) *
                                           static_cast<double>(std::numeric_limits<size_t>::max())));
    PutRNGstate();
  } else {
    cpp11::stop("Invalid type for 'seed'");
  }
  auto* rng = new b200_rng();
  const int rc = mb200_rng_create(seed, raw, raw_len, static_cast<size_t>(n_streams), deterministic ? 1 : 0,
                                  /*stream_offset=*/0, /*long_jumps=*/0, /*device=*/0, &rng->h);
  if (rc != MB200_OK) {
    delete rng;
    cpp11::stop("%s", mb200_last_error());
  }
  return cpp11::external_pointer<b200_rng>(rng);
}

[[cpp11::register]]
cpp11::sexp cpp_monty_rng_state(cpp11::sexp ptr) {
  b200_rng* rng = read_ptr(ptr, "state");
  const size_t len = 32 * mb200_rng_n_streams(rng->h);
  cpp11::writable::raws ret(static_cast<R_xlen_t>(len));
  check(mb200_rng_state(rng->h, reinterpret_cast<uint8_t*>(RAW(ret)), len), rng);
  return ret;
}

[[cpp11::register]]
void cpp_monty_rng_set_state(cpp11::sexp ptr, cpp11::raws r_value) {
  b200_rng* rng = read_ptr(ptr, "set_state");
  // a wrong length fails inside with "'value' must be a raw vector of length %d (but was %d)"
  check(mb200_rng_set_state(rng->h, reinterpret_cast<const uint8_t*>(RAW(r_value)),
                            static_cast<size_t>(r_value.size())), rng);
}

[[cpp11::register]]
void cpp_monty_rng_jump(cpp11::sexp ptr, int n) {
  b200_rng* rng = read_ptr(ptr, "jump");
  check(mb200_rng_jump(rng->h, n), rng);          // n jumps of every stream in one launch
}

[[cpp11::register]]
void cpp_monty_rng_long_jump(cpp11::sexp ptr, int n) {
  b200_rng* rng = read_ptr(ptr, "long_jump");
  check(mb200_rng_long_jump(rng->h, n), rng);
}

// ---- samplers (ref: src/random.cpp:398-881) -----------------------------------
// Each distribution registers monty_random_<dist>(params..., ptr) and
// monty_random_n_<dist>(n_samples, params..., ptr).  The C-ABI checks each
// parameter's length against the number of streams (1 or n_streams) and names
// it in the error exactly as the reference does.
#define MB200_ARG(x) REAL(x), static_cast<size_t>(x.size())

#define MB200_GLUE0(dist)                                                                      \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_##dist(cpp11::sexp ptr) {               \
    return sample(ptr, #dist, false, 1, [&](mb200_rng* h, size_t n, double* y) {              \
      return mb200_random_##dist(h, n, y); });                                                \
  }                                                                                            \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_n_##dist(size_t n_samples, cpp11::sexp ptr) { \
    return sample(ptr, #dist, true, n_samples, [&](mb200_rng* h, size_t n, double* y) {       \
      return mb200_random_##dist(h, n, y); });                                                \
  }
#define MB200_GLUE1(dist, a)                                                                   \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_##dist(cpp11::doubles a, cpp11::sexp ptr) { \
    return sample(ptr, #dist, false, 1, [&](mb200_rng* h, size_t n, double* y) {              \
      return mb200_random_##dist(h, n, MB200_ARG(a), y); });                                  \
  }                                                                                            \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_n_##dist(size_t n_samples, cpp11::doubles a, \
                                                               cpp11::sexp ptr) {             \
    return sample(ptr, #dist, true, n_samples, [&](mb200_rng* h, size_t n, double* y) {       \
      return mb200_random_##dist(h, n, MB200_ARG(a), y); });                                  \
  }
#define MB200_GLUE2(dist, a, b)                                                                \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_##dist(cpp11::doubles a, cpp11::doubles b, \
                                                             cpp11::sexp ptr) {               \
    return sample(ptr, #dist, false, 1, [&](mb200_rng* h, size_t n, double* y) {              \
      return mb200_random_##dist(h, n, MB200_ARG(a), MB200_ARG(b), y); });                    \
  }                                                                                            \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_n_##dist(size_t n_samples, cpp11::doubles a, \
                                                               cpp11::doubles b, cpp11::sexp ptr) { \
    return sample(ptr, #dist, true, n_samples, [&](mb200_rng* h, size_t n, double* y) {       \
      return mb200_random_##dist(h, n, MB200_ARG(a), MB200_ARG(b), y); });                    \
  }
#define MB200_GLUE3(dist, a, b, c)                                                             \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_##dist(cpp11::doubles a, cpp11::doubles b, \
                                                             cpp11::doubles c, cpp11::sexp ptr) { \
    return sample(ptr, #dist, false, 1, [&](mb200_rng* h, size_t n, double* y) {              \
      return mb200_random_##dist(h, n, MB200_ARG(a), MB200_ARG(b), MB200_ARG(c), y); });      \
  }                                                                                            \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_n_##dist(size_t n_samples, cpp11::doubles a, \
                                                               cpp11::doubles b, cpp11::doubles c, \
                                                               cpp11::sexp ptr) {             \
    return sample(ptr, #dist, true, n_samples, [&](mb200_rng* h, size_t n, double* y) {       \
      return mb200_random_##dist(h, n, MB200_ARG(a), MB200_ARG(b), MB200_ARG(c), y); });      \
  }
#define MB200_GLUE4(dist, a, b, c, d)                                                          \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_##dist(cpp11::doubles a, cpp11::doubles b, \
                                                             cpp11::doubles c, cpp11::doubles d, \
                                                             cpp11::sexp ptr) {               \
    return sample(ptr, #dist, false, 1, [&](mb200_rng* h, size_t n, double* y) {              \
      return mb200_random_##dist(h, n, MB200_ARG(a), MB200_ARG(b), MB200_ARG(c), MB200_ARG(d), y); }); \
  }                                                                                            \
  [[cpp11::register]] cpp11::doubles cpp_monty_random_n_##dist(size_t n_samples, cpp11::doubles a, \
                                                               cpp11::doubles b, cpp11::doubles c, \
                                                               cpp11::doubles d, cpp11::sexp ptr) { \
    return sample(ptr, #dist, true, n_samples, [&](mb200_rng* h, size_t n, double* y) {       \
      return mb200_random_##dist(h, n, MB200_ARG(a), MB200_ARG(b), MB200_ARG(c), MB200_ARG(d), y); }); \
  }

MB200_GLUE0(real)
MB200_GLUE1(exponential_rate, rate)
MB200_GLUE1(exponential_mean, mean)
MB200_GLUE1(poisson, lambda)
MB200_GLUE2(beta, a, b)
MB200_GLUE2(binomial, size, prob)
MB200_GLUE2(cauchy, location, scale)
MB200_GLUE2(gamma_scale, shape, scale)
MB200_GLUE2(gamma_rate, shape, rate)
MB200_GLUE2(negative_binomial_prob, size, prob)
MB200_GLUE2(negative_binomial_mu, size, mu)
MB200_GLUE2(normal_box_muller, mean, sd)
MB200_GLUE2(normal_polar, mean, sd)
MB200_GLUE2(normal_ziggurat, mean, sd)
MB200_GLUE2(uniform, min, max)
MB200_GLUE2(weibull, shape, scale)
MB200_GLUE2(log_normal, meanlog, sdlog)
MB200_GLUE2(zi_poisson, pi0, lambda)
MB200_GLUE3(beta_binomial_prob, size, prob, rho)
MB200_GLUE3(beta_binomial_ab, size, a, b)
MB200_GLUE3(hypergeometric, n1, n2, k)
MB200_GLUE3(zi_negative_binomial_prob, pi0, size, prob)
MB200_GLUE3(zi_negative_binomial_mu, pi0, size, mu)
MB200_GLUE4(truncated_normal, mean, sd, min, max)

// ---- multinomial (ref: src/random.cpp:884-934) ---------------------------------
namespace {

// `prob`: a vector (shared by all streams) or a matrix with one column per
// stream (or one column); ref: input_array, src/random.cpp:44-83
struct prob_matrix {
  const double* data;
  size_t len, cols;
  prob_matrix(cpp11::doubles r_prob, size_t n_streams) : data(REAL(r_prob)) {
    const SEXP dim = r_prob.attr("dim");
    if (dim == R_NilValue) {
      len = static_cast<size_t>(r_prob.size());
      cols = 1;
    } else {
      if (Rf_length(dim) != 2) cpp11::stop("Expected '%s' to be a matrix", "prob");
      len = static_cast<size_t>(INTEGER(dim)[0]);
      cols = static_cast<size_t>(INTEGER(dim)[1]);
      if (cols != 1 && cols != n_streams) {
        if (n_streams == 1) cpp11::stop("Expected '%s' to have 1 column, not %d", "prob", static_cast<int>(cols));
        cpp11::stop("Expected '%s' to have %d or 1 columns, not %d", "prob", static_cast<int>(n_streams),
                    static_cast<int>(cols));
      }
    }
  }
};

cpp11::doubles multinomial(bool is_n, size_t n_samples, cpp11::doubles r_size, cpp11::doubles r_prob,
                           cpp11::sexp ptr) {
  b200_rng* rng = read_ptr(ptr, "multinomial");
  const size_t n_streams = mb200_rng_n_streams(rng->h);
  const prob_matrix prob(r_prob, n_streams);
  const size_t ns = is_n ? n_samples : 1;
  cpp11::writable::doubles r_y(static_cast<R_xlen_t>(prob.len * ns * n_streams));
  if (ns > 0) {
    check(mb200_rng_multinomial(rng->h, ns, REAL(r_size), static_cast<size_t>(r_size.size()), prob.data, prob.len,
                                prob.cols, REAL(r_y)), rng);
  }
  const int len = static_cast<int>(prob.len), n = static_cast<int>(n_samples), s = static_cast<int>(n_streams);
  if (!is_n) {
    if (preserve_stream_dimension(n_streams, ptr)) r_y.attr("dim") = cpp11::writable::integers{len, s};
  } else if (preserve_stream_dimension(n_streams, ptr)) {
    r_y.attr("dim") = cpp11::writable::integers{len, n, s};
  } else {
    r_y.attr("dim") = cpp11::writable::integers{len, n};
  }
  return r_y;
}

}  // namespace

[[cpp11::register]]
cpp11::doubles cpp_monty_random_multinomial(cpp11::doubles r_size, cpp11::doubles r_prob, cpp11::sexp ptr) {
  return multinomial(false, 1, r_size, r_prob, ptr);
}

[[cpp11::register]]
cpp11::doubles cpp_monty_random_n_multinomial(size_t n_samples, cpp11::doubles r_size, cpp11::doubles r_prob,
                                              cpp11::sexp ptr) {
  return multinomial(true, n_samples, r_size, r_prob, ptr);
}
