// r_glue/density.cpp -- replaces src/density.cpp of mrc-ide/monty (ref:
// src/density.cpp:6-254, registered in src/cpp11.cpp:609-629): the same 21
// routines, each one call of mb200_density.  Every argument is indexed by i
// (no recycling), integer arguments stay integers, the result is a double
// vector of x's length -- as in the reference.
#include <cpp11/doubles.hpp>
#include <cpp11/integers.hpp>
#include <cpp11/protect.hpp>

#include <monty_b200.h>

namespace {

struct arg {
  const void* data;
  int is_i32;
  arg(const cpp11::doubles& v) : data(REAL(v)), is_i32(0) {}
  arg(const cpp11::integers& v) : data(INTEGER(v)), is_i32(1) {}
};

template <typename X, size_t K>
SEXP density(int dens, bool is_float, const X& x, const arg (&p)[K], bool log) {
  const size_t n = static_cast<size_t>(x.size());
  cpp11::writable::doubles ret(static_cast<R_xlen_t>(n));
  const arg xa(x);
  const void* ptrs[MB200_MAX_PARAMS] = {nullptr, nullptr, nullptr, nullptr};
  int is_i32[MB200_MAX_PARAMS] = {0, 0, 0, 0};
  for (size_t k = 0; k < K; ++k) { ptrs[k] = p[k].data; is_i32[k] = p[k].is_i32; }
  const int rc = mb200_density(dens, is_float ? MB200_REAL_F32 : MB200_REAL_F64, n, xa.data, xa.is_i32, ptrs, is_i32,
                               log ? 1 : 0, REAL(ret), nullptr, /*device=*/0);
  if (rc != MB200_OK) cpp11::stop("%s", mb200_last_error());
  return ret;
}

}  // namespace

using cpp11::doubles;
using cpp11::integers;

[[cpp11::register]] SEXP density_binomial(integers x, integers size, doubles prob, bool log) {
  const arg p[] = {size, prob};
  return density(MB200_DENS_BINOMIAL, false, x, p, log);
}
[[cpp11::register]] SEXP density_normal(doubles x, doubles mu, doubles sd, bool log) {
  const arg p[] = {mu, sd};
  return density(MB200_DENS_NORMAL, false, x, p, log);
}
// is_float: evaluate in single precision (ref: src/density.cpp:28-45)
[[cpp11::register]] SEXP density_negative_binomial_mu(integers x, doubles size, doubles mu, bool log, bool is_float) {
  const arg p[] = {size, mu};
  return density(MB200_DENS_NEGATIVE_BINOMIAL_MU, is_float, x, p, log);
}
[[cpp11::register]] SEXP density_negative_binomial_prob(integers x, doubles size, doubles prob, bool log) {
  const arg p[] = {size, prob};
  return density(MB200_DENS_NEGATIVE_BINOMIAL_PROB, false, x, p, log);
}
[[cpp11::register]] SEXP density_beta_binomial_prob(integers x, integers size, doubles prob, doubles rho, bool log) {
  const arg p[] = {size, prob, rho};
  return density(MB200_DENS_BETA_BINOMIAL_PROB, false, x, p, log);
}
[[cpp11::register]] SEXP density_beta_binomial_ab(integers x, integers size, doubles a, doubles b, bool log) {
  const arg p[] = {size, a, b};
  return density(MB200_DENS_BETA_BINOMIAL_AB, false, x, p, log);
}
[[cpp11::register]] SEXP density_poisson(integers x, doubles lambda, bool log) {
  const arg p[] = {lambda};
  return density(MB200_DENS_POISSON, false, x, p, log);
}
[[cpp11::register]] SEXP density_uniform(doubles x, doubles min, doubles max, bool log) {
  const arg p[] = {min, max};
  return density(MB200_DENS_UNIFORM, false, x, p, log);
}
[[cpp11::register]] SEXP density_exponential_rate(doubles x, doubles rate, bool log) {
  const arg p[] = {rate};
  return density(MB200_DENS_EXPONENTIAL_RATE, false, x, p, log);
}
[[cpp11::register]] SEXP density_exponential_mean(doubles x, doubles mean, bool log) {
  const arg p[] = {mean};
  return density(MB200_DENS_EXPONENTIAL_MEAN, false, x, p, log);
}
[[cpp11::register]] SEXP density_gamma_rate(doubles x, doubles shape, doubles rate, bool log) {
  const arg p[] = {shape, rate};
  return density(MB200_DENS_GAMMA_RATE, false, x, p, log);
}
[[cpp11::register]] SEXP density_gamma_scale(doubles x, doubles shape, doubles scale, bool log) {
  const arg p[] = {shape, scale};
  return density(MB200_DENS_GAMMA_SCALE, false, x, p, log);
}
[[cpp11::register]] SEXP density_beta(doubles x, doubles a, doubles b, bool log) {
  const arg p[] = {a, b};
  return density(MB200_DENS_BETA, false, x, p, log);
}
[[cpp11::register]] SEXP density_hypergeometric(integers x, integers n1, integers n2, integers k, bool log) {
  const arg p[] = {n1, n2, k};
  return density(MB200_DENS_HYPERGEOMETRIC, false, x, p, log);
}
[[cpp11::register]] SEXP density_log_normal(doubles x, doubles mulog, doubles sdlog, bool log) {
  const arg p[] = {mulog, sdlog};
  return density(MB200_DENS_LOG_NORMAL, false, x, p, log);
}
[[cpp11::register]] SEXP density_truncated_normal(doubles x, doubles mu, doubles sd, doubles min, doubles max, bool log) {
  const arg p[] = {mu, sd, min, max};
  return density(MB200_DENS_TRUNCATED_NORMAL, false, x, p, log);
}
[[cpp11::register]] SEXP density_cauchy(doubles x, doubles location, doubles scale, bool log) {
  const arg p[] = {location, scale};
  return density(MB200_DENS_CAUCHY, false, x, p, log);
}
[[cpp11::register]] SEXP density_weibull(doubles x, doubles shape, doubles scale, bool log) {
  const arg p[] = {shape, scale};
  return density(MB200_DENS_WEIBULL, false, x, p, log);
}
[[cpp11::register]] SEXP density_zi_poisson(integers x, doubles pi0, doubles lambda, bool log) {
  const arg p[] = {pi0, lambda};
  return density(MB200_DENS_ZI_POISSON, false, x, p, log);
}
[[cpp11::register]] SEXP density_zi_negative_binomial_mu(integers x, doubles pi0, doubles size, doubles mu, bool log) {
  const arg p[] = {pi0, size, mu};
  return density(MB200_DENS_ZI_NEGATIVE_BINOMIAL_MU, false, x, p, log);
}
[[cpp11::register]] SEXP density_zi_negative_binomial_prob(integers x, doubles pi0, doubles size, doubles prob, bool log) {
  const arg p[] = {pi0, size, prob};
  return density(MB200_DENS_ZI_NEGATIVE_BINOMIAL_PROB, false, x, p, log);
}
