// Minimal stand-in for cpp11 (declarations with the signatures r_glue/ relies on),
// for `g++ -fsyntax-only`.  TEST INFRASTRUCTURE; see tests/test_r_glue.py.
#pragma once
#include <initializer_list>
#include <string>
#include <vector>
#include "../Rinternals.h"
namespace cpp11 {
class sexp;
struct attribute_proxy {
  operator SEXP() const;
  template <typename T> attribute_proxy& operator=(const T&);
};
class sexp {
 public:
  sexp();
  sexp(SEXP);
  operator SEXP() const;
  attribute_proxy attr(const char*) const;
};
bool operator==(const sexp&, SEXP);
template <typename T> T as_cpp(SEXP);
template <typename... A> [[noreturn]] void stop(const char*, A...);
template <typename T> class r_vector {
 public:
  r_vector();
  r_vector(SEXP);
  operator SEXP() const;
  operator sexp() const;
  R_xlen_t size() const;
  attribute_proxy attr(const char*) const;
};
namespace writable {
template <typename T> class r_vector : public cpp11::r_vector<T> {
 public:
  explicit r_vector(R_xlen_t);
  r_vector(std::initializer_list<T>);
};
}  // namespace writable
}  // namespace cpp11
