#pragma once
#include "sexp.hpp"
namespace cpp11 {
template <typename T> class external_pointer {
 public:
  external_pointer(T*);
  external_pointer(SEXP);
  operator SEXP() const;
  T* get() const;
};
}
