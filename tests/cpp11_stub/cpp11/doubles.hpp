#pragma once
#include <cstdint>
#include "sexp.hpp"
namespace cpp11 {
typedef r_vector<double> doubles;
namespace writable { typedef r_vector<double> doubles; }
}
