#pragma once
#include <cstdint>
#include "sexp.hpp"
namespace cpp11 {
typedef r_vector<int> integers;
namespace writable { typedef r_vector<int> integers; }
}
