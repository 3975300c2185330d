#pragma once
#include <cstdint>
#include "sexp.hpp"
namespace cpp11 {
typedef r_vector<uint8_t> raws;
namespace writable { typedef r_vector<uint8_t> raws; }
}
