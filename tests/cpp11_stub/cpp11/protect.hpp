#pragma once
#include "sexp.hpp"
