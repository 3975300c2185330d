// r_glue/test_rng.cpp -- replaces src/test_rng.cpp of mrc-ide/monty (ref:
// src/test_rng.cpp:18-38, registered in src/cpp11.cpp:631): 30 values of
// stream 0 with a jump before value 10 and a long jump before value 20, as
// decimal strings (the golden files of tests/testthat/xoshiro-ref).  Only
// stream 0 moves, as in the reference: it is lifted into a one-stream handle,
// run there, and its final state written back.
#include <cpp11/external_pointer.hpp>
#include <cpp11/sexp.hpp>
#include <cpp11/protect.hpp>

#include <cstdint>
#include <string>
#include <vector>

#include <monty_b200.h>

namespace {
struct b200_rng {           // the same handle type as r_glue/random.cpp
  mb200_rng* h = nullptr;
  ~b200_rng() { mb200_rng_free(h); }
};
void ok(int rc) {
  if (rc != MB200_OK) cpp11::stop("%s", mb200_last_error());
}
}  // namespace

[[cpp11::register]]
std::vector<std::string> test_xoshiro_run(cpp11::sexp ptr) {
  b200_rng* rng = cpp11::as_cpp<cpp11::external_pointer<b200_rng>>(ptr).get();
  const size_t n_streams = mb200_rng_n_streams(rng->h);
  std::vector<uint8_t> all(32 * n_streams);
  ok(mb200_rng_state(rng->h, all.data(), all.size()));
  b200_rng one;
  ok(mb200_rng_create(0, all.data(), 32, 1, 0, 0, 0, mb200_rng_device(rng->h), &one.h));

  constexpr int n = 10;
  std::vector<std::string> ret;
  uint64_t x = 0;
  for (int i = 0; i < 3 * n; ++i) {
    if (i == n - 1) ok(mb200_rng_jump(one.h, 1));
    else if (i == 2 * n - 1) ok(mb200_rng_long_jump(one.h, 1));
    ok(mb200_rng_raw(one.h, MB200_GEN_XOSHIRO256PLUS, 1, &x));
    ret.push_back(std::to_string(x));
  }
  ok(mb200_rng_state(one.h, all.data(), 32));
  ok(mb200_rng_set_state(rng->h, all.data(), all.size()));
  return ret;
}
