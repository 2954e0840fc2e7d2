// batched.cu -- batch of independent KKT systems sharing one pattern (config 5).
#include <cstring>

#include "../../include/cannoles_b200.h"
#include "b2_cuda.h"

namespace {
int nyi(const char* f) {
  snprintf(b2::g_last_error, sizeof(b2::g_last_error), "%s: batched engine not built yet", f);
  return -1;
}
}  // namespace

extern "C" {
int b2b_analyze(int64_t, int64_t, const int64_t*, const int64_t*, int64_t, int64_t, int64_t, int64_t,
                int, const int64_t*, int, b2b_handle**) { return nyi("b2b_analyze"); }
int b2b_factorize(b2b_handle*, const double*, const uint8_t*, double, int64_t*, int64_t*, int64_t*,
                  int32_t*) { return nyi("b2b_factorize"); }
int b2b_refactorize_shift(b2b_handle*, const double*, const double*, const uint8_t*, double, int64_t*,
                          int64_t*, int64_t*, int32_t*) { return nyi("b2b_refactorize_shift"); }
int b2b_solve(b2b_handle*, const double*, double*, const uint8_t*, int) { return nyi("b2b_solve"); }
int b2b_factorize_dev(b2b_handle*, const double*, const uint8_t*, double, int64_t*) { return nyi("b2b_factorize_dev"); }
int b2b_solve_dev(b2b_handle*, const double*, double*, const uint8_t*, int) { return nyi("b2b_solve_dev"); }
int b2b_stats(const b2b_handle*, b2_stats_t*) { return nyi("b2b_stats"); }
int b2b_free(b2b_handle*) { return 0; }
}
