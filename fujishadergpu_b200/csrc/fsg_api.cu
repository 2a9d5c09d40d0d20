// Library-wide plumbing: thread-local error text, launch counter, shared host helpers.
#include <stdarg.h>

#include "fsg_common.cuh"

namespace fsg {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

char* last_error_buf() { return g_err; }
void count_launch(int n) { g_launches += n; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int gauss_half_taps(double sigma, double* w, int max_radius) {
  // scipy.ndimage._filters._gaussian_kernel1d with truncate=4.0
  int r = (int)(4.0 * sigma + 0.5);
  if (r > max_radius) return -1;
  double s2 = sigma * sigma;
  double tot = 0.0;
  for (int x = -r; x <= r; ++x) tot += exp(-0.5 / s2 * (double)(x * x));
  for (int x = 0; x <= r; ++x) w[x] = exp(-0.5 / s2 * (double)(x * x)) / tot;
  return r;
}

}  // namespace fsg

extern "C" {

const char* fsg_last_error(void) { return fsg::last_error_buf(); }
int fsg_version(void) { return 100; }
int64_t fsg_launch_count(void) { return fsg::g_launches; }
void fsg_reset_launch_count(void) { fsg::g_launches = 0; }

}  // extern "C"
