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

// ---- optional per-thread kernel timing (bench.py's roofline measurement) -------------------
// When enabled, instrumented call sites bracket their dominant kernel with CUDA events recorded on
// the launch stream; fsg_profile_read() synchronises the events and returns the durations.
constexpr int PROF_MAX = 4096;   // (a bench step records ~20 launches; 20 steps overflowed the 256 of round 1)
struct ProfState {
  int enabled = 0;
  int n = 0;
  cudaEvent_t ev[2 * PROF_MAX] = {};
  int tag[PROF_MAX] = {};
};
static thread_local ProfState g_prof;

int prof_begin(int tag, cudaStream_t s) {
  ProfState& p = g_prof;
  if (!p.enabled || p.n >= PROF_MAX) return -1;
  int i = p.n;
  for (int k = 0; k < 2; ++k)
    if (!p.ev[2 * i + k] && cudaEventCreate(&p.ev[2 * i + k]) != cudaSuccess) return -1;
  p.tag[i] = tag;
  cudaEventRecord(p.ev[2 * i], s);
  return i;
}
void prof_end(int slot, cudaStream_t s) {
  if (slot < 0) return;
  cudaEventRecord(g_prof.ev[2 * slot + 1], s);
  g_prof.n = slot + 1;
}

}  // namespace fsg

extern "C" {

void fsg_profile_enable(int on) { fsg::g_prof.enabled = on; fsg::g_prof.n = 0; }

// Writes up to `max` (tag, milliseconds) pairs; returns the number of records; resets the list.
int fsg_profile_read(int* tags, float* ms, int max) {
  fsg::ProfState& p = fsg::g_prof;
  int n = p.n < max ? p.n : max;
  for (int i = 0; i < n; ++i) {
    cudaEventSynchronize(p.ev[2 * i + 1]);
    float t = 0.f;
    cudaEventElapsedTime(&t, p.ev[2 * i], p.ev[2 * i + 1]);
    tags[i] = p.tag[i];
    ms[i] = t;
  }
  p.n = 0;
  return n;
}


const char* fsg_last_error(void) { return fsg::last_error_buf(); }
int fsg_version(void) { return 100; }
int64_t fsg_launch_count(void) { return fsg::g_launches; }
void fsg_reset_launch_count(void) { fsg::g_launches = 0; }

}  // extern "C"
