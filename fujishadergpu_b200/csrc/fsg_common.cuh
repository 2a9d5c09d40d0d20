// Shared device/host helpers for libfsg_b200 (sm_100a).
// Compiled with -fmad=false: every f32 product/sum below is an individually rounded IEEE op,
// like the NumPy/CuPy elementwise expressions it mirrors.  Where a fused multiply-add is wanted
// (f64 helpers) it is written explicitly with fma().
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fsg_b200.h"

namespace fsg {

// ---- host-side error plumbing ------------------------------------------------------------
char* last_error_buf();
void count_launch(int n = 1);
int fail(int code, const char* fmt, ...);
int prof_begin(int tag, cudaStream_t s);   // returns a slot (or -1 when profiling is off)
void prof_end(int slot, cudaStream_t s);
enum { PROF_TOPOUSM_FUSED = 1, PROF_TOPOUSM_PYRAMID = 2, PROF_TOPOUSM_COARSE = 3, PROF_GRADIENT = 4, PROF_OPENNESS = 5,
       PROF_TOPOUSM_FUSED_ROI = 6 /* fused pass of a region of interest (statistics windows) */ };

#define FSG_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return fsg::fail(FSG_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define FSG_LAUNCH_OK()                                                                \
  do {                                                                                 \
    fsg::count_launch();                                                               \
    cudaError_t _e = cudaPeekAtLastError();                                            \
    if (_e != cudaSuccess)                                                             \
      return fsg::fail(FSG_E_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

inline bool is_none(double v) { return v != v; }

// Step sizes as the reference resolves them (algorithms/_nan_utils.py:64-71,
// _impl_curvature.py:23-28): `signed_steps` keeps the geotransform sign.
inline void resolve_steps(double pixel_size, double psx, double psy, bool signed_steps, double* sy, double* sx) {
  double y = is_none(psy) ? pixel_size : psy;
  double x = is_none(psx) ? pixel_size : psx;
  if (!signed_steps) { y = fabs(y); x = fabs(x); }
  if (fabs(y) < 1e-9) y = (pixel_size != 0.0) ? pixel_size : 1.0;
  if (fabs(x) < 1e-9) x = (pixel_size != 0.0) ? pixel_size : 1.0;
  *sy = y; *sx = x;
}

// Gaussian taps exactly as scipy.ndimage builds them (radius int(4*sigma+0.5), normalised f64).
// Returns radius; w[0..radius] = centre..outermost (symmetric).
int gauss_half_taps(double sigma, double* w, int max_radius);

// ---- device helpers ----------------------------------------------------------------------
struct EncodeDev {
  int kind;
  float a, b, lo, hi;
};
inline EncodeDev make_encode(const fsg_encode* e) {
  EncodeDev d;
  d.kind = e ? e->kind : FSG_OUT_F32;
  d.a = e ? (float)e->a_coef : 1.f;
  d.b = e ? (float)e->b_coef : 0.f;
  d.lo = e ? (float)e->dn_min : 0.f;
  d.hi = e ? (float)e->dn_max : 0.f;
  return d;
}
inline size_t out_elem_size(int kind) { return kind == FSG_OUT_U8 ? 1 : (kind == FSG_OUT_I16 ? 2 : 4); }

#ifdef __CUDACC__
// DN = clip(rint(a*v + b), lo, hi); non-finite -> 0  (core/dask_processor.py:1004-1008)
__device__ __forceinline__ float encode_dn(float v, const EncodeDev& e) {
  float dn = rintf(e.a * v + e.b);
  dn = fminf(fmaxf(dn, e.lo), e.hi);
  return isfinite(v) ? dn : 0.f;
}
__device__ __forceinline__ void store_out(void* out, int64_t idx, float v, const EncodeDev& e) {
  if (e.kind == FSG_OUT_F32) {
    ((float*)out)[idx] = v;
  } else if (e.kind == FSG_OUT_U8) {
    ((uint8_t*)out)[idx] = (uint8_t)(int)encode_dn(v, e);
  } else {
    ((int16_t*)out)[idx] = (int16_t)(int)encode_dn(v, e);
  }
}

// scipy 'reflect' (edge-inclusive mirror, d c b a | a b c d | d c b a), any distance.
__device__ __forceinline__ int64_t reflect_index(int64_t i, int64_t n) {
  if (n == 1) return 0;
  int64_t p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}
// the same for an index at most one period outside [0, n) (the common case: window reach < n); falls back otherwise
__device__ __forceinline__ int64_t reflect_near(int64_t i, int64_t n) {
  int64_t j = i < 0 ? -i - 1 : i;
  j = j >= n ? 2 * n - 1 - j : j;
  return (j >= 0 && j < n) ? j : reflect_index(i, n);
}
__device__ __forceinline__ int64_t clamp_index(int64_t i, int64_t n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// ---- approximate SFU forms ------------------------------------------------------------------------
// For the outputs whose contract is 1e-5 relative / 1e-6 absolute against the reference (hillshade, slope,
// curvature, openness, ambient occlusion) -- NOT for topousm_fast, whose window means must be exact.
// rsqrt / sqrt / rcp / ex2 / lg2 .approx are within 2 ulp (<= 2.4e-7 relative); the arctangent polynomial within
// 1.7e-7 relative.  The measured distance to the oracle is asserted in tests/test_gpu_parity.py.
__device__ __forceinline__ float sfu_rsqrt(float x) { float r; asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_sqrt(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_rcp(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_ex2(float x) { float r; asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_lg2(float x) { float r; asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// arctan(x) for x >= 0 (NaN -> NaN, +inf -> pi/2): atan(t) = t * P(t^2) on [0, 1] (degree-8 least-squares fit on
// Chebyshev nodes), atan(x) = pi/2 - atan(1/x) above 1
__device__ __forceinline__ float fast_atan_pos(float x) {
  const bool big = x > 1.f;
  const float t = big ? sfu_rcp(x) : x;
  const float u = t * t;
  float a = 0.0028340641874819994f;
  a = fmaf(a, u, -0.016005029901862144f);
  a = fmaf(a, u, 0.042587608098983765f);
  a = fmaf(a, u, -0.07495445758104324f);
  a = fmaf(a, u, 0.10636754333972931f);
  a = fmaf(a, u, -0.14202570915222168f);
  a = fmaf(a, u, 0.19992484152317047f);
  a = fmaf(a, u, -0.3333306610584259f);
  a = fmaf(a * u, t, t);   // t * (1 + u * q(u))
  return big ? 1.57079637050628662f - a : a;
}
__device__ __forceinline__ float fast_atan(float x) { return copysignf(fast_atan_pos(fabsf(x)), x); }
// x^(1/2.2) for x in [0, 1] (0 -> 0, NaN -> NaN)
__device__ __forceinline__ float fast_gamma22(float x) { return sfu_ex2((float)(1 / 2.2) * sfu_lg2(x)); }

// Correctly rounded f64 quotient s/n for a small positive integer n given inv = 1/n (rounded):
// one Newton correction of the rounded product (Markstein).
__device__ __forceinline__ double div_by_count(double s, double n, double inv) {
  double q = s * inv;
  double r = fma(-q, n, s);
  return fma(r, inv, q);
}
#endif

}  // namespace fsg
