// Elementwise helpers: integer encoding, normalise, display stretch, synthetic DEM generator.
#include "fsg_filters.cuh"

namespace fsg {

__global__ void encode_kernel(const float* __restrict__ in, void* out, int64_t n, EncodeDev e) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) store_out(out, i, in[i], e);
}

// out = in / f32(scale); NaN stays NaN (topousm_fast_norm_func, algorithms/_normalization.py:35-41)
__global__ void scale_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float scale, int zero) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    float v = in[i];
    out[i] = (v != v) ? v : (zero ? 0.f : v / scale);
  }
}

// max((x - f32(lo)) / f32(scale), 0)  (algorithms/tile/dask_bridge.py:173-187); NaN propagates
__global__ void stretch_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float lo, float scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    float v = (in[i] - lo) / scale;
    out[i] = (v != v) ? v : fmaxf(v, 0.f);  // np.maximum propagates NaN
  }
}

__device__ __forceinline__ uint32_t hash32(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return (uint32_t)k;
}

struct SynthParams {
  float amp[8], kr[8], kc[8], ph[8];
};

// z = 400 + sum_k A_k sin(2 pi (r u_k + c v_k)/lambda_k + phi_k) + 0.25*noise, lambda_k = 4096/2^k,
// A_k = 250*2^(-0.9k)  (SURVEY.md 8d).  Noise: sum of 4 uniform hashes (approx. normal, sigma 0.25 m).
__global__ void synth_kernel(float* out, int64_t H, int64_t W, int64_t row0, int64_t rows, int64_t ld,
                             uint64_t seed, int nodata, SynthParams sp) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t ry = (int64_t)blockIdx.y * 8;
  if (x >= W) return;
  for (int j = 0; j < 8; ++j) {
    int64_t r = ry + j;
    if (r >= rows) return;
    int64_t y = row0 + r;
    float z = 400.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // reduce the phase in double so large coordinates keep full precision
      double t = (double)y * (double)sp.kr[k] + (double)x * (double)sp.kc[k];
      t -= floor(t);
      z += sp.amp[k] * __sinf(6.283185307f * (float)t + sp.ph[k]);
    }
    uint64_t key = (uint64_t)y * (uint64_t)W + (uint64_t)x + seed * 0x9E3779B97F4A7C15ULL;
    uint32_t h1 = hash32(key), h2 = hash32(key ^ 0xD1B54A32D192ED03ULL);
    float u = (float)(h1 & 0xffff) + (float)(h1 >> 16) + (float)(h2 & 0xffff) + (float)(h2 >> 16);
    z += 0.25f * ((u * (1.f / 65536.f) - 2.f) * 1.7320508f);  // var of sum of 4 U(0,1) = 1/3
    if (nodata) {
      float fw = (float)W, fh = (float)H;
      bool hole = (float)x < 0.02f * fw * (1.f + __sinf((float)y / 977.f));
      const float cy[3] = {0.3f, 0.7f, 0.55f}, cx[3] = {0.6f, 0.35f, 0.8f};
      const float ay[3] = {0.02f, 0.06f, 0.01f}, ax[3] = {0.05f, 0.03f, 0.012f};
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        float dy = ((float)y - cy[e] * fh) / fmaxf(2.f, ay[e] * fh);
        float dx = ((float)x - cx[e] * fw) / fmaxf(2.f, ax[e] * fw);
        hole |= (dy * dy + dx * dx) <= 1.f;
      }
      if (hole) z = nanf("");
    }
    out[r * ld + x] = z;
  }
}

// Gaussian taps computed on the HOST (glibc exp, like scipy's kernel) and handed over by value
constexpr int TAPS_MAX = 448;
struct TapsParam {
  int n;
  double w[TAPS_MAX];
};
__global__ void store_taps_kernel(const __grid_constant__ TapsParam tp, double* dst) {
  for (int i = threadIdx.x; i < tp.n; i += blockDim.x) dst[i] = tp.w[i];
}

// _combine_direct (algorithms/tile/dask_bridge.py:28-69), one response at a time.
// mode 0: acc = a*w            (first weighted response)
//      1: acc = acc + a*w      (further responses; also the equal-weight mean with w = 1/n)
//      2: acc = 0 + a*w        (first response of the equal-weight mean / sum: the reference starts from zeros)
//      3: acc = np.maximum(acc, a)   4: acc = np.minimum(acc, a)   (NaN propagates like NumPy)
//      5: acc = a
__global__ void combine_kernel(const float* __restrict__ a, float* __restrict__ acc, int64_t n, float w, int mode) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    float v = a[i];
    float r;
    if (mode == 0) r = v * w;
    else if (mode == 1) r = acc[i] + v * w;
    else if (mode == 2) r = 0.f + v * w;
    else if (mode == 5) r = v;
    else {
      float c = acc[i];
      if (v != v || c != c) r = nanf("");
      else r = mode == 3 ? fmaxf(c, v) : fminf(c, v);
    }
    acc[i] = r;
  }
}

static int grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace fsg

extern "C" {

int fsg_encode_f32(const float* in, void* out, int64_t n, const fsg_encode* enc, void* stream) {
  using namespace fsg;
  if (!in || !out || !enc) return fail(FSG_E_INVALID, "fsg_encode_f32: NULL argument");
  if (enc->kind != FSG_OUT_I16 && enc->kind != FSG_OUT_U8 && enc->kind != FSG_OUT_F32)
    return fail(FSG_E_INVALID, "fsg_encode_f32: unknown output kind %d", enc->kind);
  if (n <= 0) return FSG_OK;
  encode_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, make_encode(enc));
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_scale_f32(const float* in, float* out, int64_t n, double scale, void* stream) {
  using namespace fsg;
  if (!in || !out) return fail(FSG_E_INVALID, "fsg_scale_f32: NULL argument");
  if (n <= 0) return FSG_OK;
  int zero = !(scale > 0.0);
  scale_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, (float)scale, zero);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_stretch_f32(const float* in, float* out, int64_t n, double lo, double scale, void* stream) {
  using namespace fsg;
  if (!in || !out) return fail(FSG_E_INVALID, "fsg_stretch_f32: NULL argument");
  if (n <= 0) return FSG_OK;
  stretch_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, (float)lo, (float)scale);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

size_t fsg_gaussian_nan_workspace_bytes(int64_t H, int64_t W, double sigma) {
  if (H < 1 || W < 1 || !(sigma > 0.0)) return 0;
  int r = (int)(4.0 * sigma + 0.5);
  size_t plane = ((size_t)H * (size_t)W * 4 + 255) / 256 * 256;
  return 2 * plane + ((size_t)(r + 1) * 8 + 255) / 256 * 256;
}

/* handle_nan_with_gaussian(block, sigma, mode='nearest')[0]  (algorithms/_nan_utils.py:18-31) */
int fsg_gaussian_nan(const float* in, float* out, int64_t H, int64_t W, int64_t ld_in, double sigma, void* workspace,
                     size_t workspace_bytes, void* stream) {
  using namespace fsg;
  if (!in || !out || H < 1 || W < 1 || ld_in < W || !(sigma > 0.0))
    return fail(FSG_E_INVALID, "fsg_gaussian_nan: bad argument");
  size_t need = fsg_gaussian_nan_workspace_bytes(H, W, sigma);
  if (!workspace || workspace_bytes < need)
    return fail(FSG_E_WORKSPACE, "fsg_gaussian_nan: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  size_t plane = ((size_t)H * (size_t)W * 4 + 255) / 256 * 256;
  float* tv = (float*)workspace;
  float* tw = (float*)((unsigned char*)workspace + plane);
  double* taps = (double*)((unsigned char*)workspace + 2 * plane);
  int radius = (int)(4.0 * sigma + 0.5);
  int rc;
  if (radius + 1 <= TAPS_MAX) {
    TapsParam tp;
    tp.n = radius + 1;
    if (gauss_half_taps(sigma, tp.w, TAPS_MAX - 1) != radius) return fail(FSG_E_INVALID, "fsg_gaussian_nan: taps");
    store_taps_kernel<<<1, 256, 0, s>>>(tp, taps);
    FSG_LAUNCH_OK();
  } else if ((rc = launch_gauss_taps(sigma, radius, taps, s))) {
    return rc;
  }
  Grid g{in, H, W, ld_in};
  if ((rc = launch_gauss_axis0(g, taps, radius, tv, tw, nullptr, 0, H, s))) return rc;
  return launch_gauss_axis1(tv, tw, H, W, taps, radius, COMBINE_MEAN, out, nullptr, nullptr, s);
}

int fsg_combine_f32(const float* a, float* acc, int64_t n, double w, int mode, void* stream) {
  using namespace fsg;
  if (!a || !acc) return fail(FSG_E_INVALID, "fsg_combine_f32: NULL argument");
  if (mode < 0 || mode > 5) return fail(FSG_E_INVALID, "fsg_combine_f32: unknown mode %d", mode);
  if (n <= 0) return FSG_OK;
  combine_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(a, acc, n, (float)w, mode);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_synth_dem(float* out, int64_t H, int64_t W, int64_t row0, int64_t rows, int64_t ld, uint64_t seed,
                  int nodata, void* stream) {
  using namespace fsg;
  if (!out || H <= 0 || W <= 0 || rows < 0 || ld < W) return fail(FSG_E_INVALID, "fsg_synth_dem: bad argument");
  if (rows == 0) return FSG_OK;
  SynthParams sp;
  uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 12345;
  for (int k = 0; k < 8; ++k) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    double th = (double)(s >> 11) / 9007199254740992.0 * 6.283185307179586;
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    double ph = (double)(s >> 11) / 9007199254740992.0 * 6.283185307179586;
    double lam = 4096.0 / (double)(1 << k);
    sp.amp[k] = (float)(250.0 * pow(2.0, -0.9 * k));
    sp.kr[k] = (float)(sin(th) / lam);
    sp.kc[k] = (float)(cos(th) / lam);
    sp.ph[k] = (float)ph;
  }
  dim3 block(256), grid((unsigned)((W + 255) / 256), (unsigned)((rows + 7) / 8));
  synth_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(out, H, W, row0, rows, ld, seed, nodata, sp);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

}  // extern "C"

int fsg_copy_rect_f32(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int64_t cols,
                      void* stream) {
  using namespace fsg;
  if (rows < 0 || cols < 0 || ld_dst < cols || ld_src < cols) return fail(FSG_E_INVALID, "fsg_copy_rect_f32: bad shape");
  if (rows == 0 || cols == 0) return FSG_OK;
  if (!dst || !src) return fail(FSG_E_INVALID, "fsg_copy_rect_f32: NULL buffer");
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)ld_dst * 4, src, (size_t)ld_src * 4, (size_t)cols * 4, (size_t)rows,
                                    cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(FSG_E_CUDA, "fsg_copy_rect_f32: %s", cudaGetErrorString(e));
  return FSG_OK;
}

