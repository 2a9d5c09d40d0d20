// Elementwise helpers: integer encoding, normalise, display stretch, synthetic DEM generator.
#include "fsg_common.cuh"

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
