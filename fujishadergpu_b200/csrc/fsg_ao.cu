// Stylised screen-space ambient occlusion -- compute_ambient_occlusion_block
// (algorithms/_impl_ambient_occlusion.py:33-118) as a gather kernel + the existing sigma=1 Gaussian + a
// gamma / NoData epilogue.  SURVEY.md 8f rank 4: same ray-sample structure as openness.
//
// Per pixel: 4 rings (0.25 / 0.5 / 0.75 / 1.0 x radius) x num_samples azimuths; sample offset
// (np.round(r cos a), np.round(r sin a)), the (0,0) offsets skipped; sample value = the raster
// EDGE-REPLICATED (cp.pad(mode='edge'), so out-of-raster samples are valid unless the edge pixel is NaN);
// occlusion = min(max(0, arctan((sample - centre)/dist)) / (pi/4), 1) * (1 - 0.3 ring factor), summed in the
// reference's loop order (ring outer, azimuth inner) over the valid samples; ao = clip(1 - mean*intensity).
// NoData pixels are filled with 1.0 before the smoothing (:105-107) -- they come out as 1.0 of the formula
// anyway (no valid sample) -- then gaussian_filter(sigma=1, 'nearest'), power(ao, 1/2.2), NaN restore.
// The host layer builds the sample table with NumPy exactly as the reference does (kernels.ao_table).
#include <stdlib.h>

#include "fsg_common.cuh"
#include "fsg_filters.cuh"

extern "C" size_t fsg_gaussian_nan_workspace_bytes(int64_t H, int64_t W, double sigma);
extern "C" int fsg_gaussian_nan(const float* in, float* out, int64_t H, int64_t W, int64_t ld_in, double sigma,
                                void* workspace, size_t workspace_bytes, void* stream);

namespace fsg {

constexpr int AO_MAX_SAMPLES = 4 * 64;

struct AoSample {
  short ox, oy;
  float dist;    // f32(max(hypot(ox*sx, oy*sy), 1e-9))
  float factor;  // f32(1 - 0.3 * ring factor)
  float rinv;    // f32(1 / dist): division by the constant distance = multiply + one FMA correction
};

// a / b for a constant divisor b with rb = f32(1/b): q = a*rb corrected by one FMA residual step is the
// correctly rounded quotient (Markstein), i.e. the value of the IEEE division, at a third of the instructions
__device__ __forceinline__ float div_const(float a, float b, float rb) {
  const float q = a * rb;
  const float r = fmaf(-q, b, a);
  return (r == r) ? fmaf(r, rb, q) : q;   // residual is NaN only for infinite / NaN quotients
}

struct AoTable {   // by value through the kernel parameter bank: no global symbol, re-entrant
  int n;
  AoSample s[AO_MAX_SAMPLES];
};

// 64 x 4 pixels per CTA, thread per pixel.  Interior CTAs (every sample inside the raster) skip the clamps.
__global__ void __launch_bounds__(256) ao_raw_kernel(const float* __restrict__ dem, float* __restrict__ ao, int64_t H,
                                                     int64_t W, int64_t ld, float intensity, int halo,
                                                     const __grid_constant__ AoTable tab) {
  const int64_t x0 = (int64_t)blockIdx.x * 64, y0 = (int64_t)blockIdx.y * 4;
  const int64_t x = x0 + (threadIdx.x & 63), y = y0 + (threadIdx.x >> 6);
  if (x >= W || y >= H) return;
  const bool interior = x0 >= halo && x0 + 64 + halo <= W && y0 >= halo && y0 + 4 + halo <= H;
  const float* pc = dem + y * ld + x;
  const float c = __ldg(pc);
  const float max_angle = (float)(3.14159265358979323846 / 4);
  float total = 0.f, count = 0.f;
  if (c == c) {
    if (interior) {
      // a sample at or below the centre occludes nothing: maximum(0, arctan(negative)) = 0, no arctan needed
      // (terrain slopes are coherent, so whole warps usually agree and skip it together)
      const float inv_ma = 1.0f / max_angle;
#pragma unroll 4
      for (int k = 0; k < tab.n; ++k) {
        const AoSample sm = tab.s[k];
        const float v = __ldg(pc + ((int64_t)sm.oy * ld + sm.ox));
        const float d = v - c;
        if (d > 0.f) {   // (approximate SFU forms, fsg_common.cuh: contract 1e-5 relative / 1e-6 absolute)
          float o = fast_atan_pos(d * sm.rinv) * inv_ma;
          o = fminf(o, 1.0f);
          total = total + o * sm.factor;
        }
        count = count + ((v == v) ? 1.f : 0.f);
      }
    } else {
      for (int k = 0; k < tab.n; ++k) {
        const AoSample sm = tab.s[k];
        const int64_t sy = clamp_index(y + sm.oy, H), sx = clamp_index(x + sm.ox, W);
        const float v = __ldg(dem + sy * ld + sx);
        if (v == v) {
          const float d = v - c;
          float o = d > 0.f ? fast_atan_pos(d * sm.rinv) * (1.0f / max_angle) : 0.f;
          o = fminf(o, 1.0f);
          total = total + o * sm.factor;
          count = count + 1.f;
        }
      }
    }
  }
  const float mean = total / fmaxf(count, 1.0f);
  float a = 1.0f - mean * intensity;
  a = fminf(fmaxf(a, 0.f), 1.f);
  ao[y * W + x] = (c == c) ? a : 1.0f;
}

// power(ao, 1/2.2), optional display stretch (tile/dask_bridge.py:173-187), NaN restore, encoding
__global__ void __launch_bounds__(256) ao_finish_kernel(const float* __restrict__ dem, const float* __restrict__ sm,
                                                        void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out,
                                                        int stretch, float lo, float scale, EncodeDev enc) {
  const int64_t x = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (x >= W) return;
  for (int64_t y = blockIdx.y; y < H; y += gridDim.y) {
    const float c = dem[y * ld_in + x];
    float r = fast_gamma22(sm[y * W + x]);
    if (stretch) r = fmaxf((r - lo) / scale, 0.f);
    if (c != c) r = nanf("");
    store_out(out, y * ld_out + x, r, enc);
  }
}

}  // namespace fsg

extern "C" {

size_t fsg_ambient_occlusion_workspace_bytes(int64_t H, int64_t W) {
  if (H < 1 || W < 1) return 0;
  size_t plane = ((size_t)H * (size_t)W * 4 + 255) / 256 * 256;
  return 2 * plane + fsg_gaussian_nan_workspace_bytes(H, W, 1.0);
}

/* compute_ambient_occlusion_block on the whole H x W block.  Sample table (n <= 256 entries, the (0,0)
 * offsets already dropped): ox/oy pixel offsets, dist = f32 physical distance, factor = f32(1 - 0.3 ring). */
int fsg_ambient_occlusion(const float* dem, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out,
                          int n_samples, const int32_t* ox_host, const int32_t* oy_host, const float* dist_host,
                          const float* factor_host, double intensity, double stretch_lo, double stretch_scale,
                          const fsg_encode* enc, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace fsg;
  if (!dem || !out || H < 1 || W < 1 || ld_in < W || ld_out < W) return fail(FSG_E_INVALID, "fsg_ambient_occlusion: bad argument");
  if (n_samples < 0 || n_samples > AO_MAX_SAMPLES || (n_samples > 0 && (!ox_host || !oy_host || !dist_host || !factor_host)))
    return fail(FSG_E_INVALID, "fsg_ambient_occlusion: 0..%d samples supported", AO_MAX_SAMPLES);
  if ((H + 3) / 4 > 65535) return fail(FSG_E_UNSUPPORTED, "fsg_ambient_occlusion: more than 262140 rows per block");
  const size_t need = fsg_ambient_occlusion_workspace_bytes(H, W);
  if (!workspace || workspace_bytes < need)
    return fail(FSG_E_WORKSPACE, "fsg_ambient_occlusion: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  AoTable tab{};
  tab.n = n_samples;
  int halo = 0;
  for (int k = 0; k < n_samples; ++k) {
    if (abs(ox_host[k]) > 32000 || abs(oy_host[k]) > 32000) return fail(FSG_E_INVALID, "fsg_ambient_occlusion: offset out of range");
    tab.s[k].ox = (short)ox_host[k]; tab.s[k].oy = (short)oy_host[k];
    tab.s[k].dist = dist_host[k]; tab.s[k].factor = factor_host[k]; tab.s[k].rinv = 1.0f / dist_host[k];
    const int a = abs(ox_host[k]) > abs(oy_host[k]) ? abs(ox_host[k]) : abs(oy_host[k]);
    if (a > halo) halo = a;
  }
  const size_t plane = ((size_t)H * (size_t)W * 4 + 255) / 256 * 256;
  float* raw = (float*)workspace;
  float* smooth = (float*)((unsigned char*)workspace + plane);
  void* gws = (unsigned char*)workspace + 2 * plane;
  dim3 grid((unsigned)((W + 63) / 64), (unsigned)((H + 3) / 4));
  ao_raw_kernel<<<grid, 256, 0, s>>>(dem, raw, H, W, ld_in, (float)intensity, halo, tab);
  FSG_LAUNCH_OK();
  int rc = fsg_gaussian_nan(raw, smooth, H, W, W, 1.0, gws, workspace_bytes - 2 * plane, stream);
  if (rc) return rc;
  const int stretch = (!is_none(stretch_scale) && !is_none(stretch_lo) && stretch_scale > 1e-12) ? 1 : 0;
  dim3 g2((unsigned)((W + 255) / 256), (unsigned)(H < 65535 ? H : 65535));
  ao_finish_kernel<<<g2, 256, 0, s>>>(dem, smooth, out, H, W, ld_in, ld_out, stretch, (float)stretch_lo,
                                      (float)stretch_scale, make_encode(enc));
  FSG_LAUNCH_OK();
  return FSG_OK;
}

}  // extern "C"
