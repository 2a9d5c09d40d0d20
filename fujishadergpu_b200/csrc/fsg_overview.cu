// 2 x 2 AVERAGE overview level of an encoded raster (the pyramid a COG carries; reference:
// get_cog_options, core/dask_processor.py:201-228 -- OVERVIEW_RESAMPLING=AVERAGE, 8 levels; io/cog_builder.py:295-310).
// GDAL's average skips NoData members (DN 0 for the integer encodings, NaN for float32); a cell without valid
// members is NoData.  Integers: floor(mean + 0.5) evaluated in f64; a mean that would collide with the NoData
// value is moved to the nearest valid DN (GDAL >= 3.5 does the same).  SURVEY.md 8f rank 3.
#include "fsg_common.cuh"

namespace fsg {

template <typename T>
struct OvTraits;
template <>
struct OvTraits<uint8_t> { static constexpr bool is_float = false; };
template <>
struct OvTraits<int16_t> { static constexpr bool is_float = false; };
template <>
struct OvTraits<float> { static constexpr bool is_float = true; };

template <typename T>
__global__ void __launch_bounds__(256) overview_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t H, int64_t W,
                                                       int64_t ld_in, int64_t oh, int64_t ow, int64_t ld_out, double nodata,
                                                       int has_nodata) {
  const int64_t x = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (x >= ow) return;
  for (int64_t y = blockIdx.y; y < oh; y += gridDim.y) {
    double sum = 0.0;
    int n = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int64_t sy = 2 * y + dy, sx = 2 * x + dx;
        if (sy < H && sx < W) {
          const T v = in[sy * ld_in + sx];
          bool ok;
          if (OvTraits<T>::is_float) ok = (float)v == (float)v;
          else ok = !has_nodata || (double)v != nodata;
          if (ok) { sum += (double)v; ++n; }
        }
      }
    }
    T r;
    if (n == 0) {
      r = OvTraits<T>::is_float ? (T)nanf("") : (T)nodata;
    } else if (OvTraits<T>::is_float) {
      r = (T)(sum / (double)n);
    } else {
      double m = floor(sum / (double)n + 0.5);
      if (has_nodata && m == nodata) m = (sum / (double)n >= nodata) ? nodata + 1.0 : nodata - 1.0;
      r = (T)m;
    }
    out[y * ld_out + x] = r;
  }
}

}  // namespace fsg

extern "C" {

/* kind: FSG_OUT_F32 / FSG_OUT_I16 / FSG_OUT_U8 (element type of in and out); out is ceil(H/2) x ceil(W/2). */
int fsg_overview_average(const void* in, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out, int kind,
                         double nodata, int has_nodata, void* stream) {
  using namespace fsg;
  if (!in || !out || H < 1 || W < 1 || ld_in < W) return fail(FSG_E_INVALID, "fsg_overview_average: bad argument");
  const int64_t oh = (H + 1) / 2, ow = (W + 1) / 2;
  if (ld_out < ow) return fail(FSG_E_INVALID, "fsg_overview_average: output stride too small");
  dim3 grid((unsigned)((ow + 255) / 256), (unsigned)(oh < 65535 ? oh : 65535));
  cudaStream_t s = (cudaStream_t)stream;
  if (kind == FSG_OUT_F32)
    overview_kernel<float><<<grid, 256, 0, s>>>((const float*)in, (float*)out, H, W, ld_in, oh, ow, ld_out, nodata, has_nodata);
  else if (kind == FSG_OUT_I16)
    overview_kernel<int16_t><<<grid, 256, 0, s>>>((const int16_t*)in, (int16_t*)out, H, W, ld_in, oh, ow, ld_out, nodata, has_nodata);
  else if (kind == FSG_OUT_U8)
    overview_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)in, (uint8_t*)out, H, W, ld_in, oh, ow, ld_out, nodata, has_nodata);
  else
    return fail(FSG_E_INVALID, "fsg_overview_average: unknown element kind %d", kind);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

}  // extern "C"
