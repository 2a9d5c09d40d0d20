// Generic NaN-aware separable filters on f32 grids (used on the decimated pyramid levels, for the
// sigma=1 Gaussian of local mode, for the coarse enclosed-void fill, and as the general fallback).
//
// Arithmetic contract (== scipy.ndimage on f32 input, which cupyx.scipy.ndimage mirrors):
//   * each axis pass accumulates in f64 and rounds ONCE to f32; axis 0 first;
//   * box ('reflect'): exact window sum / size, correctly rounded;
//   * gaussian ('nearest'): centre*w0 + sum_{j=R..1} (x[-j]+x[+j])*w[j]  (scipy's symmetric order);
//   * NaN-aware form of handle_nan_with_uniform/_gaussian (algorithms/_nan_utils.py:18-47):
//       mean = U(filled)/U(valid) where U(valid) > 0 else 0.
//     Evaluated per pixel; on a NaN-free grid U(valid) == 1.0f exactly, so it coincides bit for bit
//     with the reference's plain branch and no block-level `.any()` sync is needed.
#pragma once
#include "fsg_common.cuh"

namespace fsg {

struct Grid {
  const float* src;  // input grid (may hold NaN); src row 0 is GLOBAL row `row_off`
  int64_t h, w, ld;  // h = GLOBAL number of rows (edge rules), w columns, ld row stride
  int64_t row_off = 0;  // row-band shards: global row of src[0]; the band must hold every row the
                        // pass touches after mirroring / clamping at the GLOBAL edges
};

// pass A (axis 0) for global output rows [oy0, oy0+oh); tv/tw: f32 planes oh x w (ld = w).
int launch_box_axis0(const Grid& g, int size, float* tv, float* tw, int64_t oy0, int64_t oh, cudaStream_t s);
// need_scratch (void_fill_need_bytes(h, w) bytes, whole-grid void fill only, may be NULL): the pass is evaluated only
// on the (row, 256-column tile) pairs within `radius` columns of a NaN cell -- all the fill's axis-1 pass reads.
size_t void_fill_need_bytes(int64_t h, int64_t w);
int launch_gauss_axis0(const Grid& g, const double* taps_dev, int radius, float* tv, float* tw, const int* run_flag,
                       int64_t oy0, int64_t oh, cudaStream_t s, unsigned char* need_scratch = nullptr);

// pass B (axis 1) + combine.  mode 0: mean = tw>0 ? tv/tw : 0 -> out
//                            mode 1 (void fill, _nan_utils.py:655-667): where isnan(orig) & (sw > 0.5):
//                                    orig = sv / max(sw, 1e-6) (in place); sets *still_nan if NaN remains
enum { COMBINE_MEAN = 0, COMBINE_VOIDFILL = 1 };
int launch_box_axis1(const float* tv, const float* tw, int64_t h, int64_t w, int size, float* out, cudaStream_t s);
// void_tiles (COMBINE_VOIDFILL only, may be NULL): the first h * ceil(w / 256) bytes of the need scratch
// launch_gauss_axis0 filled (1 = the tile of that row holds a void cell)
int launch_gauss_axis1(const float* tv, const float* tw, int64_t h, int64_t w, const double* taps_dev, int radius,
                       int combine, float* out, const int* run_flag, int* still_nan, cudaStream_t s,
                       const unsigned char* void_tiles = nullptr);

// both box passes in one kernel (intermediate planes of a 16-row batch in shared memory), sizes 3..257: returns 1
// when launched (*rc = status), 0 when the size is out of range -- same values as axis0 + axis1 bit for bit
int launch_box_mean2d(const Grid& g, int size, float* out, int64_t oy0, int64_t oh, cudaStream_t s, int* rc);

// taps for sigma (device side, f64): w[0..radius]
int launch_gauss_taps(double sigma, int radius, double* taps_dev, cudaStream_t s);

}  // namespace fsg
