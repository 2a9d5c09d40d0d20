// Gradient family: hillshade / slope / curvature as ONE fused pass per algorithm.
//
// Reference arithmetic (every f32 op individually rounded, no FMA):
//   handle_nan_for_gradient  algorithms/_nan_utils.py:50-74   (NaN gap fill + np.gradient edge_order=2)
//   compute_hillshade_block  algorithms/_impl_hillshade.py:20-54
//   compute_slope_block      algorithms/_impl_slope.py:19-35
//   compute_curvature_block  algorithms/_impl_curvature.py:19-57
//
// Layout: a CTA of 256 threads owns a 64x32 output tile.  The gap-filled, z-scaled DEM tile plus a
// 2-px (hillshade/slope) or 3-px (curvature) halo is staged once in shared memory; curvature also
// stages the first derivatives on tile+2.  Global traffic: one read of the DEM (+halo re-reads that
// hit L2) and one write of the result -- 8 B/px (f32 out), 5 B/px (u8 out).
#include <stdlib.h>

#include "fsg_common.cuh"

namespace fsg {

constexpr int GT_W = 64;
constexpr int GT_H = 32;
constexpr int G_THREADS = 256;

struct AxisCoef {  // np.gradient(edge_order=2) coefficients for one axis, already f32
  float two_h;     // f32(2*h)           interior: (f[i+1]-f[i-1]) / two_h
  float inv_two_h; // 1/two_h when two_h is a power of two (the division is then an exact multiply), else 0
  float a0, b0, c0;  // first sample:   a0*f0 + b0*f1 + c0*f2
  float a1, b1, c1;  // last sample:    a1*f[n-3] + b1*f[n-2] + c1*f[n-1]
};

static AxisCoef make_axis(double h) {
  AxisCoef c;
  c.two_h = (float)(2.0 * h);
  int ex = 0;
  c.inv_two_h = (frexpf(fabsf(c.two_h), &ex) == 0.5f && ex > -100 && ex < 100) ? 1.0f / c.two_h : 0.f;
  c.a0 = (float)(-1.5 / h); c.b0 = (float)(2.0 / h); c.c0 = (float)(-0.5 / h);
  c.a1 = (float)(0.5 / h);  c.b1 = (float)(-2.0 / h); c.c1 = (float)(1.5 / h);
  return c;
}

struct GradParams {
  const float* dem;
  void* out;
  int64_t H, W, buf_row0, buf_rows, out_row0, out_rows, ld_in, ld_out;
  float zscale;
  AxisCoef y1, x1;  // first derivatives (|step|)
  AxisCoef y2, x2;  // second derivatives (signed step; curvature only)
  float lx, ly, lz, sgx, sgy;  // hillshade light vector / orientation signs
  // interior pixels work on the raw central differences d = f[i+1] - f[i-1] (see grad_interior): light vector and
  // signs folded with 1/(2h), squared reciprocal steps
  float kx, ky, ix2, iy2;
  int sub;                     // slope unit or curvature type
  double gw[5];                // sigma=1 Gaussian half taps (centre..outer)
  EncodeDev enc;
};

// NaN-aware sigma=1 Gaussian at one pixel (handle_nan_with_gaussian, mode='nearest'): separable,
// axis 0 first, f64 accumulation in scipy's symmetric order, f32 between the passes.
__device__ __noinline__ float gap_fill(const GradParams& p, int64_t gy, int64_t gx) {
  float colv[9], colw[9];
#pragma unroll 1
  for (int k = 0; k < 9; ++k) {
    int64_t x = clamp_index(gx + k - 4, p.W);
    auto at = [&](int64_t y, float* ok) {
      y = clamp_index(y, p.H) - p.buf_row0;
      y = y < 0 ? 0 : (y >= p.buf_rows ? p.buf_rows - 1 : y);
      float v = p.dem[y * p.ld_in + x];
      bool nan = v != v;
      *ok = nan ? 0.f : 1.f;
      return nan ? 0.f : v;
    };
    float o0;
    float v0 = at(gy, &o0);
    double sv = (double)v0 * p.gw[0];
    double sw = (double)o0 * p.gw[0];
#pragma unroll
    for (int j = 4; j >= 1; --j) {
      float oa, ob;
      float va = at(gy - j, &oa);
      float vb = at(gy + j, &ob);
      sv += ((double)va + (double)vb) * p.gw[j];
      sw += ((double)oa + (double)ob) * p.gw[j];
    }
    colv[k] = (float)sv;
    colw[k] = (float)sw;
  }
  double sv = (double)colv[4] * p.gw[0];
  double sw = (double)colw[4] * p.gw[0];
#pragma unroll
  for (int j = 4; j >= 1; --j) {
    sv += ((double)colv[4 - j] + (double)colv[4 + j]) * p.gw[j];
    sw += ((double)colw[4 - j] + (double)colw[4 + j]) * p.gw[j];
  }
  float fv = (float)sv, fw = (float)sw;
  return fw > 0.f ? fv / fw : 0.f;
}

// x^1.5 for x >= 0 as x*sqrt(x): two correctly rounded ops (<= 1.5 ulp; the reference's powf is not
// correctly rounded either) instead of the ~100-instruction powf
__device__ __forceinline__ float pow15(float x) { return x * sfu_sqrt(x); }

// derivative along one axis of a shared-memory plane; `g` is the global index along the axis,
// `n` the raster extent, `s` the element stride along that axis.
__device__ __forceinline__ float deriv(const float* f, int s, int64_t g, int64_t n, const AxisCoef& c) {
  if (g == 0) return (c.a0 * f[0] + c.b0 * f[s]) + c.c0 * f[2 * s];
  if (g == n - 1) return (c.a1 * f[-2 * s] + c.b1 * f[-s]) + c.c1 * f[0];
  float d = f[s] - f[-s];
  return c.inv_two_h != 0.f ? d * c.inv_two_h : d / c.two_h;   // exact reciprocal when 2h is a power of two
}

// curvature from first / second derivatives + display mapping (_impl_curvature.py:34-55), op for op
__device__ __forceinline__ float curv_result(const GradParams& p, float dy, float dx, float dyy, float dyx, float dxy,
                                             float dxx) {
  float k;
  if (p.sub == FSG_CURV_MEAN) {
    float pp = dx, q = dy, r = dxx, t = dyy;
    float s = (dxy + dyx) / 2.f;
    float den = pow15((1.f + pp * pp) + q * q);
    float num = ((1.f + q * q) * r - ((2.f * pp) * q) * s) + (1.f + pp * pp) * t;
    k = (-num) * sfu_rcp(2.f * den + 1e-10f);
  } else if (p.sub == FSG_CURV_GAUSSIAN) {
    float b = (1.f + dx * dx) + dy * dy;
    k = (dxx * dyy - dxy * dxy) * sfu_rcp(b * b);
  } else if (p.sub == FSG_CURV_PLANFORM) {
    float num = ((dy * dy) * dxx - ((2.f * dx) * dy) * dxy) + (dx * dx) * dyy;
    k = (-num) * sfu_rcp(pow15(dx * dx + dy * dy) + 1e-10f);
  } else {
    float num = ((dx * dx) * dxx + ((2.f * dx) * dy) * dxy) + (dy * dy) * dyy;
    float g2 = dx * dx + dy * dy;
    k = (-num) * sfu_rcp(g2 * pow15((1.f + dx * dx) + dy * dy) + 1e-10f);
  }
  // display value ((tanh(100 k) + 1) / 2)^(1/2.2): (tanh(z) + 1) / 2 = 1 / (1 + e^(-2 z)), so the value is
  // 2^(-c * log2(1 + 2^(-200 log2(e) k))) -- three SFU operations instead of tanhf, a division, log2f and exp2f, and
  // accurate in RELATIVE terms where tanh saturates at -1 (the f32 tanh of the reference is not; the parity
  // test compares that range before the gamma).  The divisions above are reciprocal-multiplies (<= 2 ulp).
  // k = +-inf / NaN behave like the reference: 1, 0, NaN.
  const float E = sfu_ex2(k * -288.53900817779268f);
  return sfu_ex2((float)(-1 / 2.2) * sfu_lg2(1.f + E));
}

// (approximate SFU forms and the arctangent polynomial: fsg_common.cuh)
template <int CLASS>
__device__ __forceinline__ float grad_result(const GradParams& p, float dy, float dx) {
  if (CLASS == 0) {
    float e = dx * p.sgx, n = dy * p.sgy;
    float hs = (((-e) * p.lx + (-n) * p.ly) + p.lz) * sfu_rsqrt((e * e + n * n) + 1.0f);
    return fminf(fmaxf(hs, 0.f), 1.f);
  } else {
    const float g = sfu_sqrt(dx * dx + dy * dy);
    if (p.sub == FSG_SLOPE_PERCENT) return g * 100.f;   // tan(arctan(g)) * 100 (_impl_slope.py:30-31)
    const float s = fast_atan_pos(g);
    if (p.sub == FSG_SLOPE_DEGREE) return s * (180.0f / 3.14159274101257324f);  // npy_rad2degf
    return s;
  }
}

// Interior pixels (both neighbours exist on both axes) from the raw central differences dyr = f[y+1] - f[y-1],
// dxr = f[x+1] - f[x-1]: the reciprocal steps, the orientation signs and the light vector are folded into
// constants (host, f64 -> f32) and the sums are fused multiply-adds -- ~12 instead of ~30 instructions per pixel, same
// 1e-5 / 1e-6 contract (measured: <= 3e-7 relative).  Used by the streaming AND the tile kernel, so a pixel does
// not depend on which of the two computed it (row-band calls == whole-raster call).
template <int CLASS>
__device__ __forceinline__ float grad_interior(const GradParams& p, float dyr, float dxr) {
  const float g2 = fmaf(dxr * dxr, p.ix2, (dyr * dyr) * p.iy2);   // dx^2 + dy^2
  if (CLASS == 0) {
    const float num = fmaf(dxr, p.kx, fmaf(dyr, p.ky, p.lz));    // (-e) lx + (-n) ly + lz
    return __saturatef(num * sfu_rsqrt(g2 + 1.0f));
  } else {
    const float g = sfu_sqrt(g2);
    if (p.sub == FSG_SLOPE_PERCENT) return g * 100.f;
    const float s = fast_atan_pos(g);
    if (p.sub == FSG_SLOPE_DEGREE) return s * (180.0f / 3.14159274101257324f);
    return s;
  }
}

template <int CLASS>
struct TileGeom {
  static constexpr int HALO = (CLASS == 2) ? 3 : 2;
  static constexpr int FW = GT_W + 2 * HALO + 1;  // +1: odd stride, fewer bank conflicts
  static constexpr int FH = GT_H + 2 * HALO;
  static constexpr int DW = GT_W + 4 + 1;
  static constexpr int DH = GT_H + 4;
};

// one 64x32 output tile with origin (ty0, tx0); all G_THREADS threads of the CTA take part
template <int CLASS>  // 0 hillshade, 1 slope, 2 curvature
__device__ __forceinline__ void grad_tile(const GradParams& p, int64_t ty0, int64_t tx0, float* F, unsigned char* M,
                                          float* DY, float* DX) {
  constexpr int HALO = TileGeom<CLASS>::HALO, FW = TileGeom<CLASS>::FW, FH = TileGeom<CLASS>::FH;
  constexpr int DW = TileGeom<CLASS>::DW, DH = TileGeom<CLASS>::DH;
  const int tid = threadIdx.x;

  // ---- stage gap-filled, scaled DEM ----
  for (int i = tid; i < FH * (GT_W + 2 * HALO); i += G_THREADS) {
    int fy = i / (GT_W + 2 * HALO), fx = i - fy * (GT_W + 2 * HALO);
    int64_t gy = ty0 - HALO + fy, gx = tx0 - HALO + fx;
    float v = 0.f;
    bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
    bool nan = false;
    if (inside) {
      int64_t by = gy - p.buf_row0;
      if (by >= 0 && by < p.buf_rows) {
        v = p.dem[by * p.ld_in + gx];
        nan = v != v;
        if (nan) v = gap_fill(p, gy, gx);
        v = v * p.zscale;
      }
    }
    F[fy * FW + fx] = v;
    int cy = fy - HALO, cx = fx - HALO;
    if (cy >= 0 && cy < GT_H && cx >= 0 && cx < GT_W) M[cy * GT_W + cx] = nan ? 1 : 0;
  }
  __syncthreads();

  if (CLASS == 2) {
    for (int i = tid; i < DH * (GT_W + 4); i += G_THREADS) {
      int dyi = i / (GT_W + 4), dxi = i - dyi * (GT_W + 4);
      int64_t gy = ty0 - 2 + dyi, gx = tx0 - 2 + dxi;
      float vy = 0.f, vx = 0.f;
      if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
        const float* f = &F[(dyi + 1) * FW + (dxi + 1)];
        vy = deriv(f, FW, gy, p.H, p.y1);
        vx = deriv(f, 1, gx, p.W, p.x1);
      }
      DY[dyi * DW + dxi] = vy;
      DX[dyi * DW + dxi] = vx;
    }
    __syncthreads();
  }

  for (int i = tid; i < GT_H * GT_W; i += G_THREADS) {
    int cy = i / GT_W, cx = i - cy * GT_W;
    int64_t gy = ty0 + cy, gx = tx0 + cx;
    if (gy >= p.out_row0 + p.out_rows || gx >= p.W) continue;
    float res;
    if (M[i]) {
      res = nanf("");
    } else if (CLASS == 0) {
      const float* f = &F[(cy + HALO) * FW + (cx + HALO)];
      if (gy > 0 && gy < p.H - 1 && gx > 0 && gx < p.W - 1) {
        res = grad_interior<0>(p, f[FW] - f[-FW], f[1] - f[-1]);
      } else {
        float dy = deriv(f, FW, gy, p.H, p.y1);
        float dx = deriv(f, 1, gx, p.W, p.x1);
        res = grad_result<0>(p, dy, dx);
      }
    } else if (CLASS == 1) {
      const float* f = &F[(cy + HALO) * FW + (cx + HALO)];
      if (gy > 0 && gy < p.H - 1 && gx > 0 && gx < p.W - 1) {
        res = grad_interior<1>(p, f[FW] - f[-FW], f[1] - f[-1]);
      } else {
        float dy = deriv(f, FW, gy, p.H, p.y1);
        float dx = deriv(f, 1, gx, p.W, p.x1);
        res = grad_result<1>(p, dy, dx);
      }
    } else {
      const float* gyp = &DY[(cy + 2) * DW + (cx + 2)];
      const float* gxp = &DX[(cy + 2) * DW + (cx + 2)];
      float dy = gyp[0], dx = gxp[0];
      float dyy = deriv(gyp, DW, gy, p.H, p.y2);
      float dyx = deriv(gyp, 1, gx, p.W, p.x2);
      float dxy = deriv(gxp, DW, gy, p.H, p.y2);
      float dxx = deriv(gxp, 1, gx, p.W, p.x2);
      res = curv_result(p, dy, dx, dyy, dyx, dxy, dxx);
    }
    store_out(p.out, (gy - p.out_row0) * p.ld_out + gx, res, p.enc);
  }
}

template <int CLASS>
__global__ void __launch_bounds__(G_THREADS) grad_kernel(const __grid_constant__ GradParams p) {
  __shared__ float F[TileGeom<CLASS>::FH * TileGeom<CLASS>::FW];
  __shared__ unsigned char M[GT_H * GT_W];
  __shared__ float DY[(CLASS == 2) ? TileGeom<CLASS>::DH * TileGeom<CLASS>::DW : 1];
  __shared__ float DX[(CLASS == 2) ? TileGeom<CLASS>::DH * TileGeom<CLASS>::DW : 1];
  grad_tile<CLASS>(p, p.out_row0 + (int64_t)blockIdx.y * GT_H, (int64_t)blockIdx.x * GT_W, F, M, DY, DX);
}

// ------------------------------------------------------------------------------------------
// Streaming variant for hillshade / slope (16-byte aligned rows).
// A CTA owns 1024 columns x 64 rows.  A thread owns four adjacent columns and marches down the rows
// with the previous, current and next row in registers plus three rows of loads in flight (float4
// loads; the two horizontal neighbours outside its four columns are scalar loads that hit L1).  The hot
// loop has no NaN / edge handling at all: a CTA that met a NaN anywhere in the rows it loaded redoes
// its block with the tile routine above (gap fill), and raster-edge pixels are redone with the
// one-sided np.gradient forms.  Arithmetic: the same op-for-op sequence as grad_tile (reference:
// _nan_utils.py:50-74, _impl_hillshade.py:20-54, _impl_slope.py:19-35).
// ------------------------------------------------------------------------------------------
constexpr int GS_THREADS = G_THREADS;
constexpr int GS_VEC = 4;
constexpr int GS_COLS = GS_THREADS * GS_VEC;
constexpr int GS_BAND = 64;
constexpr int GS_SLOTS = 6;   // row slots of the hillshade / slope streaming kernel (registers)
static_assert(GS_COLS % GT_W == 0 && GS_BAND % GT_H == 0, "the NaN redo walks whole tiles");

__device__ __forceinline__ float filled_at(const GradParams& p, int64_t gy, int64_t gx) {
  int64_t by = gy - p.buf_row0;
  by = by < 0 ? 0 : (by >= p.buf_rows ? p.buf_rows - 1 : by);
  float v = __ldg(p.dem + by * p.ld_in + gx);
  if (v != v) v = gap_fill(p, gy, gx);
  return v * p.zscale;
}

__device__ __forceinline__ float central(float hi, float lo, const AxisCoef& c) {
  float d = hi - lo;
  return c.inv_two_h != 0.f ? d * c.inv_two_h : d / c.two_h;
}

// POW2: every 2h is a power of two (the usual case: 1 m, 0.5 m, 2 m ... pixels), so each central difference is one
// subtraction and one exact multiplication -- no per-call test for the division form
template <bool POW2>
__device__ __forceinline__ float central_t(float hi, float lo, const AxisCoef& c) {
  if (POW2) return (hi - lo) * c.inv_two_h;
  return central(hi, lo, c);
}

// one raster-edge pixel from global memory (one-sided second-order forms of np.gradient)
template <int CLASS>
__device__ __noinline__ void edge_px(const GradParams& p, int64_t gy, int64_t gx) {
  float c0 = __ldg(p.dem + (gy - p.buf_row0) * p.ld_in + gx);
  float res;
  if (c0 != c0) {
    res = nanf("");
  } else {
    float dy, dx;
    if (gy == 0) dy = (p.y1.a0 * filled_at(p, 0, gx) + p.y1.b0 * filled_at(p, 1, gx)) + p.y1.c0 * filled_at(p, 2, gx);
    else if (gy == p.H - 1)
      dy = (p.y1.a1 * filled_at(p, p.H - 3, gx) + p.y1.b1 * filled_at(p, p.H - 2, gx)) + p.y1.c1 * filled_at(p, p.H - 1, gx);
    else dy = (filled_at(p, gy + 1, gx) - filled_at(p, gy - 1, gx)) / p.y1.two_h;
    if (gx == 0) dx = (p.x1.a0 * filled_at(p, gy, 0) + p.x1.b0 * filled_at(p, gy, 1)) + p.x1.c0 * filled_at(p, gy, 2);
    else if (gx == p.W - 1)
      dx = (p.x1.a1 * filled_at(p, gy, p.W - 3) + p.x1.b1 * filled_at(p, gy, p.W - 2)) + p.x1.c1 * filled_at(p, gy, p.W - 1);
    else dx = (filled_at(p, gy, gx + 1) - filled_at(p, gy, gx - 1)) / p.x1.two_h;
    res = grad_result<CLASS>(p, dy, dx);
  }
  store_out(p.out, (gy - p.out_row0) * p.ld_out + gx, res, p.enc);
}

struct GsRow {
  float v[GS_VEC + 2];   // columns c-1 .. c+4
};

// L2 prefetch of the 16 bytes a thread will load from row gy (no register, no scoreboard): the rows in flight in
// registers hide the L2 latency, these hide the DRAM latency behind them.
constexpr int GS_PF = 8;   // rows ahead of the register pipeline
__device__ __forceinline__ void gs_prefetch(const GradParams& p, int64_t gy, int64_t c) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dem + (gy - p.buf_row0) * p.ld_in + c));
}

__device__ __forceinline__ void gs_issue(const GradParams& p, int64_t gy, int64_t c, GsRow& r) {
  const float* row = p.dem + (gy - p.buf_row0) * p.ld_in;
  if (c + GS_VEC <= p.W) {
    float4 q = __ldg(reinterpret_cast<const float4*>(row + c));
    r.v[1] = q.x; r.v[2] = q.y; r.v[3] = q.z; r.v[4] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < GS_VEC; ++k) r.v[1 + k] = (c + k < p.W) ? __ldg(row + c + k) : 0.f;
  }
  r.v[0] = c > 0 ? __ldg(row + c - 1) : 0.f;
  r.v[5] = c + GS_VEC < p.W ? __ldg(row + c + GS_VEC) : 0.f;
}

template <int CLASS, bool POW2>
__global__ void __launch_bounds__(GS_THREADS, 3) grad_stream_kernel(const __grid_constant__ GradParams p) {
  __shared__ float F[TileGeom<CLASS>::FH * TileGeom<CLASS>::FW];
  __shared__ unsigned char M[GT_H * GT_W];
  const int64_t cx0 = (int64_t)blockIdx.x * GS_COLS;
  const int64_t c = cx0 + (int64_t)threadIdx.x * GS_VEC;
  const int64_t y0 = p.out_row0 + (int64_t)blockIdx.y * GS_BAND;
  const int64_t yend = p.out_row0 + p.out_rows;
  const int64_t y1 = y0 + GS_BAND < yend ? y0 + GS_BAND : yend;
  const int64_t last_needed = y1 < p.H ? y1 : p.H - 1;   // last row any pixel of the block reads
  const bool live = c < p.W;
  float nanprobe = 0.f;   // becomes NaN as soon as one loaded value is NaN (or +-Inf)
  if (live) {
    // GS_SLOTS row slots rotate through the roles (previous, current, next, GS_SLOTS - 4 rows whose loads are in
    // flight, the row being requested): the loop is unrolled GS_SLOTS times, so the rotation is a renaming, not
    // register moves.  (10 slots / 8 rows in flight at 2 CTAs per SM were measured: slower, 0.64 -> 0.73 ms.)
    // (Handing the neighbour columns across lanes by shuffle instead of the two scalar loads was measured: slower,
    // slope 0.78 -> 0.93 ms at 16384^2.)
    constexpr int NS = GS_SLOTS, AHEAD = GS_SLOTS - 2;   // row y + AHEAD is requested while row y is computed
    GsRow r[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q)
#pragma unroll
      for (int k = 0; k < GS_VEC + 2; ++k) r[q].v[k] = 0.f;
    if (y0 > 0) gs_issue(p, y0 - 1, c, r[0]);
#pragma unroll
    for (int q = 0; q < AHEAD; ++q)
      if (y0 + q <= last_needed) gs_issue(p, y0 + q, c, r[1 + q]);
#pragma unroll
    for (int q = AHEAD; q < AHEAD + GS_PF; ++q)
      if (y0 + q <= last_needed) gs_prefetch(p, y0 + q, c);
    const bool vec_out = (p.ld_out % 4 == 0) && (c + GS_VEC <= p.W);
    const bool scale = p.zscale != 1.0f;
    nanprobe += ((r[0].v[1] + r[0].v[2]) + (r[0].v[3] + r[0].v[4]));
#pragma unroll 1
    for (int64_t yb = y0; yb < y1; yb += NS) {
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int64_t y = yb + j;
        if (y < y1) {
          GsRow& prev = r[j];
          GsRow& cur = r[(j + 1) % NS];
          GsRow& next = r[(j + 2) % NS];
          if (y + AHEAD <= last_needed) gs_issue(p, y + AHEAD, c, r[(j + 1 + AHEAD) % NS]);
          if (y + AHEAD + GS_PF <= last_needed) gs_prefetch(p, y + AHEAD + GS_PF, c);
          float res[GS_VEC];
#pragma unroll
          for (int k = 0; k < GS_VEC; ++k) {
            float up = prev.v[k + 1], dn = next.v[k + 1], lf = cur.v[k], rt = cur.v[k + 2];
            if (scale) { up = up * p.zscale; dn = dn * p.zscale; lf = lf * p.zscale; rt = rt * p.zscale; }
            res[k] = grad_interior<CLASS>(p, dn - up, rt - lf);
          }
          nanprobe += ((cur.v[0] + cur.v[1]) + (cur.v[2] + cur.v[3])) + (cur.v[4] + cur.v[5]);
          nanprobe *= 0.f;
          const int64_t o = (y - p.out_row0) * p.ld_out + c;
          if (vec_out && p.enc.kind == FSG_OUT_F32) {
            *reinterpret_cast<float4*>((float*)p.out + o) = make_float4(res[0], res[1], res[2], res[3]);
          } else if (vec_out && p.enc.kind == FSG_OUT_U8) {
            uchar4 q = make_uchar4((unsigned char)(int)encode_dn(res[0], p.enc), (unsigned char)(int)encode_dn(res[1], p.enc),
                                   (unsigned char)(int)encode_dn(res[2], p.enc), (unsigned char)(int)encode_dn(res[3], p.enc));
            *reinterpret_cast<uchar4*>((uint8_t*)p.out + o) = q;
          } else {
#pragma unroll
            for (int k = 0; k < GS_VEC; ++k)
              if (c + k < p.W) store_out(p.out, o + k, res[k], p.enc);
          }
          if (y + 1 == y1) {   // the row below the block was read as `next` of the last row
            nanprobe += ((next.v[1] + next.v[2]) + (next.v[3] + next.v[4]));
            nanprobe *= 0.f;
          }
        }
      }
    }
  }
  // ---- cold paths ----
  if (__syncthreads_or((int)(nanprobe != nanprobe))) {
    // a NaN (NoData) somewhere in the rows this block loaded: redo the block tile by tile with gap fill
    for (int64_t ty = y0; ty < y1; ty += GT_H) {
      for (int64_t tx = cx0; tx < cx0 + GS_COLS && tx < p.W; tx += GT_W) {
        __syncthreads();
        grad_tile<CLASS>(p, ty, tx, F, M, nullptr, nullptr);
      }
    }
    return;
  }
  if (!live) return;
  if (y0 == 0) {
    for (int k = 0; k < GS_VEC; ++k) if (c + k < p.W) edge_px<CLASS>(p, 0, c + k);
  }
  if (y1 == p.H && p.H > 1) {
    for (int k = 0; k < GS_VEC; ++k) if (c + k < p.W) edge_px<CLASS>(p, p.H - 1, c + k);
  }
  if (c == 0) {
    for (int64_t y = y0; y < y1; ++y) edge_px<CLASS>(p, y, 0);
  }
  if (c <= p.W - 1 && p.W - 1 < c + GS_VEC) {
    for (int64_t y = y0; y < y1; ++y) edge_px<CLASS>(p, y, p.W - 1);
  }
}


// ------------------------------------------------------------------------------------------
// Streaming curvature (16-byte aligned rows): same scheme with a 5 x 5 footprint.  A thread owns four adjacent
// columns and keeps rows y-2 .. y+2 of columns c-2 .. c+5 in registers (three more rows of loads in flight);
// first derivatives are formed where they are needed and the second derivatives from them -- the op sequence of
// np.gradient applied twice (_impl_curvature.py:23-33: |step| for the first, signed step for the second pass).
// The hot loop uses the central forms everywhere and clamped loads; the two-pixel frame of the raster (where
// np.gradient's one-sided forms enter a first or a second derivative) is redone afterwards tile by tile with
// grad_tile<2>, and so is the whole block when a NaN / Inf was loaded (gap fill).
// ------------------------------------------------------------------------------------------
struct CsRow {
  float v[GS_VEC + 4];   // columns c-2 .. c+5
};

__device__ __forceinline__ void cs_issue(const GradParams& p, int64_t gy, int64_t c, CsRow& r) {
  gy = gy < 0 ? 0 : (gy > p.H - 1 ? p.H - 1 : gy);           // clamped: only frame pixels see the difference
  int64_t by = gy - p.buf_row0;                               // (row bands: the look-ahead rows may lie outside the buffer)
  by = by < 0 ? 0 : (by > p.buf_rows - 1 ? p.buf_rows - 1 : by);
  const float* row = p.dem + by * p.ld_in;
  if (c + GS_VEC <= p.W) {
    float4 q = __ldg(reinterpret_cast<const float4*>(row + c));
    r.v[2] = q.x; r.v[3] = q.y; r.v[4] = q.z; r.v[5] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < GS_VEC; ++k) r.v[2 + k] = (c + k < p.W) ? __ldg(row + c + k) : 0.f;
  }
  if (c >= 2) {
    float2 l = __ldg(reinterpret_cast<const float2*>(row + c - 2));
    r.v[0] = l.x; r.v[1] = l.y;
  } else {
    r.v[0] = r.v[1] = 0.f;
  }
  if (c + GS_VEC + 2 <= p.W) {
    float2 h = __ldg(reinterpret_cast<const float2*>(row + c + GS_VEC));
    r.v[6] = h.x; r.v[7] = h.y;
  } else {
    r.v[6] = c + GS_VEC < p.W ? __ldg(row + c + GS_VEC) : 0.f;
    r.v[7] = 0.f;
  }
}

template <bool POW2>
__global__ void __launch_bounds__(GS_THREADS, 2) curv_stream_kernel(const __grid_constant__ GradParams p) {
  __shared__ float F[TileGeom<2>::FH * TileGeom<2>::FW];
  __shared__ unsigned char M[GT_H * GT_W];
  __shared__ float DY[TileGeom<2>::DH * TileGeom<2>::DW];
  __shared__ float DX[TileGeom<2>::DH * TileGeom<2>::DW];
  const int64_t cx0 = (int64_t)blockIdx.x * GS_COLS;
  const int64_t c = cx0 + (int64_t)threadIdx.x * GS_VEC;
  const int64_t y0 = p.out_row0 + (int64_t)blockIdx.y * GS_BAND;
  const int64_t yend = p.out_row0 + p.out_rows;
  const int64_t y1 = y0 + GS_BAND < yend ? y0 + GS_BAND : yend;
  const bool live = c < p.W;
  float nanprobe = 0.f;
  if (live) {
    CsRow r0, r1, r2, r3, r4, q1, q2, q3;   // rows y-2 .. y+2, then rows whose loads are in flight
    cs_issue(p, y0 - 2, c, r0);
    cs_issue(p, y0 - 1, c, r1);
    cs_issue(p, y0, c, r2);
    cs_issue(p, y0 + 1, c, r3);
    cs_issue(p, y0 + 2, c, r4);
    cs_issue(p, y0 + 3, c, q1);
    cs_issue(p, y0 + 4, c, q2);
    const bool vec_out = (p.ld_out % 4 == 0) && (c + GS_VEC <= p.W);
#pragma unroll
    for (int i = 0; i < GS_VEC + 4; ++i) nanprobe += (r0.v[i] + r1.v[i]) + (r2.v[i] + r3.v[i]);
    nanprobe *= 0.f;
#pragma unroll 1
    for (int64_t y = y0; y < y1; ++y) {
      cs_issue(p, y + 5, c, q3);
      {   // L2 prefetch GS_PF rows ahead of the register pipeline (clamped like the loads)
        int64_t by = y + 5 + GS_PF - p.buf_row0;
        by = by < 0 ? 0 : (by > p.buf_rows - 1 ? p.buf_rows - 1 : by);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dem + by * p.ld_in + c));
      }
      float res[GS_VEC];
#pragma unroll
      for (int k = 0; k < GS_VEC; ++k) {
        const int i = k + 2;
        // first derivatives (|step|): dy at (y-1, y, y+1; x) and (y; x-1, x+1); dx at (y; x-1, x, x+1) and (y-1, y+1; x)
        const float dy_m = central_t<POW2>(r2.v[i], r0.v[i], p.y1);
        const float dy_0 = central_t<POW2>(r3.v[i], r1.v[i], p.y1);
        const float dy_p = central_t<POW2>(r4.v[i], r2.v[i], p.y1);
        const float dy_l = central_t<POW2>(r3.v[i - 1], r1.v[i - 1], p.y1);
        const float dy_r = central_t<POW2>(r3.v[i + 1], r1.v[i + 1], p.y1);
        const float dx_l = central_t<POW2>(r2.v[i], r2.v[i - 2], p.x1);
        const float dx_0 = central_t<POW2>(r2.v[i + 1], r2.v[i - 1], p.x1);
        const float dx_r = central_t<POW2>(r2.v[i + 2], r2.v[i], p.x1);
        const float dx_m = central_t<POW2>(r1.v[i + 1], r1.v[i - 1], p.x1);
        const float dx_p = central_t<POW2>(r3.v[i + 1], r3.v[i - 1], p.x1);
        // second derivatives (signed step)
        const float dyy = central_t<POW2>(dy_p, dy_m, p.y2);
        const float dyx = central_t<POW2>(dy_r, dy_l, p.x2);
        const float dxy = central_t<POW2>(dx_p, dx_m, p.y2);
        const float dxx = central_t<POW2>(dx_r, dx_l, p.x2);
        res[k] = curv_result(p, dy_0, dx_0, dyy, dyx, dxy, dxx);
      }
#pragma unroll
      for (int i = 0; i < GS_VEC + 4; ++i) nanprobe += r4.v[i];
      nanprobe *= 0.f;
      const int64_t o = (y - p.out_row0) * p.ld_out + c;
      if (vec_out && p.enc.kind == FSG_OUT_F32) {
        *reinterpret_cast<float4*>((float*)p.out + o) = make_float4(res[0], res[1], res[2], res[3]);
      } else if (vec_out && p.enc.kind == FSG_OUT_U8) {
        uchar4 q = make_uchar4((unsigned char)(int)encode_dn(res[0], p.enc), (unsigned char)(int)encode_dn(res[1], p.enc),
                               (unsigned char)(int)encode_dn(res[2], p.enc), (unsigned char)(int)encode_dn(res[3], p.enc));
        *reinterpret_cast<uchar4*>((uint8_t*)p.out + o) = q;
      } else {
#pragma unroll
        for (int k = 0; k < GS_VEC; ++k)
          if (c + k < p.W) store_out(p.out, o + k, res[k], p.enc);
      }
      r0 = r1; r1 = r2; r2 = r3; r3 = r4; r4 = q1; q1 = q2; q2 = q3;
    }
#pragma unroll
    for (int i = 0; i < GS_VEC + 4; ++i) nanprobe += r3.v[i] + r4.v[i];   // rows y1, y1+1 (read by the last rows)
    nanprobe *= 0.f;
  }
  // ---- cold paths: NaN in the block -> everything again with gap fill; otherwise only the raster frame ----
  const bool redo_all = __syncthreads_or((int)(nanprobe != nanprobe)) != 0;
  for (int64_t ty = y0; ty < y1; ty += GT_H) {
    for (int64_t tx = cx0; tx < cx0 + GS_COLS && tx < p.W; tx += GT_W) {
      const bool frame = ty < 2 || ty + GT_H > p.H - 2 || tx < 2 || tx + GT_W > p.W - 2;
      if (!redo_all && !frame) continue;
      __syncthreads();
      grad_tile<2>(p, ty, tx, F, M, DY, DX);
    }
  }
}

static int check_window(const fsg_window* w, int need_halo, const char* who) {
  if (!w) return fail(FSG_E_INVALID, "%s: window is NULL", who);
  if (w->H_global < 3 || w->W < 3)
    return fail(FSG_E_INVALID,
                "%s: Shape of array too small to calculate a numerical gradient, at least (edge_order + 1) elements are required.", who);
  if (w->out_rows < 0 || w->out_row0 < 0 || w->out_row0 + w->out_rows > w->H_global)
    return fail(FSG_E_INVALID, "%s: output rows [%lld,+%lld) outside raster of %lld rows", who,
                (long long)w->out_row0, (long long)w->out_rows, (long long)w->H_global);
  int64_t lo = w->out_row0 - need_halo; if (lo < 0) lo = 0;
  int64_t hi = w->out_row0 + w->out_rows + need_halo; if (hi > w->H_global) hi = w->H_global;
  if (w->buf_row0 > lo || w->buf_row0 + w->buf_rows < hi)
    return fail(FSG_E_INVALID, "%s: buffer rows [%lld,%lld) do not cover the %d-row halo [%lld,%lld)", who,
                (long long)w->buf_row0, (long long)(w->buf_row0 + w->buf_rows), need_halo, (long long)lo, (long long)hi);
  if (w->ld_in < w->W || w->ld_out < w->W) return fail(FSG_E_INVALID, "%s: row stride smaller than width", who);
  return FSG_OK;
}

static int run_grad(int cls, const float* dem, void* out, const fsg_window* win, GradParams& p,
                    const fsg_encode* enc, void* stream, const char* who) {
  int halo = (cls == 2) ? 3 : 2;
  int rc = check_window(win, halo, who);
  if (rc) return rc;
  if (!dem || !out) return fail(FSG_E_INVALID, "%s: NULL buffer", who);
  p.dem = dem; p.out = out;
  p.H = win->H_global; p.W = win->W; p.buf_row0 = win->buf_row0; p.buf_rows = win->buf_rows;
  p.out_row0 = win->out_row0; p.out_rows = win->out_rows; p.ld_in = win->ld_in; p.ld_out = win->ld_out;
  p.enc = make_encode(enc);
  if (gauss_half_taps(1.0, p.gw, 4) != 4) return fail(FSG_E_INVALID, "%s: gaussian taps", who);
  if (p.out_rows == 0) return FSG_OK;
  dim3 grid((unsigned)((p.W + GT_W - 1) / GT_W), (unsigned)((p.out_rows + GT_H - 1) / GT_H));
  cudaStream_t s = (cudaStream_t)stream;
  int slot = prof_begin(PROF_GRADIENT, s);
  // streaming kernel: aligned rows, whole buffer rows available for every row it touches
  const bool stream_ok = cls != 2 && (((uintptr_t)p.dem & 15) == 0) && (p.ld_in % 4 == 0) &&
                         (((uintptr_t)p.out & 15) == 0) && p.buf_row0 <= (p.out_row0 > 0 ? p.out_row0 - 1 : 0) &&
                         p.buf_row0 + p.buf_rows >= (p.out_row0 + p.out_rows < p.H ? p.out_row0 + p.out_rows + 1 : p.H) &&
                         !getenv("FSG_GRAD_TILED");
  // curvature: the buffer must hold the two rows above / below the output rows (clamped at the raster edge)
  const bool cstream_ok = cls == 2 && (((uintptr_t)p.dem & 15) == 0) && (p.ld_in % 4 == 0) &&
                          (((uintptr_t)p.out & 15) == 0) && p.H >= 8 && p.W >= 8 &&
                          p.buf_row0 <= (p.out_row0 > 2 ? p.out_row0 - 2 : 0) &&
                          p.buf_row0 + p.buf_rows >= (p.out_row0 + p.out_rows + 2 < p.H ? p.out_row0 + p.out_rows + 2 : p.H) &&
                          !getenv("FSG_GRAD_TILED");
  if (cstream_ok) {
    dim3 sgrid((unsigned)((p.W + GS_COLS - 1) / GS_COLS), (unsigned)((p.out_rows + GS_BAND - 1) / GS_BAND));
    const bool pow2 = p.y1.inv_two_h != 0.f && p.x1.inv_two_h != 0.f && p.y2.inv_two_h != 0.f && p.x2.inv_two_h != 0.f;
    if (pow2) curv_stream_kernel<true><<<sgrid, GS_THREADS, 0, s>>>(p);
    else curv_stream_kernel<false><<<sgrid, GS_THREADS, 0, s>>>(p);
  } else if (stream_ok) {
    dim3 sgrid((unsigned)((p.W + GS_COLS - 1) / GS_COLS), (unsigned)((p.out_rows + GS_BAND - 1) / GS_BAND));
    const bool pow2 = p.y1.inv_two_h != 0.f && p.x1.inv_two_h != 0.f;
    if (cls == 0 && pow2) grad_stream_kernel<0, true><<<sgrid, GS_THREADS, 0, s>>>(p);
    else if (cls == 0) grad_stream_kernel<0, false><<<sgrid, GS_THREADS, 0, s>>>(p);
    else if (pow2) grad_stream_kernel<1, true><<<sgrid, GS_THREADS, 0, s>>>(p);
    else grad_stream_kernel<1, false><<<sgrid, GS_THREADS, 0, s>>>(p);
  } else if (cls == 0) grad_kernel<0><<<grid, G_THREADS, 0, s>>>(p);
  else if (cls == 1) grad_kernel<1><<<grid, G_THREADS, 0, s>>>(p);
  else grad_kernel<2><<<grid, G_THREADS, 0, s>>>(p);
  prof_end(slot, s);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

}  // namespace fsg

extern "C" {

int fsg_hillshade(const float* dem, void* out, const fsg_window* win, double azimuth, double altitude,
                  double z_factor, double pixel_size, double psx, double psy, const fsg_encode* enc, void* stream) {
  using namespace fsg;
  GradParams p{};
  double sy, sx;
  resolve_steps(pixel_size, psx, psy, false, &sy, &sx);
  p.y1 = make_axis(sy); p.x1 = make_axis(sx); p.y2 = p.y1; p.x2 = p.x1;
  p.zscale = (float)(is_none(z_factor) ? 1.0 : z_factor);
  const double d2r = 3.14159265358979323846 / 180.0;
  double alt = altitude * d2r, az = azimuth * d2r;
  p.lx = (float)(sin(az) * cos(alt)); p.ly = (float)(cos(az) * cos(alt)); p.lz = (float)sin(alt);
  p.sgx = (is_none(psx) || psx >= 0.0) ? 1.f : -1.f;
  p.sgy = (is_none(psy) || psy >= 0.0) ? 1.f : -1.f;
  p.kx = (float)(-(double)p.sgx * (double)p.lx / (2.0 * sx)); p.ky = (float)(-(double)p.sgy * (double)p.ly / (2.0 * sy));
  p.ix2 = (float)(1.0 / (4.0 * sx * sx)); p.iy2 = (float)(1.0 / (4.0 * sy * sy));
  return run_grad(0, dem, out, win, p, enc, stream, "fsg_hillshade");
}

int fsg_slope(const float* dem, void* out, const fsg_window* win, int unit, double pixel_size, double psx,
              double psy, const fsg_encode* enc, void* stream) {
  using namespace fsg;
  if (unit < 0 || unit > 2) return fail(FSG_E_INVALID, "fsg_slope: unknown unit %d", unit);
  GradParams p{};
  double sy, sx;
  resolve_steps(pixel_size, psx, psy, false, &sy, &sx);
  p.y1 = make_axis(sy); p.x1 = make_axis(sx); p.y2 = p.y1; p.x2 = p.x1;
  p.zscale = 1.f; p.sub = unit;
  p.ix2 = (float)(1.0 / (4.0 * sx * sx)); p.iy2 = (float)(1.0 / (4.0 * sy * sy));
  return run_grad(1, dem, out, win, p, enc, stream, "fsg_slope");
}

int fsg_curvature(const float* dem, void* out, const fsg_window* win, int curvature_type, double pixel_size,
                  double psx, double psy, const fsg_encode* enc, void* stream) {
  using namespace fsg;
  if (curvature_type < 0 || curvature_type > 3) return fail(FSG_E_INVALID, "fsg_curvature: unknown type %d", curvature_type);
  GradParams p{};
  // first derivatives use |step| (handle_nan_for_gradient), second derivatives the signed step
  double sy2, sx2;
  resolve_steps(pixel_size, psx, psy, true, &sy2, &sx2);
  p.y1 = make_axis(fabs(sy2) < 1e-9 ? 1.0 : fabs(sy2));
  p.x1 = make_axis(fabs(sx2) < 1e-9 ? 1.0 : fabs(sx2));
  p.y2 = make_axis(sy2); p.x2 = make_axis(sx2);
  p.zscale = 1.f; p.sub = curvature_type;
  return run_grad(2, dem, out, win, p, enc, stream, "fsg_curvature");
}

}  // extern "C"
