// Topographic openness (Yokoyama et al. 2002) -- compute_openness_vectorized
// (algorithms/_impl_openness.py:31-132) as one fused gather kernel.
//
// Per pixel: D azimuths x <=10 ray samples; sample offset (round(r cos a), round(r sin a)) with
// Python's half-even round; sample value = NaN->0 raster, edge-replicated; sample validity = inside
// the raster and not NaN (and the centre not NaN).  Per azimuth the reference keeps
// max/min over samples of arctan(dz/dist); arctan is monotone, so the kernel keeps the extreme of the
// f32 quotient dz/dist (same individually rounded ops) and applies arctan once per azimuth.
// Then: zenith/nadir angle, mean over azimuths with >=1 valid sample, /(pi/2), clip, gamma,
// optional display stretch (tile/dask_bridge.py:173-187), NaN restore, optional integer encoding.
//
// Traffic: the 80 gathers of a CTA tile are 80 shifted copies of the tile, served by L1/L2; DRAM sees
// the DEM about once when CTAs sweep the raster in row-major order (a +-256-row band of a 32768-wide
// raster is 67 MB < 126 MB L2).  Algorithmic bytes: 8 B/px.
#include <stdlib.h>

#include "fsg_common.cuh"

namespace fsg {

constexpr int OP_MAX_SAMPLES = 64 * 10;
constexpr int OPI_ROWS = 8;    // rows of an interior tile (128 columns x 8 rows, four pixels per thread)

struct OpenParams {
  const float* dem;
  void* out;
  int64_t H, W, buf_row0, buf_rows, out_row0, out_rows, ld_in, ld_out;
  int n_dirs;
  int negative;
  int stretch;
  float stretch_lo, stretch_scale, stretch_rinv;
  EncodeDev enc;
  int64_t rx0, ry0, rx1, ry1;   // openness_kernel: the rectangle of pixels this launch covers (global coordinates)
};

struct OpenSample {
  short ox, oy;
  float dist;  // f32(max(hypot(ox*sx, oy*sy), 1e-9))
  int off;     // oy * ld_in + ox (elements)
  float rinv;  // f32(1 / dist): the division by the constant distance is a multiplication (<= 1.5 ulp; the contract
               // of this output is 1e-5 relative / 1e-6 absolute, tests/test_gpu_parity.py asserts the measured error)
};

// Passed BY VALUE as a kernel parameter (read through the constant bank, broadcast to the warp):
// no global/constant symbol is written, so concurrent calls from several host threads are safe.
struct OpenTable {
  int dir_start[65];
  OpenSample s[OP_MAX_SAMPLES];
};

// The interior kernel's view of the table: exactly OPI_NS slots per direction (the reference's ten ring distances; a
// direction with fewer distinct distances is padded), byte offsets, so that the sample loop is fully unrolled and
// its table reads are constant-bank loads at immediate offsets.  A padding slot has rinv = NaN: its quotient is NaN,
// which fmaxf / fminf ignore exactly like a NaN sample.
constexpr int OPI_NS = 10;
constexpr int OPI_PANEL = 64;   // tiles (of 128 columns) per column panel of the interior kernel's block order
struct OpenSlot {
  int off_bytes;   // (oy * ld_in + ox) * 4
  float rinv;
};
struct OpenFixed {
  OpenSlot s[64 * OPI_NS];
};

// mean zenith / nadir angle -> display value: / (pi/2), clip, gamma 1/2.2, optional stretch (same ops in both kernels)
__device__ __forceinline__ float open_finish(const OpenParams& p, float asum, float acnt) {
  float o = asum * sfu_rcp(fmaxf(acnt, 1.f));
  o = __saturatef(o * (float)(2.0 / 3.14159265358979323846));
  float res = fast_gamma22(o);
  if (p.stretch) res = fmaxf((res - p.stretch_lo) * p.stretch_rinv, 0.f);
  return res;
}

__global__ void __launch_bounds__(256) openness_kernel(OpenParams p, const __grid_constant__ OpenTable tab, int fast_halo) {
  if (fast_halo >= 0) {   // interior 128 x 8 tiles are done by openness_interior_kernel
    const int64_t px = p.rx0 + (int64_t)blockIdx.x * 64, py = p.ry0 + (int64_t)blockIdx.y * 4;
    const int64_t tx0 = px / 128 * 128, ty0 = p.out_row0 + (py - p.out_row0) / OPI_ROWS * OPI_ROWS;
    if (tx0 >= fast_halo && tx0 + 128 + fast_halo <= p.W && ty0 >= fast_halo && ty0 + OPI_ROWS + fast_halo <= p.H &&
        ty0 + OPI_ROWS <= p.out_row0 + p.out_rows)
      return;
  }
  const int64_t x = p.rx0 + (int64_t)blockIdx.x * 64 + (threadIdx.x & 63);
  const int64_t y = p.ry0 + (int64_t)blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= p.rx1 || y >= p.ry1) return;
  const float* base = p.dem - p.buf_row0 * p.ld_in;  // address of global row 0
  const float c = base[y * p.ld_in + x];
  float res;
  if (c != c) {
    res = nanf("");
  } else {
    const float half_pi = (float)(3.14159265358979323846 / 2);
    float asum = 0.f, acnt = 0.f;
    for (int d = 0; d < p.n_dirs; ++d) {
      float ext = 0.f;
      bool seen = false;
      for (int k = tab.dir_start[d]; k < tab.dir_start[d + 1]; ++k) {
        const OpenSample sm = tab.s[k];
        int64_t sy = y + sm.oy, sx = x + sm.ox;
        if (sy < 0 || sy >= p.H || sx < 0 || sx >= p.W) continue;  // padded validity = False
        float v = base[sy * p.ld_in + sx];
        if (v != v) continue;
        float t = (v - c) * sm.rinv;
        if (!seen) ext = t;
        else ext = p.negative ? fminf(ext, t) : fmaxf(ext, t);
        seen = true;
      }
      if (seen) {
        // np.maximum(ext, angle) starting from -pi/2 (f32): atan never drops below it
        float a = fast_atan(ext);
        a = p.negative ? fminf(half_pi, a) : fmaxf(-half_pi, a);
        float dang = p.negative ? half_pi + a : half_pi - a;
        asum = asum + dang;
        acnt = acnt + 1.f;
      }
    }
    res = open_finish(p, asum, acnt);
  }
  store_out(p.out, (y - p.out_row0) * p.ld_out + x, res, p.enc);
}


// Interior tiles (every sample of every pixel lies inside the raster): no bounds checks, linear sample offsets, the
// division by the constant distance as a multiplication.  Four pixels per thread (2 columns 64 apart x 2 rows 4
// apart).  The ncu profile of round 1's version showed the kernel issue bound (90 % issue-slot use, L1 28 %, L2
// 31 %): 13 instructions per sample and pixel, 10 of 30 per sample on 64-bit address arithmetic.  Now: one add per
// row pointer serves two loads, the quotient is one multiply.  (Eight pixels per thread, 128 x 16 tiles: measured
// slower, 34 -> 52 ms at 32768^2 -- fewer resident warps, larger L1 footprint.)
template <bool NEG>
__global__ void __launch_bounds__(256) openness_interior_kernel(const __grid_constant__ OpenParams p,
                                                                const __grid_constant__ OpenFixed tab, int halo) {
  // CTAs are handed out in launch order; walking the raster in panels of OPI_PANEL tiles (8192 columns) keeps the rows
  // the resident CTAs sample (+- max_distance around ~5 tile rows) inside the L2: 18 MB per panel instead of 72 MB for
  // the full width of a 32768-column raster
  int64_t tx, ty;
  {
    const int64_t ntx = gridDim.x, nty = gridDim.y;
    const int64_t b = (int64_t)blockIdx.y * ntx + blockIdx.x;
    const int64_t nfull = ntx / OPI_PANEL, full = nfull * OPI_PANEL * nty;
    if (b < full) {
      const int64_t panel = b / (OPI_PANEL * nty), r = b - panel * (OPI_PANEL * nty);
      ty = r / OPI_PANEL;
      tx = panel * OPI_PANEL + (r - ty * OPI_PANEL);
    } else {
      const int64_t r = b - full, wl = ntx - nfull * OPI_PANEL;
      ty = r / wl;
      tx = nfull * OPI_PANEL + (r - ty * wl);
    }
  }
  const int64_t x0 = tx * 128, y0 = p.out_row0 + ty * OPI_ROWS;
  // tiles that touch the border band are left to openness_kernel (launched over the same area)
  const bool interior = x0 >= halo && x0 + 128 + halo <= p.W && y0 >= halo && y0 + OPI_ROWS + halo <= p.H &&
                        y0 + OPI_ROWS <= p.out_row0 + p.out_rows;
  if (!interior) return;
  const int64_t x = x0 + (threadIdx.x & 63);
  const int64_t y = y0 + (threadIdx.x >> 6);
  constexpr int NR = OPI_ROWS / 4, NP = 2 * NR;
  const char* pr[NR];   // this thread's rows y, y + 4 at column x (byte pointers: a sample is one 64-bit add away)
#pragma unroll
  for (int r = 0; r < NR; ++r) pr[r] = reinterpret_cast<const char*>(p.dem + (y + 4 * r - p.buf_row0) * p.ld_in + x);
  float c[NP], asum[NP], acnt[NP];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    c[2 * r] = __ldg(reinterpret_cast<const float*>(pr[r]));
    c[2 * r + 1] = __ldg(reinterpret_cast<const float*>(pr[r]) + 64);
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) { asum[j] = 0.f; acnt[j] = 0.f; }
  const float half_pi = (float)(3.14159265358979323846 / 2);
  const float start = NEG ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
  for (int d = 0; d < p.n_dirs; ++d) {
    float e[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) e[j] = start;
    const OpenSlot* slot = tab.s + d * OPI_NS;
#pragma unroll
    for (int k = 0; k < OPI_NS; ++k) {
      const int64_t off = (int64_t)slot[k].off_bytes;
      const float rinv = slot[k].rinv;
      float v[NP];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const float* ps = reinterpret_cast<const float*>(pr[r] + off);
        v[2 * r] = __ldg(ps);
        v[2 * r + 1] = __ldg(ps + 64);
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float q = (v[j] - c[j]) * rinv;
        // a NaN sample gives a NaN quotient, which fminf / fmaxf ignore: invalid samples drop out by themselves
        e[j] = NEG ? fminf(e[j], q) : fmaxf(e[j], q);
      }
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (e[j] != start) {   // at least one valid sample in this direction
        float a = fast_atan(e[j]);
        a = NEG ? fminf(half_pi, a) : fmaxf(-half_pi, a);
        asum[j] = asum[j] + (NEG ? half_pi + a : half_pi - a);
        acnt[j] = acnt[j] + 1.f;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    float res = open_finish(p, asum[j], acnt[j]);
    if (c[j] != c[j]) res = nanf("");
    store_out(p.out, (y + 4 * (j >> 1) - p.out_row0) * p.ld_out + x + (j & 1) * 64, res, p.enc);
  }
}

static double py_round(double v) { return nearbyint(v); }  // half-to-even, like Python's round()

static int run_openness(const float* dem, void* out, const fsg_window* win, int negative, int n_dirs,
                        const OpenTable& tab, int D, double stretch_lo, double stretch_scale, const fsg_encode* enc,
                        void* stream) {
  if (!dem || !out || !win) return fail(FSG_E_INVALID, "fsg_openness: NULL argument");
  if (win->H_global < 1 || win->W < 1) return fail(FSG_E_INVALID, "fsg_openness: empty raster");
  if (win->out_row0 < 0 || win->out_rows < 0 || win->out_row0 + win->out_rows > win->H_global)
    return fail(FSG_E_INVALID, "fsg_openness: output rows outside the raster");
  int64_t lo = win->out_row0 - D; if (lo < 0) lo = 0;
  int64_t hi = win->out_row0 + win->out_rows + D; if (hi > win->H_global) hi = win->H_global;
  if (win->buf_row0 > lo || win->buf_row0 + win->buf_rows < hi)
    return fail(FSG_E_INVALID, "fsg_openness: buffer rows do not cover the %d-row halo", D);
  OpenParams p{};
  p.dem = dem; p.out = out; p.H = win->H_global; p.W = win->W; p.buf_row0 = win->buf_row0; p.buf_rows = win->buf_rows;
  p.out_row0 = win->out_row0; p.out_rows = win->out_rows; p.ld_in = win->ld_in; p.ld_out = win->ld_out;
  p.n_dirs = n_dirs; p.negative = negative ? 1 : 0;
  p.stretch = (!is_none(stretch_scale) && !is_none(stretch_lo) && stretch_scale > 1e-12) ? 1 : 0;
  p.stretch_lo = (float)stretch_lo; p.stretch_scale = (float)stretch_scale;
  p.stretch_rinv = p.stretch ? (float)(1.0 / (double)p.stretch_scale) : 0.f;
  p.enc = make_encode(enc);
  if (p.out_rows == 0) return FSG_OK;
  dim3 grid((unsigned)((p.W + 63) / 64), (unsigned)((p.out_rows + 3) / 4));
  int slot = prof_begin(PROF_OPENNESS, (cudaStream_t)stream);
  // interior fast path: sample offsets fit in int32,
  // the buffer holds every row a sample of an interior tile can touch
  bool fast = (double)p.ld_in * (double)(D + 1) < 2.0e9 && p.buf_row0 <= (p.out_row0 - D > 0 ? p.out_row0 - D : 0) &&
              p.buf_row0 + p.buf_rows >= (p.out_row0 + p.out_rows + D < p.H ? p.out_row0 + p.out_rows + D : p.H) &&
              !getenv("FSG_OPENNESS_GENERIC");
  OpenTable t2 = tab;
  {
    const int ns = t2.dir_start[n_dirs];
    for (int k = 0; k < ns; ++k) {
      t2.s[k].off = fast ? (int)((int64_t)t2.s[k].oy * p.ld_in + t2.s[k].ox) : 0;
      t2.s[k].rinv = 1.0f / t2.s[k].dist;
    }
  }
  for (int d = 0; d < n_dirs && fast; ++d)
    if (t2.dir_start[d + 1] - t2.dir_start[d] > OPI_NS) fast = false;
  if (fast && (double)p.ld_in * (double)(D + 1) * 4.0 >= 2.0e9) fast = false;   // byte offsets in int32
  if (fast) {
    OpenFixed fx;
    for (int d = 0; d < 64; ++d)
      for (int k = 0; k < OPI_NS; ++k) {
        OpenSlot& o = fx.s[d * OPI_NS + k];
        o.off_bytes = 0;
        o.rinv = nanf("");
        if (d < n_dirs && t2.dir_start[d] + k < t2.dir_start[d + 1]) {
          const OpenSample& sm = t2.s[t2.dir_start[d] + k];
          o.off_bytes = sm.off * 4;
          o.rinv = sm.rinv;
        }
      }
    dim3 g2((unsigned)((p.W + 127) / 128), (unsigned)((p.out_rows + OPI_ROWS - 1) / OPI_ROWS));
    if (p.negative) openness_interior_kernel<true><<<g2, 256, 0, (cudaStream_t)stream>>>(p, fx, D);
    else openness_interior_kernel<false><<<g2, 256, 0, (cudaStream_t)stream>>>(p, fx, D);
    FSG_LAUNCH_OK();
    // the border band (and whatever the tile grid leaves over): four rectangles of the generic kernel instead of a
    // launch over the whole raster whose CTAs mostly exit at once (4.2 M empty CTAs at 32768^2)
    const int64_t R0 = p.out_row0, R1 = p.out_row0 + p.out_rows;
    int64_t XL = ((int64_t)D + 127) / 128 * 128, XR = p.W - D >= 128 ? (p.W - D) / 128 * 128 : 0;
    if (XL > p.W) XL = p.W;
    if (XR < XL) XR = XL;
    // interior tile rows are anchored at out_row0: first row y with y >= D, last tile end <= min(H - D, R1)
    int64_t YT = R0 + ((D > R0 ? D - R0 : 0) + OPI_ROWS - 1) / OPI_ROWS * OPI_ROWS;
    const int64_t ylim = (p.H - D < R1 ? p.H - D : R1);
    int64_t YB = ylim > R0 ? R0 + (ylim - R0) / OPI_ROWS * OPI_ROWS : R0;
    if (YT > R1) YT = R1;
    if (YB < YT) YB = YT;
    auto rect = [&](int64_t x0, int64_t x1, int64_t y0, int64_t y1) {
      if (x1 <= x0 || y1 <= y0) return;
      OpenParams q = p;
      q.rx0 = x0; q.rx1 = x1; q.ry0 = y0; q.ry1 = y1;
      dim3 g((unsigned)((x1 - x0 + 63) / 64), (unsigned)((y1 - y0 + 3) / 4));
      openness_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(q, t2, D);
      count_launch();
    };
    rect(0, XL, R0, R1);
    rect(XR, p.W, R0, R1);
    rect(XL, XR, R0, YT);
    rect(XL, XR, YB, R1);
  } else {
    p.rx0 = 0; p.rx1 = p.W; p.ry0 = p.out_row0; p.ry1 = p.out_row0 + p.out_rows;
    openness_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, t2, -1);
  }
  prof_end(slot, (cudaStream_t)stream);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

}  // namespace fsg

extern "C" {

// Explicit ray-sample table (the Python host layer builds it with NumPy exactly as the reference
// does, so the integer offsets cannot drift with the libm in use).
int fsg_openness_samples(const float* dem, void* out, const fsg_window* win, int negative, int num_directions,
                         const int32_t* dir_start_host, const int32_t* ox_host, const int32_t* oy_host,
                         const float* dist_host, double stretch_lo, double stretch_scale, const fsg_encode* enc,
                         void* stream) {
  using namespace fsg;
  if (num_directions < 1 || num_directions > 64) return fail(FSG_E_INVALID, "fsg_openness: num_directions must be 1..64");
  if (!dir_start_host || !ox_host || !oy_host || !dist_host) return fail(FSG_E_INVALID, "fsg_openness: NULL sample table");
  int ns = dir_start_host[num_directions];
  if (ns < 0 || ns > OP_MAX_SAMPLES) return fail(FSG_E_INVALID, "fsg_openness: too many ray samples (%d)", ns);
  OpenTable tab{};
  int D = 0;
  for (int d = 0; d <= 64; ++d) tab.dir_start[d] = d <= num_directions ? dir_start_host[d] : ns;
  for (int k = 0; k < ns; ++k) {
    tab.s[k].ox = (short)ox_host[k]; tab.s[k].oy = (short)oy_host[k]; tab.s[k].dist = dist_host[k];
    int a = abs(ox_host[k]) > abs(oy_host[k]) ? abs(ox_host[k]) : abs(oy_host[k]);
    if (a > D) D = a;
  }
  return run_openness(dem, out, win, negative, num_directions, tab, D, stretch_lo, stretch_scale, enc, stream);
}

int fsg_openness(const float* dem, void* out, const fsg_window* win, int negative, int num_directions,
                 int max_distance, double pixel_size, double psx, double psy, double stretch_lo,
                 double stretch_scale, const fsg_encode* enc, void* stream) {
  using namespace fsg;
  if (num_directions < 1 || num_directions > 64) return fail(FSG_E_INVALID, "fsg_openness: num_directions must be 1..64");
  if (max_distance < 0 || max_distance > 32000) return fail(FSG_E_INVALID, "fsg_openness: max_distance out of range");
  // distances: np.unique((np.linspace(0.1, 1.0, 10) * max_distance).astype(int)), > 0  (:67-68)
  int dist[10], nd = 0;
  for (int i = 0; i < 10; ++i) {
    double step = (1.0 - 0.1) / 9.0;
    double f = (i == 9) ? 1.0 : (double)i * step + 0.1;
    int v = (int)(f * (double)max_distance);
    if (v <= 0) continue;
    if (nd == 0 || v != dist[nd - 1]) dist[nd++] = v;
  }
  int D = nd ? dist[nd - 1] : 0;
  double sx = is_none(psx) ? pixel_size : fabs(psx);
  double sy = is_none(psy) ? pixel_size : fabs(psy);
  if (sx < 1e-9) sx = pixel_size != 0.0 ? pixel_size : 1.0;
  if (sy < 1e-9) sy = pixel_size != 0.0 ? pixel_size : 1.0;
  OpenTable tab{};
  int ns = 0;
  const double two_pi = 2.0 * 3.14159265358979323846;
  for (int d = 0; d < num_directions; ++d) {
    tab.dir_start[d] = ns;
    double ang = (double)d * (two_pi / (double)num_directions);  // np.linspace(0, 2*pi, D, endpoint=False)
    double cx = cos(ang), cy = sin(ang);
    for (int k = 0; k < nd; ++k) {
      int ox = (int)py_round((double)dist[k] * cx);
      int oy = (int)py_round((double)dist[k] * cy);
      if (ox == 0 && oy == 0) continue;
      double pd = hypot((double)ox * sx, (double)oy * sy);
      if (pd < 1e-9) pd = 1e-9;
      tab.s[ns].ox = (short)ox; tab.s[ns].oy = (short)oy; tab.s[ns].dist = (float)pd;
      ++ns;
    }
  }
  for (int d = num_directions; d <= 64; ++d) tab.dir_start[d] = ns;
  return run_openness(dem, out, win, negative, num_directions, tab, D, stretch_lo, stretch_scale, enc, stream);
}

}  // extern "C"
