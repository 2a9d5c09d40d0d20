// topousm_fast: multiscale unsharp mask  out = sum_i w_i * (dem - mean_i(dem))
//
// Reference: compute_topousm_fast_efficient_block (algorithms/_impl_topousm_fast.py:49-100) with
//   _radius_to_downsample_factor (algorithms/_nan_utils.py:555-601)
//   _downsample_nan_aware        (:604-668)   -> pyramid_kernel (+ enclosed-void fill)
//   handle_nan_with_uniform/_gaussian (:18-47) on the decimated grids -> fsg_filters.cu
//   _upsample_to_shape           (:671-698)   -> align-corners bilinear taps inside the fused kernel
//   apply_global_normalization   (algorithms/_global_stats.py:123-153) + integer encoding -> epilogue
//
// Pipeline (all on the caller's stream, no host sync, no allocation):
//   1. pyramid_kernel : ONE read of the DEM produces every needed f x f valid-mean level
//                       (f32 sums in NumPy's reduction order -> bit-identical coarse grids).
//   2. per coarse term: box / sigma=1 Gaussian mean on its (small) level.
//   3. fused_kernel   : streams column strips of the DEM through a shared-memory row ring; for
//                       every full-resolution radius the vertical pass keeps exact f64 running
//                       window sums per column, the horizontal pass slides f64 sums along rows
//                       (results rounded to f32 exactly where scipy rounds: after each axis);
//                       coarse means are sampled with scipy's zoom arithmetic; terms are combined
//                       in list order in f32; normalise; NaN restore; optional u8/i16 encoding.
//      Algorithmic traffic: 4 B/px read + 4 B/px (f32) or 1 B/px (u8) written.
#include <stdlib.h>
#include <functional>
#include <map>
#include <mutex>
#include <queue>
#include <utility>
#include <vector>

#include "fsg_filters.cuh"

namespace fsg {

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
constexpr int MAX_TERMS = 16;
constexpr int MAX_FUSED_R = 41;  // ds==1 for r<=41 at pixel_size>=1 (and fewer for finer pixels)
constexpr int MAX_LEVELS = 4;    // decimation factors 2,4,8,16

enum TermKind { TERM_BOX_FUSED = 0, TERM_COARSE = 1, TERM_PLANE = 2 };

struct HostTerm {
  int kind;
  int radius;     // user radius
  int ds;         // decimation factor
  int size;       // box taps on its grid (0: sigma=1 gaussian)
  int level;      // index into levels[] for coarse terms
  size_t grid_off;  // workspace offset of the mean grid (coarse) / plane
};

struct HostLevel {
  int f;
  int64_t h, w;
  size_t off;  // workspace offset of the coarse grid
};

struct HostPlan {
  int n_terms;
  HostTerm terms[MAX_TERMS];
  int n_levels;
  HostLevel levels[MAX_LEVELS];
  size_t off_tv, off_tw;   // scratch planes (max over users)
  size_t off_taps;         // f64 taps scratch
  size_t off_flags;        // ints: [0..3] level has void cell, [4..7] level still has NaN after fill
  size_t off_need;         // bytes: (row, tile) maps of the enclosed-void fill (void_fill_need_bytes of the largest level)
  size_t off_v8flags, v8flag_bytes;   // NaN-block flags of the interior fast path (fused_kernel_v8)
  size_t total;
  int taps_cap;
  int fused_R;             // max fused radius
};

static int decimation_factor(double radius, double pixel_size) {
  // algorithms/_nan_utils.py:555-601 with algorithm_name="topousm_fast"
  double r = radius < 1.0 ? 1.0 : radius;
  double px = pixel_size != 0.0 ? pixel_size : 1.0;
  if (px < 1e-3) px = 1e-3;
  double res = 1.0 / px;
  if (res < 1.0) res = 1.0;
  double score = (r / 24.0) * 1.15 * 1.0 * pow(res, 0.35);
  if (score <= 1.0) return 1;
  int f = 1 << (int)floor(log2(score));
  if (f < 1) f = 1;
  if (f > 16) f = 16;
  return f;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int make_plan(int64_t H, int64_t W, const int32_t* radii, int n, double pixel_size, HostPlan* p) {
  if (n <= 0) return fail(FSG_E_INVALID, "At least one radius value is required");
  if (n > MAX_TERMS) return fail(FSG_E_UNSUPPORTED, "at most %d radii are supported, got %d", MAX_TERMS, n);
  if (H < 1 || W < 1) return fail(FSG_E_INVALID, "empty raster %lld x %lld", (long long)H, (long long)W);
  memset(p, 0, sizeof(*p));
  p->n_terms = n;
  size_t off = 0;
  size_t scratch = 0;  // elements needed in tv/tw
  int taps_cap = 8;
  int n_fused = 0;
  for (int i = 0; i < n; ++i) {
    HostTerm& t = p->terms[i];
    t.radius = radii[i];
    t.ds = decimation_factor((double)radii[i], pixel_size);
    if (t.ds > 1) {
      int rs = (int)nearbyint((double)radii[i] / (double)t.ds);  // Python round(): half to even
      if (rs < 1) rs = 1;
      t.kind = TERM_COARSE;
      t.size = rs <= 1 ? 0 : 2 * rs + 1;
      int lv = -1;
      for (int k = 0; k < p->n_levels; ++k)
        if (p->levels[k].f == t.ds) lv = k;
      if (lv < 0) {
        lv = p->n_levels++;
        p->levels[lv].f = t.ds;
        p->levels[lv].h = (H + t.ds - 1) / t.ds;
        p->levels[lv].w = (W + t.ds - 1) / t.ds;
      }
      t.level = lv;
    } else if (radii[i] <= 1) {
      t.kind = TERM_PLANE;
      t.size = 0;
    } else if (radii[i] <= MAX_FUSED_R && n_fused < 8) {
      t.kind = TERM_BOX_FUSED;
      t.size = 2 * radii[i] + 1;
      ++n_fused;
      if (radii[i] > p->fused_R) p->fused_R = radii[i];
    } else {
      t.kind = TERM_PLANE;  // general fallback (never hit by the reference's decimation rule)
      t.size = 2 * radii[i] + 1;
    }
  }
  for (int k = 0; k < p->n_levels; ++k) {
    HostLevel& l = p->levels[k];
    l.off = off;
    off = align_up(off + (size_t)l.h * l.w * 4, 256);
    if ((size_t)l.h * l.w > scratch) scratch = (size_t)l.h * l.w;
    double sigma = (double)(l.h < l.w ? l.h : l.w) / 64.0;
    if (sigma < 1.0) sigma = 1.0;
    int r = (int)(4.0 * sigma + 0.5);
    if (r + 1 > taps_cap) taps_cap = r + 1;
  }
  for (int i = 0; i < n; ++i) {
    HostTerm& t = p->terms[i];
    if (t.kind == TERM_COARSE) {
      const HostLevel& l = p->levels[t.level];
      t.grid_off = off;
      off = align_up(off + (size_t)l.h * l.w * 4, 256);
    } else if (t.kind == TERM_PLANE) {
      t.grid_off = off;
      off = align_up(off + (size_t)H * W * 4, 256);
      if ((size_t)H * W > scratch) scratch = (size_t)H * W;
    }
  }
  p->off_tv = off; off = align_up(off + scratch * 4, 256);
  p->off_tw = off; off = align_up(off + scratch * 4, 256);
  p->off_taps = off; off = align_up(off + (size_t)taps_cap * 8, 256);
  p->off_flags = off; off = align_up(off + 64, 256);
  {
    size_t nb = 0;
    for (int k = 0; k < p->n_levels; ++k) {
      const size_t b = void_fill_need_bytes(p->levels[k].h, p->levels[k].w);
      nb = b > nb ? b : nb;
    }
    p->off_need = off; off = align_up(off + nb, 256);
  }
  p->v8flag_bytes = (size_t)((H + 255) / 256 + 1) * (size_t)(W / 192 + 2) * sizeof(int);
  p->off_v8flags = off; off = align_up(off + p->v8flag_bytes, 256);
  p->taps_cap = taps_cap;
  p->total = off;
  return FSG_OK;
}

// ------------------------------------------------------------------------------------------
// 1. pyramid: f x f mean of finite members, NumPy reduction order
// ------------------------------------------------------------------------------------------
constexpr int PY_ROWS = 16;
constexpr int PY_COLS = 512;
constexpr int PY_STRIDE = PY_COLS + 4;  // 16-byte aligned rows; +4 words: 128-bit loads by rows (factor 16) are conflict free

struct PyramidParams {
  const float* dem;
  int64_t H, W, ld;
  int n_levels;
  int f[MAX_LEVELS];
  float* grid[MAX_LEVELS];
  int64_t gw[MAX_LEVELS];
  int64_t gh[MAX_LEVELS];
  int* flags;  // flags[k] = level k has a void (all-NaN) cell
};

__device__ __forceinline__ int py_idx(int row, int col) { return row * PY_STRIDE + col; }

// One cell: `sum(axis=(1,3), dtype=float32)` of NumPy = per cell row a pairwise row sum
// (n<8: sequential; n>=8: 8 lanes then ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))), rows added in order.
template <int F>
__device__ __forceinline__ void cell_reduce(const float* sm, int row0, int col0, float* tot, float* cnt) {
  float T = 0.f, Cn = 0.f;
#pragma unroll 1
  for (int i = 0; i < F; ++i) {
    float rs, rc;
    if (F < 8) {
      float v = sm[py_idx(row0 + i, col0)];
      bool ok = isfinite(v);
      rs = ok ? v : 0.f;
      rc = ok ? 1.f : 0.f;
#pragma unroll
      for (int j = 1; j < F; ++j) {
        v = sm[py_idx(row0 + i, col0 + j)];
        ok = isfinite(v);
        rs = rs + (ok ? v : 0.f);
        rc = rc + (ok ? 1.f : 0.f);
      }
    } else {
      float r[8], c[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = sm[py_idx(row0 + i, col0 + k)];
        bool ok = isfinite(v);
        r[k] = ok ? v : 0.f;
        c[k] = ok ? 1.f : 0.f;
      }
#pragma unroll
      for (int j = 8; j < F; j += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float v = sm[py_idx(row0 + i, col0 + j + k)];
          bool ok = isfinite(v);
          r[k] = r[k] + (ok ? v : 0.f);
          c[k] = c[k] + (ok ? 1.f : 0.f);
        }
      }
      rs = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
      rc = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
    }
    if (i == 0) { T = rs; Cn = rc; } else { T = T + rs; Cn = Cn + rc; }
  }
  *tot = T;
  *cnt = Cn;
}

// 128-bit variants for the two factors the reference's rule produces at 1 m pixels (4 and 16): same
// operation order as cell_reduce<4> / row_reduce<16>, a quarter of the shared-memory instructions.
__device__ __forceinline__ void fin4(const float4 q, float* v, float* c) {
  const float x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bool ok = isfinite(x[k]);
    v[k] = ok ? x[k] : 0.f;
    c[k] = ok ? 1.f : 0.f;
  }
}
__device__ __forceinline__ void cell4_vec(const float* sm, int row0, int col0, float* tot, float* cnt) {
  float T = 0.f, Cn = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[4], c[4];
    fin4(*reinterpret_cast<const float4*>(sm + py_idx(row0 + i, col0)), v, c);
    float rs = ((v[0] + v[1]) + v[2]) + v[3];   // n < 8: sequential
    float rc = ((c[0] + c[1]) + c[2]) + c[3];
    if (i == 0) { T = rs; Cn = rc; } else { T = T + rs; Cn = Cn + rc; }
  }
  *tot = T;
  *cnt = Cn;
}
__device__ __forceinline__ void row16_vec(const float* sm, int row, int col0, float* rs_out, float* rc_out) {
  float r[8], c[8], v[4], w[4];
  fin4(*reinterpret_cast<const float4*>(sm + py_idx(row, col0)), r, c);
  fin4(*reinterpret_cast<const float4*>(sm + py_idx(row, col0 + 4)), r + 4, c + 4);
  fin4(*reinterpret_cast<const float4*>(sm + py_idx(row, col0 + 8)), v, w);
#pragma unroll
  for (int k = 0; k < 4; ++k) { r[k] = r[k] + v[k]; c[k] = c[k] + w[k]; }
  fin4(*reinterpret_cast<const float4*>(sm + py_idx(row, col0 + 12)), v, w);
#pragma unroll
  for (int k = 0; k < 4; ++k) { r[4 + k] = r[4 + k] + v[k]; c[4 + k] = c[4 + k] + w[k]; }
  *rs_out = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  *rc_out = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
}

// One cell ROW of a factor >= 8 cell (the pairwise row sum of cell_reduce): used by the two-stage path of
// pyramid_kernel, where every (cell, row) pair is reduced by its own thread and the rows are then
// added in order by one thread per cell -- same operation order, 16x more parallelism for f = 16.
template <int F>
__device__ __forceinline__ void row_reduce(const float* sm, int row, int col0, float* rs_out, float* rc_out) {
  static_assert(F >= 8, "pairwise rows only");
  float r[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float v = sm[py_idx(row, col0 + k)];
    bool ok = isfinite(v);
    r[k] = ok ? v : 0.f;
    c[k] = ok ? 1.f : 0.f;
  }
#pragma unroll
  for (int j = 8; j < F; j += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = sm[py_idx(row, col0 + j + k)];
      bool ok = isfinite(v);
      r[k] = r[k] + (ok ? v : 0.f);
      c[k] = c[k] + (ok ? 1.f : 0.f);
    }
  }
  *rs_out = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  *rc_out = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
}

// Single-column coarse grid (W <= f): NumPy coalesces the two reduced axes into ONE contiguous run of
// f*f elements and reduces it with its pairwise routine (n<8 sequential; n<=128: 8 lanes + tree;
// n=256: two 128-element halves).
template <typename GET>
__device__ __forceinline__ float pairwise_block(GET get, int start, int n) {  // 8 <= n <= 128, n % 8 == 0
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = get(start + k);
  for (int i = 8; i < n; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = r[k] + get(start + i + k);
  }
  return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
}

__device__ void cell_reduce_contig(const float* sm, int f, int row0, int col0, float* tot, float* cnt) {
  auto val = [&](int e) { float v = sm[py_idx(row0 + e / f, col0 + e % f)]; return isfinite(v) ? v : 0.f; };
  auto one = [&](int e) { float v = sm[py_idx(row0 + e / f, col0 + e % f)]; return isfinite(v) ? 1.f : 0.f; };
  const int n = f * f;
  if (n < 8) {
    float a = val(0), c = one(0);
    for (int e = 1; e < n; ++e) { a = a + val(e); c = c + one(e); }
    *tot = a; *cnt = c;
  } else if (n <= 128) {
    *tot = pairwise_block(val, 0, n); *cnt = pairwise_block(one, 0, n);
  } else {
    *tot = pairwise_block(val, 0, 128) + pairwise_block(val, 128, 128);
    *cnt = pairwise_block(one, 0, 128) + pairwise_block(one, 128, 128);
  }
}

__global__ void __launch_bounds__(256) pyramid_kernel(PyramidParams p) {
  __shared__ __align__(16) float sm[PY_ROWS * PY_STRIDE];
  const int64_t x0 = (int64_t)blockIdx.x * PY_COLS;
  const int64_t y0 = (int64_t)blockIdx.y * PY_ROWS;
  const int tid = threadIdx.x;
  const float qnan = nanf("");
  // coalesced load, ragged edge -> NaN (the reference NaN-pads before reshaping)
  const bool tile_fast = y0 + PY_ROWS <= p.H && x0 + PY_COLS <= p.W && (p.ld % 4 == 0) &&
                         ((((uintptr_t)p.dem) & 15) == 0);
  if (tile_fast) {
    // whole, aligned tile: all eight 16-byte loads of a thread are issued before the first store
    constexpr int NL = PY_ROWS * (PY_COLS / 4) / 256;
    float4 q[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const int i = tid + k * 256;
      const int r = i / (PY_COLS / 4), c4 = (i - r * (PY_COLS / 4)) * 4;
      q[k] = __ldg(reinterpret_cast<const float4*>(p.dem + (y0 + r) * p.ld + x0 + c4));
    }
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const int i = tid + k * 256;
      const int r = i / (PY_COLS / 4), c4 = (i - r * (PY_COLS / 4)) * 4;
      *reinterpret_cast<float4*>(sm + py_idx(r, c4)) = q[k];
    }
  } else {
    for (int i = tid; i < PY_ROWS * (PY_COLS / 4); i += 256) {
      int r = i / (PY_COLS / 4), c4 = (i - r * (PY_COLS / 4)) * 4;
      int64_t gy = y0 + r, gx = x0 + c4;
      float v[4] = {qnan, qnan, qnan, qnan};
      if (gy < p.H) {
        const float* src = p.dem + gy * p.ld + gx;
        if (gx + 3 < p.W && ((((uintptr_t)src) & 15) == 0)) {
          float4 q = *reinterpret_cast<const float4*>(src);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (gx + k < p.W) v[k] = src[k];
        }
      }
      *reinterpret_cast<float4*>(sm + py_idx(r, c4)) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  __syncthreads();
  __shared__ float RS[PY_ROWS * PY_COLS / 8], RC[PY_ROWS * PY_COLS / 8];
  for (int lv = 0; lv < p.n_levels; ++lv) {
    const int f = p.f[lv];
    const int cells_x = PY_COLS / f, cells_y = PY_ROWS / f;
    if (f >= 8 && p.gw[lv] != 1) {
      // stage 1: one (cell, row) pair per thread; stage 2: rows added in order, one cell per thread
      if (f == 16) {
        // lanes = the 16 rows of two adjacent cells: rows are 516 words apart -> conflict-free LDS.128
        for (int cx = tid / PY_ROWS; cx < cells_x; cx += 256 / PY_ROWS) {
          const int row = tid % PY_ROWS;
          float rs, rc;
          row16_vec(sm, row, cx * 16, &rs, &rc);
          RS[row * cells_x + cx] = rs;
          RC[row * cells_x + cx] = rc;
        }
      } else {
        for (int it = tid; it < cells_x * PY_ROWS; it += 256) {
          int row = it / cells_x, cx = it - row * cells_x;
          float rs, rc;
          row_reduce<8>(sm, row, cx * 8, &rs, &rc);
          RS[it] = rs;
          RC[it] = rc;
        }
      }
      __syncthreads();
      for (int i = tid; i < cells_x * cells_y; i += 256) {
        int cy = i / cells_x, cx = i - cy * cells_x;
        int64_t oy = y0 / f + cy, ox = x0 / f + cx;
        if (oy >= p.gh[lv] || ox >= p.gw[lv]) continue;
        float tot = RS[(cy * f) * cells_x + cx], cnt = RC[(cy * f) * cells_x + cx];
        for (int r = 1; r < f; ++r) {
          tot = tot + RS[(cy * f + r) * cells_x + cx];
          cnt = cnt + RC[(cy * f + r) * cells_x + cx];
        }
        float out;
        if (cnt > 0.f) out = tot / fmaxf(cnt, 1.f);
        else { out = qnan; p.flags[lv] = 1; }
        p.grid[lv][oy * p.gw[lv] + ox] = out;
      }
      __syncthreads();
      continue;
    }
    for (int i = tid; i < cells_x * cells_y; i += 256) {
      int cy = i / cells_x, cx = i - cy * cells_x;
      int64_t oy = y0 / f + cy, ox = x0 / f + cx;
      if (oy >= p.gh[lv] || ox >= p.gw[lv]) continue;
      float tot, cnt;
      if (p.gw[lv] == 1) cell_reduce_contig(sm, f, cy * f, cx * f, &tot, &cnt);
      else if (f == 2) cell_reduce<2>(sm, cy * 2, cx * 2, &tot, &cnt);
      else if (f == 4) cell4_vec(sm, cy * 4, cx * 4, &tot, &cnt);
      else if (f == 8) cell_reduce<8>(sm, cy * 8, cx * 8, &tot, &cnt);
      else cell_reduce<16>(sm, cy * 16, cx * 16, &tot, &cnt);
      float out;
      if (cnt > 0.f) out = tot / fmaxf(cnt, 1.f);
      else { out = qnan; p.flags[lv] = 1; }
      p.grid[lv][oy * p.gw[lv] + ox] = out;
    }
  }
}

// flags[4+lv] = grid still holds NaN (set by void fill when it leaves a cell; or copy of flags[lv]
// when... see host code) -- helper to zero the flag block
__global__ void zero_flags_kernel(int* flags) {
  if (threadIdx.x < 16) flags[threadIdx.x] = 0;
}

// ------------------------------------------------------------------------------------------
// 3. fused full-resolution kernel
// ------------------------------------------------------------------------------------------
constexpr int FK_THREADS = 352;             // 11 warps
constexpr int FK_NB = 32;                   // rows per batch
constexpr int FK_TW = 264;                  // output columns per strip (11 segments of 24)
constexpr int FK_NSEG = FK_THREADS / FK_NB; // 11 column segments in the horizontal phase
constexpr int FK_SEG = (FK_TW + FK_NSEG - 1) / FK_NSEG;  // 24
constexpr int FK_MAXF = 8;

struct DevTerm {
  int kind;
  int r;             // fused radius
  float weight;
  const float* grid; // coarse mean grid / plane
  int64_t gh, gw;    // coarse dims
  double rscale, cscale;  // (gh-1)/(H-1), (gw-1)/(W-1)
  int lvl;                // coarse terms: index of the pyramid level (column-fraction table)
  int64_t grow0;          // global row of grid row 0 (row-band shards hold a window of the grid)
};

struct FusedParams {
  const float* dem;
  void* out;
  int64_t H, W, ld_in, ld_out;   // H, W: GLOBAL raster size (edge rules, zoom mapping)
  int64_t dem_row0, dem_rows;    // global row of dem[0] and rows held (row-band shards; 0, H for a whole raster)
  int64_t out_row0, out_rows;    // global rows to produce; out[0] is row out_row0
  int n_terms;
  DevTerm terms[MAX_TERMS];
  int R;           // halo (max fused radius, 0 if none)
  int band_rows;   // rows per CTA band (multiple of FK_NB)
  int ring_rows;   // rows held by the shared-memory ring (fast kernel)
  int n_lvls;      // pyramid levels referenced by coarse terms
  double lvl_cscale[MAX_LEVELS];
  int lvl_gw[MAX_LEVELS];
  float norm_rinv; // f32(1/norm_scale)
  int norm_mode;   // 0 none, 1 divide by norm_scale, 2 zeros, 3 scale read from norm_scale_dev (NaN: none, <= 0: zeros)
  float norm_scale;
  const float* norm_scale_dev;
  EncodeDev enc;
  int bulk_ok;     // v6: DEM rows are 16-byte aligned (base pointer and row stride) -> bulk async copies
  int out_vec_ok;  // v7: output rows allow 16-byte (f32) / 4-byte (u8) vector stores
  int strip0;      // first column strip to compute (region of interest; 0 for the whole width)
  int64_t roi_col0, roi_cols;   // host side: requested output columns (0, W for the whole width)
  // v6 rectangle launches (raster borders and NaN blocks next to a v8 launch): strips start at col0, output
  // columns stop at col_end (0: W); with tile_flags, CTA (bx, by) owns the tile_w x tile_rows tile (bx, by) of the
  // grid anchored at (col0, tile_row0) and runs only if its flag is set
  int col0, col_end;
  const int* tile_flags;
  int tile_w, tile_rows;
  int64_t tile_row0, tile_row1;
  // v8 (interior fast path): strips of V8_TW columns from v8_col0, rows [v8_row0, v8_row1), NaN block flags
  int v8_col0;
  int v8_xlast;   // first column of the last strip (the last strip may overlap its neighbour to end at the region's edge)
  int64_t v8_row0, v8_row1;
  int* v8_flags;
};

// normalisation resolved on the device: the p99 scale may still be on its way when the launch is enqueued
struct NormDev {
  int mode;
  float sc, rinv;
};
__device__ __forceinline__ NormDev resolve_norm(const FusedParams& p) {
  NormDev n{p.norm_mode, p.norm_scale, p.norm_rinv};
  if (p.norm_mode == 3) {
    const float s = *p.norm_scale_dev;
    if (s != s) n.mode = 0;
    else if (s > 0.f) { n.mode = 1; n.sc = s; n.rinv = (float)(1.0 / (double)s); }
    else n.mode = 2;
  }
  return n;
}

__global__ void __launch_bounds__(FK_THREADS, 1) fused_kernel(FusedParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const NormDev nd = resolve_norm(p);
  const int R = p.R;
  const int SW = FK_TW + 2 * R;          // strip width incl. halo
  const int SWp = SW | 1;                // odd row stride (bank-conflict-free row-parallel access)
  const int NRING = FK_NB + 2 * R + 1;   // rows resident in the ring
  float* ring = reinterpret_cast<float*>(smraw);                 // NRING x SWp
  float* vplane = ring + (size_t)NRING * SWp;                    // FK_NB x SWp  (also output staging)
  unsigned char* cplane = reinterpret_cast<unsigned char*>(vplane + (size_t)FK_NB * SWp);  // FK_NB x SWp

  const int tid = threadIdx.x;
  const int64_t x0 = ((int64_t)blockIdx.x + p.strip0) * FK_TW;       // first output column of the strip
  const int64_t cs0 = x0 - R;                            // global column of strip slot 0
  const int64_t out_end = p.out_row0 + p.out_rows;
  const int64_t yb0 = p.out_row0 + (int64_t)blockIdx.y * p.band_rows;
  const int64_t yb1 = (yb0 + p.band_rows < out_end) ? yb0 + p.band_rows : out_end;
  const int64_t H = p.H, W = p.W;

  // column owned in the vertical phase
  const int vc = tid;                                    // strip slot
  const int64_t vgx = cs0 + vc;
  const bool vcol_ok = vc < SW && vgx >= 0 && vgx < W;

  // per-thread running window state for the fused radii
  double rs[FK_MAXF];
  int rc[FK_MAXF];
  int fr[FK_MAXF];
  int nf = 0;
  for (int t = 0; t < p.n_terms; ++t)
    if (p.terms[t].kind == TERM_BOX_FUSED) fr[nf++] = p.terms[t].r;

  // horizontal-phase mapping: lane <-> row, warp-group <-> column segment
  const int hi = tid % FK_NB;
  const int hg = tid / FK_NB;
  const int hj0 = hg * FK_SEG;
  int hjn = (hj0 + FK_SEG <= FK_TW) ? FK_SEG : (FK_TW - hj0 > 0 ? FK_TW - hj0 : 0);
  if (x0 + hj0 + hjn > p.W) hjn = (p.W - x0 - hj0 > 0) ? (int)(p.W - x0 - hj0) : 0;  // never slide past the raster

  unsigned nan_hist = 0;  // bit b: batch loaded b steps ago contained NaN (uniform)
  int64_t loaded_hi = -1; // highest actual row resident (exclusive upper bound - 1)

  auto ring_at = [&](int64_t row, int slot) -> float& { return ring[(size_t)(row % NRING) * SWp + slot]; };

  for (int64_t y = yb0; y < yb1; y += FK_NB) {
    // ---- (1) make rows [y-R, y+NB+R] (clipped to the raster) resident ----
    int64_t need_lo = y - R < 0 ? 0 : y - R;
    int64_t need_hi = y + FK_NB + R >= H ? H - 1 : y + FK_NB + R;
    if (need_hi > p.dem_row0 + p.dem_rows - 1) need_hi = p.dem_row0 + p.dem_rows - 1;
    int64_t from = (y == yb0) ? need_lo : loaded_hi + 1;
    __syncthreads();  // everyone is done with the rows about to be overwritten
    int my_nan = 0;
    for (int64_t row = from; row <= need_hi; ++row) {
      for (int c = tid; c < SW; c += FK_THREADS) {
        int64_t gx = cs0 + c;
        if (gx >= 0 && gx < W) {
          float v = p.dem[(row - p.dem_row0) * p.ld_in + gx];
          my_nan |= (v != v);
          ring_at(row, c) = v;
        }
      }
    }
    loaded_hi = need_hi;
    int any = __syncthreads_or(my_nan);
    nan_hist = (y == yb0) ? (any ? 0xffffffffu : 0u) : ((nan_hist << 1) | (any ? 1u : 0u));
    // ring spans NB+2R+1 rows <= 4 batches (+1 for safety)
    const bool nanmode = (nan_hist & 0x3fu) != 0;

    // ---- (2) (re)initialise the running sums at the first batch of the band ----
    if (y == yb0 && vcol_ok) {
      for (int k = 0; k < nf; ++k) {
        double s = 0.0;
        int c = 0;
        for (int d = -fr[k]; d <= fr[k]; ++d) {
          float v = ring_at(reflect_index(y + d, H), vc);
          bool ok = v == v;
          s += ok ? (double)v : 0.0;
          c += ok;
        }
        rs[k] = s;
        rc[k] = c;
      }
    }

    // ---- (3) terms in list order ----
    float acc[FK_SEG];
    int fk = 0;
    bool first = true;
    const int64_t orow = y + hi;  // output row of this thread in the horizontal phase
    const bool hrow_ok = orow < yb1;
    for (int t = 0; t < p.n_terms; ++t) {
      const DevTerm& T = p.terms[t];
      if (T.kind == TERM_BOX_FUSED) {
        const int r = T.r;
        const double n = (double)(2 * r + 1), inv = 1.0 / n;
        const float nf32 = (float)(2 * r + 1);
        // vertical phase: thread = column
        if (vcol_ok && vc >= R - r && vc < R + FK_TW + r) {
          double s = rs[fk];
          int c = rc[fk];
          for (int i = 0; i < FK_NB; ++i) {
            int64_t row = y + i;
            if (row >= yb1) break;
            vplane[i * SWp + vc] = (float)div_by_count(s, n, inv);
            float vin = ring_at(reflect_index(row + r + 1, H), vc);
            float vout = ring_at(reflect_index(row - r, H), vc);
            if (nanmode) {
              cplane[i * SWp + vc] = (unsigned char)c;
              bool oin = vin == vin, oout = vout == vout;
              s += (oin ? (double)vin : 0.0) - (oout ? (double)vout : 0.0);
              c += (int)oin - (int)oout;
            } else {
              s += (double)vin - (double)vout;
            }
          }
          rs[fk] = s;
          rc[fk] = c;
        }
        __syncthreads();
        // horizontal phase: thread = (row, column segment)
        if (hrow_ok && hjn > 0) {
          const float* vrow = vplane + hi * SWp;
          const unsigned char* crow = cplane + hi * SWp;
          const float* xrow = ring + (size_t)(orow % NRING) * SWp;
          auto slot = [&](int64_t j) -> int {  // strip slot of global column x0+j (reflected)
            int64_t gx = x0 + j;
            if (gx < 0 || gx >= W) gx = reflect_index(gx, W);
            return (int)(gx - cs0);
          };
          double sv = 0.0, sw = 0.0;
          for (int d = -r; d <= r; ++d) {
            int sidx = slot(hj0 + d);
            sv += (double)vrow[sidx];
            if (nanmode) sw += (double)((float)crow[sidx] / nf32);
          }
#pragma unroll
          for (int jj = 0; jj < FK_SEG; ++jj) {
            if (jj < hjn) {
              int64_t j = hj0 + jj;
              float mean = (float)div_by_count(sv, n, inv);
              if (nanmode) {
                float den = (float)div_by_count(sw, n, inv);
                mean = den > 0.f ? mean / den : 0.f;
              }
              float x = xrow[R + j];
              float term = T.weight * (x - mean);
              acc[jj] = first ? term : acc[jj] + term;
              if (jj + 1 < hjn) {
                int sin_ = slot(j + r + 1), sout = slot(j - r);
                sv += (double)vrow[sin_] - (double)vrow[sout];
                if (nanmode) sw += (double)((float)crow[sin_] / nf32) - (double)((float)crow[sout] / nf32);
              }
            }
          }
        }
        __syncthreads();
        ++fk;
      } else if (T.kind == TERM_COARSE) {
        if (hrow_ok && hjn > 0) {
          const float* xrow = ring + (size_t)(orow % NRING) * SWp;
          // scipy zoom(order=1): coordinate = index * (n_in-1)/(n_out-1); weights [1-t, t]; 'nearest' edge
          double ri = (double)orow * T.rscale;
          int64_t r0 = (int64_t)floor(ri);
          if (r0 > T.gh - 1) r0 = T.gh - 1;
          double tr = ri - (double)r0;
          int64_t r1 = r0 + 1 < T.gh ? r0 + 1 : T.gh - 1;
          double wr0 = 1.0 - tr, wr1 = tr;
          const float* g0 = T.grid + (r0 - T.grow0) * T.gw;
          const float* g1 = T.grid + (r1 - T.grow0) * T.gw;
          double ci = (double)(x0 + hj0) * T.cscale;
          int64_t c0 = (int64_t)floor(ci);
          if (c0 > T.gw - 1) c0 = T.gw - 1;
          double next_int = (double)(c0 + 1);
          int64_t c1 = c0 + 1 < T.gw ? c0 + 1 : T.gw - 1;
          double p00 = (double)g0[c0] * wr0, p01 = (double)g0[c1] * wr0;
          double p10 = (double)g1[c0] * wr1, p11 = (double)g1[c1] * wr1;
#pragma unroll
          for (int jj = 0; jj < FK_SEG; ++jj) {
            if (jj < hjn) {
              int64_t gx = x0 + hj0 + jj;
              if (gx < W) {
                ci = (double)gx * T.cscale;
                while (ci >= next_int && c0 < T.gw - 1) {
                  ++c0;
                  next_int += 1.0;
                  c1 = c0 + 1 < T.gw ? c0 + 1 : T.gw - 1;
                  p00 = (double)g0[c0] * wr0; p01 = (double)g0[c1] * wr0;
                  p10 = (double)g1[c0] * wr1; p11 = (double)g1[c1] * wr1;
                }
                double tc = ci - (double)c0;
                double wc0 = 1.0 - tc;
                double v = p00 * wc0;
                v += p01 * tc;
                v += p10 * wc0;
                v += p11 * tc;
                float mean = (float)v;
                float x = xrow[R + hj0 + jj];
                float term = T.weight * (x - mean);
                acc[jj] = first ? term : acc[jj] + term;
              }
            }
          }
        }
      } else {  // TERM_PLANE: precomputed full-resolution mean
        if (hrow_ok && hjn > 0) {
          const float* xrow = ring + (size_t)(orow % NRING) * SWp;
          const float* prow = T.grid + (orow - T.grow0) * W;
#pragma unroll
          for (int jj = 0; jj < FK_SEG; ++jj) {
            if (jj < hjn) {
              int64_t gx = x0 + hj0 + jj;
              if (gx < W) {
                float x = xrow[R + hj0 + jj];
                float term = T.weight * (x - prow[gx]);
                acc[jj] = first ? term : acc[jj] + term;
              }
            }
          }
        }
      }
      first = false;
    }

    // ---- (4) epilogue: normalise, stage, coalesced store ----
    float* stage = vplane;  // FK_NB x (FK_TW+1)
    if (hrow_ok && hjn > 0) {
#pragma unroll
      for (int jj = 0; jj < FK_SEG; ++jj) {
        if (jj < hjn) {
          float v = acc[jj];
          if (nd.mode == 1) v = v / nd.sc;
          else if (nd.mode == 2) v = (v != v) ? v : 0.f;
          stage[hi * (FK_TW + 1) + hj0 + jj] = v;
        }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < FK_NB * FK_TW; idx += FK_THREADS) {
      int rr = idx / FK_TW, c = idx - rr * FK_TW;
      int64_t gy = y + rr, gx = x0 + c;
      if (gy < yb1 && gx < W) store_out(p.out, (gy - p.out_row0) * p.ld_out + gx, stage[rr * (FK_TW + 1) + c], p.enc);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 3b. fast variant of the fused kernel (rasters with H, W >= R + 2: a single mirror reflection is
//     enough at the edges).  Same arithmetic as fused_kernel.
//
//   ring   : the NB+2R+1 DEM rows one batch needs, filled with cp.async (LDGSTS).  The rows of batch
//            b+1 are issued as soon as the last vertical pass of batch b has consumed the rows they
//            replace, so the copy overlaps the horizontal passes, the coarse terms and the stores.
//   vphase : thread = column; exact f64 running window sum down the rows (state in registers across
//            batches).  The mean is rounded to the f32 grid WITHOUT leaving f64 (add/subtract
//            2^(e+29): RN-even at 24 bits == scipy's f32 store after axis 0) and kept as f64 in
//            shared memory, so the horizontal pass needs no f32->f64 conversions (conversions run at
//            1/4 rate on B200 and were the top stall reason).
//   hphase : thread = (row, SEG-column segment); lanes of a warp = different rows (odd row stride ->
//            conflict-free); f64 sliding sum along the row; f32 rounding (scipy's store after axis 1);
//            weighted difference accumulated in f32 in list order.
//   coarse : align-corners bilinear tap of the decimated mean (one f64 FMA per pixel; column
//            fractions tabulated once per CTA).
//   output : staged through shared memory -> coalesced stores (f32 / i16 / u8 encoding fused).
//   NaN    : batches whose ring holds a NaN take the NaN-aware form (value and validity sums, f32
//            planes); results are bit-identical to the dense form where no NaN is in reach.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect1(int64_t i, int64_t n) {  // one mirror, valid for -n <= i < 2n
  return (int)(i < 0 ? -1 - i : (i >= n ? 2 * n - 1 - i : i));
}

template <bool EDGE>
__device__ __forceinline__ int col_slot(int j, int x0, int cs0, int W) {  // strip slot of output column j (+-r)
  if (!EDGE) return j - (cs0 - x0);
  int64_t gx = (int64_t)x0 + j;
  if (gx < 0 || gx >= W) gx = reflect1(gx, W);
  return (int)(gx - cs0);
}

// round an f64 to the nearest f32-representable value (ties to even), staying in f64
__device__ __forceinline__ double round_to_f32_grid(double q) {
  int hi = __double2hiint(q);
  double m = __hiloint2double((hi & 0xfff00000) + (29 << 20), 0);
  return (q + m) - m;
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// generic horizontal pass (edge strips, partial segments, NaN mode).  PLANE = double (dense) or float.
template <bool EDGE, int SEG, typename PLANE>
__device__ __forceinline__ void hphase_generic(const PLANE* __restrict__ vrow, const unsigned char* __restrict__ crow,
                                               const float* xr, int r, int hj0, int hjn, int x0, int cs0, int W,
                                               bool nanmode, float weight, float* acc) {
  const double inv = 1.0 / (double)(2 * r + 1);
  const double n = (double)(2 * r + 1);
  const float nf32 = (float)(2 * r + 1);
  double sv = 0.0, sw = 0.0;
  for (int d = -r; d <= r; ++d) {
    int sidx = col_slot<EDGE>(hj0 + d, x0, cs0, W);
    sv += (double)vrow[sidx];
    if (nanmode) sw += (double)((float)crow[sidx] / nf32);
  }
#pragma unroll
  for (int jj = 0; jj < SEG; ++jj) {
    if (jj < hjn) {
      const int j = hj0 + jj;
      float mean;
      if (nanmode) {
        mean = (float)div_by_count(sv, n, inv);
        float den = (float)div_by_count(sw, n, inv);
        mean = den > 0.f ? mean / den : 0.f;
      } else {
        mean = (float)(sv * inv);
      }
      acc[jj] = acc[jj] + weight * (xr[jj] - mean);
      if (jj + 1 < hjn) {
        int sin_ = col_slot<EDGE>(j + r + 1, x0, cs0, W), sout = col_slot<EDGE>(j - r, x0, cs0, W);
        sv += (double)vrow[sin_] - (double)vrow[sout];
        if (nanmode) sw += (double)((float)crow[sin_] / nf32) - (double)((float)crow[sout] / nf32);
      }
    }
  }
}

// bilinear taps of one coarse term along a thread's segment: A0/A1 are the row-interpolated coarse
// values left/right of the current pixel; `adv` marks the pixels where the coarse column advances
template <int SEG, bool FULL>
__device__ __forceinline__ void coarse_taps(const float* __restrict__ g0, const float* __restrict__ g1, double wr0,
                                            double tr, int c0, int c1, int gwm1, unsigned adv,
                                            const double* __restrict__ tcp, double A0, double A1, float wgt,
                                            const float* xr, float* acc, int hjn) {
  double dA = A1 - A0;
#pragma unroll
  for (int jj = 0; jj < SEG; ++jj) {
    if (FULL || jj < hjn) {
      if ((adv >> jj) & 1u) {
        ++c0;
        c1 = c0 + 1 < gwm1 ? c0 + 1 : gwm1;
        A0 = A1;
        A1 = (double)__ldg(g0 + c1) * wr0 + (double)__ldg(g1 + c1) * tr;
        dA = A1 - A0;
      }
      float mean = (float)fma(tcp[jj], dA, A0);
      acc[jj] = acc[jj] + wgt * (xr[jj] - mean);
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(FK_THREADS, 1) fused_kernel_fast(FusedParams p) {
  constexpr int NSEG = FK_THREADS / NB;
  constexpr int SEG = FK_TW / NSEG;
  static_assert(SEG * NSEG == FK_TW, "segments must tile the strip");
  extern __shared__ __align__(16) unsigned char smraw[];
  const NormDev nd = resolve_norm(p);
  const int R = p.R;
  const int SW = FK_TW + 2 * R;
  const int SWp = SW | 1;
  const int NRING = NB + 2 * R + 1;
  // layout (bytes from the 16-byte aligned base): ring f32 | plane (f64, or f32 + u8 counts) | tables
  const size_t off_plane = ((size_t)NRING * SWp * 4 + 15) / 16 * 16;
  const size_t off_slot = off_plane + (size_t)NB * SWp * 8;
  const size_t off_tc = (off_slot + (size_t)(NRING + 1) * 4 + 15) / 16 * 16;
  float* ring = reinterpret_cast<float*>(smraw);
  double* plane64 = reinterpret_cast<double*>(smraw + off_plane);
  float* plane32 = reinterpret_cast<float*>(smraw + off_plane);
  unsigned char* cplane = smraw + off_plane + (size_t)NB * SWp * 4;
  int* slot_tab = reinterpret_cast<int*>(smraw + off_slot);
  double* tctab = reinterpret_cast<double*>(smraw + off_tc);

  const int tid = threadIdx.x;
  const int W = (int)p.W;
  const int64_t H = p.H;
  const int x0 = ((int)blockIdx.x + p.strip0) * FK_TW;
  const int cs0 = x0 - R;
  const int64_t out_end = p.out_row0 + p.out_rows;
  const int64_t yb0 = p.out_row0 + (int64_t)blockIdx.y * p.band_rows;
  const int64_t yb1 = (yb0 + p.band_rows < out_end) ? yb0 + p.band_rows : out_end;
  const bool edge_strip = (cs0 < 0) || (cs0 + SW > W);

  const int vc = tid;
  const int vgx = cs0 + vc;
  const bool vcol_ok = vc < SW && vgx >= 0 && vgx < W;

  double rs[FK_MAXF];
  int rc[FK_MAXF];
#pragma unroll
  for (int k = 0; k < FK_MAXF; ++k) { rs[k] = 0.0; rc[k] = 0; }
  int last_fused = -1;
  for (int t = 0; t < p.n_terms; ++t)
    if (p.terms[t].kind == TERM_BOX_FUSED) last_fused = t;

  const int hi = tid % NB;
  const int hg = tid / NB;
  const int hj0 = hg * SEG;
  int hjn = SEG;
  if (x0 + hj0 + hjn > W) hjn = (W - x0 - hj0 > 0) ? (W - x0 - hj0) : 0;

  // Column coordinates of the coarse levels are the same for every row of the strip: the fraction
  // tc(j) = j*cscale - floor(j*cscale) goes to shared memory once per CTA, and each thread keeps the
  // first coarse column of its segment plus a bit mask of the pixels where that column advances.
  int lv_c0[MAX_LEVELS];
  unsigned lv_adv[MAX_LEVELS];
#pragma unroll
  for (int l = 0; l < MAX_LEVELS; ++l) {
    lv_c0[l] = 0;
    lv_adv[l] = 0u;
    if (l < p.n_lvls) {
      const double cs = p.lvl_cscale[l];
      const int gwm1 = p.lvl_gw[l] - 1;
      for (int j = tid; j < FK_TW; j += FK_THREADS) {
        double ci = (double)(x0 + j) * cs;
        double fl = floor(ci);
        if (fl > (double)gwm1) fl = (double)gwm1;
        tctab[l * FK_TW + j] = ci - fl;
      }
      int prev = 0;
#pragma unroll
      for (int jj = 0; jj < SEG; ++jj) {
        int c = (int)floor((double)(x0 + hj0 + jj) * cs);
        if (c > gwm1) c = gwm1;
        if (jj == 0) lv_c0[l] = c;
        else if (c > prev) lv_adv[l] |= 1u << jj;
        prev = c;
      }
    }
  }

  // rows live at ring slot (row - row_org) mod NRING
  const int64_t row_org = yb0 - R < 0 ? 0 : yb0 - R;
  const int64_t dem_last = p.dem_row0 + p.dem_rows - 1;   // last row the caller's buffer holds
  auto need_hi_of = [&](int64_t yy) {
    int64_t v = yy + NB + R >= H ? H - 1 : yy + NB + R;
    return v > dem_last ? dem_last : v;
  };
  const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring) + 4u * (unsigned)vc;
  const unsigned row_bytes = 4u * (unsigned)SWp;
  auto issue_rows = [&](int64_t from, int64_t to) {   // cp.async rows [from, to] of this thread's column
    if (vcol_ok && from <= to) {
      int sl = (int)((from - row_org) % NRING);
      int n = (int)(to - from + 1);
      const float* src = p.dem + (from - p.dem_row0) * p.ld_in + vgx;
      while (n > 0) {
        int run = NRING - sl < n ? NRING - sl : n;   // rows until the ring wraps
        unsigned sa = ring_sa + (unsigned)sl * row_bytes;
#pragma unroll 4
        for (int k = 0; k < run; ++k) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src) : "memory");
          sa += row_bytes;
          src += p.ld_in;
        }
        n -= run;
        sl += run;
        if (sl >= NRING) sl = 0;
      }
    }
  };

  unsigned nan_hist = 0;
  int base = (int)(yb0 - row_org);  // ring slot of row y
  issue_rows(row_org, need_hi_of(yb0));
  cp_async_commit();

  for (int64_t y = yb0; y < yb1; y += NB) {
    const int64_t need_hi = need_hi_of(y);
    const int64_t prev_hi = (y == yb0) ? row_org - 1 : need_hi_of(y - NB);
    cp_async_wait_all();
    // slot table: entry (d + R + 1) = SWp * ring slot of (mirrored) row y + d, d in [-R-1, NB+R]
    for (int e = tid; e <= NRING; e += FK_THREADS) {
      int d = e - R - 1;
      int dd = reflect1(y + d, H) - (int)y;
      int sl = (base + dd) % NRING;
      if (sl < 0) sl += NRING;
      slot_tab[e] = sl * SWp;
    }
    __syncthreads();   // this batch's rows (every thread's cp.async) and the slot table are visible
    int my_nan = 0;
    if (vcol_ok) {
      int sl = (int)((prev_hi + 1 - row_org) % NRING);
      int n = (int)(need_hi - prev_hi);
      float probe = 0.f;   // NaN is sticky under addition of 0*x: one FFMA-free test at the end
      while (n > 0) {
        int run = NRING - sl < n ? NRING - sl : n;
        const float* pp = ring + (size_t)sl * SWp + vc;
#pragma unroll 4
        for (int k = 0; k < run; ++k) { probe += pp[0] * 0.f; pp += SWp; }
        n -= run;
        sl += run;
        if (sl >= NRING) sl = 0;
      }
      my_nan = (probe != probe);
    }
    const int any = __syncthreads_or(my_nan);
    nan_hist = (y == yb0) ? (any ? 0x3fu : 0u) : ((nan_hist << 1) | (any ? 1u : 0u));
    const bool nanmode = (nan_hist & 0x3fu) != 0;
    const int nrows_b = (int)((yb1 - y) < NB ? (yb1 - y) : NB);
    const bool interior_rows = (y - R - 1 >= 0) && (y + NB + R + 1 < H);
    const bool more = y + NB < yb1;

    if (y == yb0 && vcol_ok) {
      int fk = 0;
      for (int t = 0; t < p.n_terms; ++t) {
        if (p.terms[t].kind != TERM_BOX_FUSED) continue;
        const int r = p.terms[t].r;
        double s = 0.0;
        int c = 0;
        for (int d = -r; d <= r; ++d) {
          float v = ring[slot_tab[d + R + 1] + vc];
          bool ok = v == v;
          s += ok ? (double)v : 0.0;
          c += ok;
        }
#pragma unroll
        for (int k = 0; k < FK_MAXF; ++k) if (k == fk) { rs[k] = s; rc[k] = c; }
        ++fk;
      }
    }

    const int64_t orow = y + hi;
    const bool hrow_ok = orow < yb1;
    const bool hfast = hrow_ok && !edge_strip && hjn == SEG && !nanmode;
    float acc[SEG], xr[SEG];
    {
      const float* xrow = ring + slot_tab[hi + R + 1];
#pragma unroll
      for (int jj = 0; jj < SEG; ++jj) {
        acc[jj] = 0.f;
        xr[jj] = (hrow_ok && jj < hjn) ? xrow[R + hj0 + jj] : 0.f;
      }
    }
    if (last_fused < 0) {   // no vertical pass will read the ring: the next rows can start right away
      __syncthreads();
      if (more) issue_rows(need_hi + 1, need_hi_of(y + NB));
      cp_async_commit();
    }
    int fk = 0;
    for (int t = 0; t < p.n_terms; ++t) {
      const DevTerm& T = p.terms[t];
      if (T.kind == TERM_BOX_FUSED) {
        const int r = T.r;
        const double n = (double)(2 * r + 1), inv = 1.0 / n;
        if (vcol_ok && vc >= R - r && vc < R + FK_TW + r) {
          double s = 0.0;
          int c = 0;
#pragma unroll
          for (int k = 0; k < FK_MAXF; ++k) if (k == fk) { s = rs[k]; c = rc[k]; }
          if (!nanmode && interior_rows) {
            // rows are consecutive ring slots: plain pointer increments, splitting the batch where
            // the incoming or the outgoing row wraps around the ring
            int sin = base + r + 1; if (sin >= NRING) sin -= NRING;
            int sout = base - r; if (sout < 0) sout += NRING;
            int i = 0;
            while (i < nrows_b) {
              int run = nrows_b - i;
              if (NRING - sin < run) run = NRING - sin;
              if (NRING - sout < run) run = NRING - sout;
              const float* pin = ring + (size_t)sin * SWp + vc;
              const float* pout = ring + (size_t)sout * SWp + vc;
              double* pv = plane64 + (size_t)i * SWp + vc;
              int k = 0;
              for (; k + 4 <= run; k += 4) {
                float a0 = pin[0], a1 = pin[SWp], a2 = pin[2 * SWp], a3 = pin[3 * SWp];
                float b0 = pout[0], b1 = pout[SWp], b2 = pout[2 * SWp], b3 = pout[3 * SWp];
                double d0 = (double)a0 - (double)b0, d1 = (double)a1 - (double)b1;
                double d2 = (double)a2 - (double)b2, d3 = (double)a3 - (double)b3;
                pv[0] = round_to_f32_grid(s * inv); s += d0;
                pv[SWp] = round_to_f32_grid(s * inv); s += d1;
                pv[2 * SWp] = round_to_f32_grid(s * inv); s += d2;
                pv[3 * SWp] = round_to_f32_grid(s * inv); s += d3;
                pin += 4 * SWp; pout += 4 * SWp; pv += 4 * SWp;
              }
              for (; k < run; ++k) {
                pv[0] = round_to_f32_grid(s * inv);
                s += (double)pin[0] - (double)pout[0];
                pin += SWp; pout += SWp; pv += SWp;
              }
              i += run;
              sin += run; if (sin >= NRING) sin -= NRING;
              sout += run; if (sout >= NRING) sout -= NRING;
            }
          } else {
            const int* tin = slot_tab + (r + 1 + R + 1);
            const int* tout = slot_tab + (-r + R + 1);
            if (!nanmode) {
              for (int i = 0; i < nrows_b; ++i) {
                plane64[i * SWp + vc] = round_to_f32_grid(s * inv);
                s += (double)ring[tin[i] + vc] - (double)ring[tout[i] + vc];
              }
            } else {
              for (int i = 0; i < nrows_b; ++i) {
                plane32[i * SWp + vc] = (float)div_by_count(s, n, inv);
                cplane[i * SWp + vc] = (unsigned char)c;
                float vin = ring[tin[i] + vc];
                float vout = ring[tout[i] + vc];
                bool oin = vin == vin, oout = vout == vout;
                s += (oin ? (double)vin : 0.0) - (oout ? (double)vout : 0.0);
                c += (int)oin - (int)oout;
              }
            }
          }
#pragma unroll
          for (int k = 0; k < FK_MAXF; ++k) if (k == fk) { rs[k] = s; rc[k] = c; }
        }
        __syncthreads();
        if (t == last_fused) {   // the ring rows this batch no longer needs are free: start the next copy
          if (more) issue_rows(need_hi + 1, need_hi_of(y + NB));
          cp_async_commit();
        }
        if (hfast) {
          const double* pl = plane64 + hi * SWp + R + hj0 - r;   // leftmost tap of the first window
          const double* pr = pl + 2 * r + 1;                     // first tap entering
          double sv = 0.0;
          for (int d = 0; d <= 2 * r; ++d) sv += pl[d];
          const float wgt = T.weight;
#pragma unroll
          for (int jj = 0; jj < SEG; ++jj) {
            float mean = (float)(sv * inv);
            acc[jj] = acc[jj] + wgt * (xr[jj] - mean);
            if (jj + 1 < SEG) sv += pr[jj] - pl[jj];
          }
        } else if (hrow_ok && hjn > 0) {
          const unsigned char* crow = cplane + hi * SWp;
          if (nanmode) {
            const float* vrow = plane32 + hi * SWp;
            if (edge_strip) hphase_generic<true, SEG, float>(vrow, crow, xr, r, hj0, hjn, x0, cs0, W, true, T.weight, acc);
            else hphase_generic<false, SEG, float>(vrow, crow, xr, r, hj0, hjn, x0, cs0, W, true, T.weight, acc);
          } else {
            const double* vrow = plane64 + hi * SWp;
            if (edge_strip) hphase_generic<true, SEG, double>(vrow, crow, xr, r, hj0, hjn, x0, cs0, W, false, T.weight, acc);
            else hphase_generic<false, SEG, double>(vrow, crow, xr, r, hj0, hjn, x0, cs0, W, false, T.weight, acc);
          }
        }
        __syncthreads();
        ++fk;
      } else if (T.kind == TERM_COARSE) {
        if (hrow_ok && hjn > 0) {
          // scipy zoom(order=1): coordinate = index*(n_in-1)/(n_out-1), weights [1-t, t], 'nearest' edge.
          // Rows are interpolated first (A = a0*(1-tr) + a1*tr at the two neighbouring coarse
          // columns), then one FMA along the row: A0 + tc*(A1 - A0).
          double ri = (double)orow * T.rscale;
          int64_t r0 = (int64_t)floor(ri);
          if (r0 > T.gh - 1) r0 = T.gh - 1;
          const double tr = ri - (double)r0;
          const int64_t r1 = r0 + 1 < T.gh ? r0 + 1 : T.gh - 1;
          const double wr0 = 1.0 - tr;
          const float* g0 = T.grid + (r0 - T.grow0) * T.gw;
          const float* g1 = T.grid + (r1 - T.grow0) * T.gw;
          const int gwm1 = (int)T.gw - 1;
          int c0 = 0;
          unsigned adv = 0u;
#pragma unroll
          for (int l = 0; l < MAX_LEVELS; ++l) if (l == T.lvl) { c0 = lv_c0[l]; adv = lv_adv[l]; }
          int c1 = c0 + 1 < gwm1 ? c0 + 1 : gwm1;
          double A0 = (double)__ldg(g0 + c0) * wr0 + (double)__ldg(g1 + c0) * tr;
          double A1 = (double)__ldg(g0 + c1) * wr0 + (double)__ldg(g1 + c1) * tr;
          const float wgt = T.weight;
          const double* tcp = tctab + T.lvl * FK_TW + hj0;
          if (hjn == SEG)
            coarse_taps<SEG, true>(g0, g1, wr0, tr, c0, c1, gwm1, adv, tcp, A0, A1, wgt, xr, acc, hjn);
          else
            coarse_taps<SEG, false>(g0, g1, wr0, tr, c0, c1, gwm1, adv, tcp, A0, A1, wgt, xr, acc, hjn);
        }
      } else {
        if (hrow_ok && hjn > 0) {
          const float* prow = T.grid + (orow - T.grow0) * W;
#pragma unroll
          for (int jj = 0; jj < SEG; ++jj)
            if (jj < hjn) acc[jj] = acc[jj] + T.weight * (xr[jj] - __ldg(prow + x0 + hj0 + jj));
        }
      }
    }

    float* stage = plane32;   // NB x (FK_TW + 1) floats, inside the plane region
    if (hrow_ok && hjn > 0) {
      float* sp = stage + hi * (FK_TW + 1) + hj0;
      if (nd.mode == 1) {
        // v / s, correctly rounded: q = v*rinv, one FMA residual correction (Markstein); NaN stays NaN
        const float sc = nd.sc, rinv = nd.rinv;
#pragma unroll
        for (int jj = 0; jj < SEG; ++jj) {
          if (jj < hjn) {
            float q = acc[jj] * rinv;
            float rem = fmaf(-q, sc, acc[jj]);
            sp[jj] = fmaf(rem, rinv, q);
          }
        }
      } else if (nd.mode == 2) {
#pragma unroll
        for (int jj = 0; jj < SEG; ++jj) if (jj < hjn) sp[jj] = (acc[jj] != acc[jj]) ? acc[jj] : 0.f;
      } else {
#pragma unroll
        for (int jj = 0; jj < SEG; ++jj) if (jj < hjn) sp[jj] = acc[jj];
      }
    }
    __syncthreads();
    {
      const int ncols = (W - x0) < FK_TW ? (W - x0) : FK_TW;
      for (int rr = tid / 32; rr < nrows_b; rr += FK_THREADS / 32) {   // one warp per output row
        const float* srow = stage + rr * (FK_TW + 1) + (tid & 31);
        const int64_t obase = (y + rr - p.out_row0) * p.ld_out + x0 + (tid & 31);
        if (p.enc.kind == FSG_OUT_F32 && ncols == FK_TW) {
          float* o = (float*)p.out + obase;
#pragma unroll
          for (int c = 0; c < FK_TW / 32; ++c) o[c * 32] = srow[c * 32];
          if ((tid & 31) < FK_TW % 32) o[(FK_TW / 32) * 32] = srow[(FK_TW / 32) * 32];
        } else if (p.enc.kind == FSG_OUT_F32) {
          float* o = (float*)p.out + obase;
          for (int c = 0; c + (tid & 31) < ncols; c += 32) o[c] = srow[c];
        } else {
          for (int c = 0; c + (tid & 31) < ncols; c += 32) store_out(p.out, obase + c, srow[c], p.enc);
        }
      }
    }
    __syncthreads();   // the plane (staging) is free again before the next batch's vertical pass
    base += NB;
    if (base >= NRING) base -= NRING;
  }
  cp_async_wait_all();
}

template <int NB>
static size_t fused_fast_smem_bytes(int R, int n_lvls) {
  size_t SWp = (size_t)((FK_TW + 2 * R) | 1);
  size_t nring = NB + 2 * R + 1;
  size_t off_plane = align_up(nring * SWp * 4, 16);
  size_t off_slot = off_plane + (size_t)NB * SWp * 8;
  size_t off_tc = align_up(off_slot + (nring + 1) * 4, 16);
  return off_tc + (size_t)n_lvls * FK_TW * 8;
}

static size_t fused_smem_bytes(int R) {
  size_t SWp = (size_t)((FK_TW + 2 * R) | 1);
  size_t nring = FK_NB + 2 * R + 1;
  size_t vp = (size_t)FK_NB * SWp;
  size_t stage = (size_t)FK_NB * (FK_TW + 1);
  if (stage > vp) vp = stage;
  return nring * SWp * 4 + vp * 4 + align_up((size_t)FK_NB * SWp, 16);
}

}  // namespace fsg
#include "fsg_topousm_v6.cuh"
#include "fsg_topousm_v8.cuh"
namespace fsg {

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
// Debug switches, read once per process (FSG_FORCE_GENERIC, FSG_FUSED_V5, FSG_NO_BULK, FSG_V6_CFGB, FSG_NO_V8).
struct FusedSwitches {
  bool force_generic, v5, no_bulk, cfgb, no_v8, no_side, box_two_pass;
  FusedSwitches()
      : force_generic(getenv("FSG_FORCE_GENERIC") != nullptr), v5(getenv("FSG_FUSED_V5") != nullptr),
        no_bulk(getenv("FSG_NO_BULK") != nullptr), cfgb(getenv("FSG_V6_CFGB") != nullptr),
        no_v8(getenv("FSG_NO_V8") != nullptr), no_side(getenv("FSG_NO_SIDE_STREAM") != nullptr),
        box_two_pass(getenv("FSG_BOX_TWO_PASS") != nullptr) {}
};
static const FusedSwitches& fused_switches() {
  static const FusedSwitches sw;
  return sw;
}
extern "C" void fsg_debug_reload_switches(void) { const_cast<FusedSwitches&>(fused_switches()) = FusedSwitches(); }

static int mean_on_grid(const Grid& g, int size, float* out, float* tv, float* tw, double* taps, int64_t oy0,
                        int64_t oh, cudaStream_t s) {
  int rc;
  if (size == 0) {  // sigma = 1 gaussian, 'nearest'
    if ((rc = launch_gauss_taps(1.0, 4, taps, s))) return rc;
    if ((rc = launch_gauss_axis0(g, taps, 4, tv, tw, nullptr, oy0, oh, s))) return rc;
    return launch_gauss_axis1(tv, tw, oh, g.w, taps, 4, COMBINE_MEAN, out, nullptr, nullptr, s);
  }
  int slot = prof_begin(PROF_TOPOUSM_COARSE, s);
  if (!fused_switches().box_two_pass && launch_box_mean2d(g, size, out, oy0, oh, s, &rc)) {
    prof_end(slot, s);
    return rc;
  }
  if ((rc = launch_box_axis0(g, size, tv, tw, oy0, oh, s))) { prof_end(slot, s); return rc; }
  rc = launch_box_axis1(tv, tw, oh, g.w, size, out, s);
  prof_end(slot, s);
  return rc;
}

static int launch_pyramid(const float* dem, int64_t rows, int64_t W, int64_t ld, int n_levels, const int* factors,
                          float* const* grids, int* flags, cudaStream_t s) {
  PyramidParams pp{};
  pp.dem = dem; pp.H = rows; pp.W = W; pp.ld = ld; pp.n_levels = n_levels; pp.flags = flags;
  for (int k = 0; k < n_levels; ++k) {
    pp.f[k] = factors[k];
    pp.grid[k] = grids[k];
    pp.gw[k] = (W + factors[k] - 1) / factors[k];
    pp.gh[k] = (rows + factors[k] - 1) / factors[k];
  }
  dim3 grid((unsigned)((W + PY_COLS - 1) / PY_COLS), (unsigned)((rows + PY_ROWS - 1) / PY_ROWS));
  int pslot = prof_begin(PROF_TOPOUSM_PYRAMID, s);
  pyramid_kernel<<<grid, 256, 0, s>>>(pp);
  prof_end(pslot, s);
  FSG_LAUNCH_OK();
  return FSG_OK;
}


// Rows per CTA band: few enough that the (2R+1)-row warm-up per band stays small, enough CTAs (>= 6 per SM when
// the raster allows) that the slower edge strips and the SM-to-SM spread average out.  (A "fewest waves" model
// was tried and lost 25 %: one or two waves leave the chip waiting for the edge-strip CTAs.)
static int64_t fused_band_rows(int64_t rows, int64_t strips, int fused_R, int64_t multiple) {
  int64_t want_bands = (rows + 1023) / 2048;
  if (want_bands < 1) want_bands = 1;
  const int64_t min_rows = 8 * (int64_t)(2 * fused_R + 1);
  while (strips * want_bands < 148 * 6 && (rows + 2 * want_bands - 1) / (2 * want_bands) >= min_rows) want_bands *= 2;
  // small launches (the statistics windows): filling the 148 SMs matters more than the warm-up rows
  while (strips * want_bands < 148 && (rows + 2 * want_bands - 1) / (2 * want_bands) >= 3 * (int64_t)(2 * fused_R + 1))
    want_bands *= 2;
  int64_t band_rows = (rows + want_bands - 1) / want_bands;
  band_rows = (band_rows + multiple - 1) / multiple * multiple;
  if (band_rows > (1 << 30)) band_rows = 1 << 30;
  return band_rows;
}

// One fused_kernel_v6 launch over output rows [r0, r1) x columns [c0, c1) of the raster (`base` describes the
// whole call; its out pointer belongs to row base.out_row0).
static int launch_v6_rect(const FusedParams& base, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaStream_t s,
                          bool spread = false) {
  if (r1 <= r0 || c1 <= c0) return FSG_OK;
  FusedParams fp = base;
  fp.out = (unsigned char*)base.out + (size_t)(r0 - base.out_row0) * (size_t)base.ld_out * out_elem_size(base.enc.kind);
  fp.out_row0 = r0;
  fp.out_rows = r1 - r0;
  fp.strip0 = 0;
  fp.col0 = (int)c0;
  fp.col_end = (int)c1;
  fp.tile_flags = nullptr;
  if (c0 % 4 != 0) fp.bulk_ok = 0;
  const int64_t strips = (c1 - c0 + FK_TW - 1) / FK_TW;
  int64_t band_rows = fused_band_rows(fp.out_rows, strips, fp.R, FK_NB);
  // a rectangle one or two strips wide next to a large launch (the left / right raster border): many short bands,
  // so that it finishes inside the large launch instead of trailing it as a few long serial CTAs
  while (spread && band_rows > 4 * (int64_t)(2 * fp.R + 1) && strips * ((fp.out_rows + band_rows - 1) / band_rows) < 148)
    band_rows = ((band_rows + 1) / 2 + FK_NB - 1) / FK_NB * FK_NB;
  fp.band_rows = (int)band_rows;
  const int64_t bands = (fp.out_rows + band_rows - 1) / band_rows;
  if (bands > 65535) return fail(FSG_E_UNSUPPORTED, "fsg_topousm_fast: raster too tall");
  const size_t smem = V6Geom<32, V6CfgA>::BYTES;
  FSG_CUDA_OK(cudaFuncSetAttribute(fused_kernel_v6<32, V6CfgA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fused_kernel_v6<32, V6CfgA><<<dim3((unsigned)strips, (unsigned)bands), V6Geom<32, V6CfgA>::THREADS, smem, s>>>(fp);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

// Rows per CTA of the v8 launch (a multiple of the 16-row batch; a CTA that meets a NaN flags the 256-row block
// and restarts behind it, wherever its band ends).  One CTA per SM, CTAs handed out in launch order: the makespan of
// every candidate is computed with a CTA costing (rows + the rows a restart is worth) and the shortest one wins --
// 8128 rows x 341 strips: two bands of 3792 rows and a 544-row remainder that fills the SMs with one full CTA fewer
// (fixed 2048-row bands left a 0.2-wave tail: 6.1 -> 5.44 ms); a 4128^2 statistics window: one wave of 154 CTAs.
static int64_t v8_band_rows(int64_t rows, int64_t strips) {
  static std::mutex mu;
  static std::map<std::pair<int64_t, int64_t>, int64_t> memo;
  std::lock_guard<std::mutex> lock(mu);
  auto it = memo.find({rows, strips});
  if (it != memo.end()) return it->second;
  const int64_t restart = 96;   // fill of five ring groups + the window sums, in rows of steady-state work
  int64_t best = V8_BLK;
  double best_t = 1e300;
  for (int64_t br = 4 * V8_NB; br <= 4096 && br < rows + V8_NB; br += V8_NB) {
    const int64_t bands = (rows + br - 1) / br;
    if (bands > 65535) continue;
    const int64_t last = rows - (bands - 1) * br;
    const int64_t c1 = br + restart, c2 = last + restart;
    // list scheduling on 148 SMs in launch order: the strips * (bands - 1) equal CTAs fill the SMs evenly ...
    const int64_t n_full = strips * (bands - 1);
    const int64_t q = n_full / 148, r = n_full % 148;
    std::priority_queue<int64_t, std::vector<int64_t>, std::greater<int64_t>> sm;
    for (int i = 0; i < 148; ++i) sm.push((q + (i < r ? 1 : 0)) * c1);
    int64_t mk = (q + (r ? 1 : 0)) * c1;
    // ... and the CTAs of the last (shorter) band go to whichever SM frees up first
    for (int64_t c = 0; c < strips; ++c) {
      const int64_t t = sm.top() + c2;
      sm.pop();
      sm.push(t);
      mk = t > mk ? t : mk;
    }
    if ((double)mk < best_t) { best_t = (double)mk; best = br; }
  }
  memo[{rows, strips}] = best;
  return best;
}

// The border rectangles of a v8 launch run on a side stream (forked from / joined to the caller's stream with
// events) so that their few CTAs share the chip with the interior launch instead of running after it.  One lane
// per host thread and device.
struct SideLane {
  cudaStream_t st = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int dev = -1;
};
static SideLane* side_lane() {
  thread_local SideLane lanes[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  SideLane& L = lanes[dev & 15];
  if (L.st && L.dev == dev) return &L;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&L.st, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&L.join, cudaEventDisableTiming) != cudaSuccess) {
    L.st = nullptr;
    return nullptr;
  }
  L.dev = dev;
  return &L;
}

// bytes of the v8 NaN-block flags for a raster (or band) of H x W
static size_t v8_flag_bytes(int64_t rows, int64_t W) {
  return (size_t)((rows + V8_BLK - 1) / V8_BLK + 1) * (size_t)(W / V8_TW + 2) * sizeof(int);
}

// Interior fast path: fused_kernel_v8 on the rows / strips whose whole neighbourhood lies inside the raster (and
// inside the DEM rows the caller holds), fused_kernel_v6 on the four border rectangles and on the 256-row
// blocks v8 flagged (NaN or Inf in reach).  Returns 1 when the launch was taken, 0 when v8 does not apply.
static int launch_fused_v8(FusedParams& fp, int* flags, size_t flag_bytes, cudaStream_t s, int* rc_out) {
  *rc_out = FSG_OK;
  const int64_t H = fp.H, W = fp.W;
  if (fp.n_terms < 3 || fp.n_terms > 3 + V8_MAXC || !fp.bulk_ok || !flags) return 0;
  if (fp.terms[0].kind != TERM_BOX_FUSED || fp.terms[0].r != 2 || fp.terms[1].kind != TERM_BOX_FUSED ||
      fp.terms[1].r != 8 || fp.terms[2].kind != TERM_BOX_FUSED || fp.terms[2].r != 32)
    return 0;
  for (int i = 3; i < fp.n_terms; ++i)   // decimated terms only, decimation >= 4 (staged coarse cells: V8_CROWS x V8_CCOLS)
    if (fp.terms[i].kind != TERM_COARSE || fp.terms[i].cscale > 0.25 || fp.terms[i].rscale > 0.25) return 0;
  const int64_t R0 = fp.out_row0, R1 = fp.out_row0 + fp.out_rows;
  int64_t C0 = 0, C1 = W;
  if (fp.roi_cols > 0 && fp.roi_cols < W) { C0 = fp.roi_col0; C1 = fp.roi_col0 + fp.roi_cols; }
  const int64_t lo = fp.dem_row0 > 0 ? fp.dem_row0 : 0;
  const int64_t hi = fp.dem_row0 + fp.dem_rows < H ? fp.dem_row0 + fp.dem_rows : H;
  // rows: the window of row y needs rows y - 32 .. y + 32 (a batch of 16 rows: the five ring groups around it)
  int64_t Ya = R0 > lo + V8_RH ? R0 : lo + V8_RH;
  int64_t ymax = R1 < hi - V8_RH ? R1 : hi - V8_RH;
  int64_t Yb = Ya + (ymax - Ya) / V8_NB * V8_NB;
  // columns: strips of V8_TW from a multiple of 4 (16-byte aligned bulk copies) with 32 columns on either side
  int64_t Xa = C0 > V8_RH ? C0 : V8_RH;
  Xa = (Xa + 3) / 4 * 4;
  int64_t xmax = C1 < W - V8_RH ? C1 : W - V8_RH;
  int64_t n8 = xmax > Xa ? (xmax - Xa) / V8_TW : 0;
  int64_t Xb = Xa + n8 * V8_TW;
  if (n8 < 1 || Yb - Ya < 4 * V8_NB) return 0;
  int64_t xlast = Xb - V8_TW;
  if (Xb < xmax && (xmax - V8_TW) % 4 == 0) {   // one more strip, shifted left to end at xmax (the overlap is computed
    xlast = xmax - V8_TW;                        // twice, same values): no narrow rectangle left for the general kernel
    ++n8;
    Xb = xmax;
  }
  const int64_t nblk = (Yb - Ya + V8_BLK - 1) / V8_BLK;
  if ((size_t)nblk * (size_t)n8 * sizeof(int) > flag_bytes) return 0;

  const bool is_roi = (fp.roi_cols > 0 && fp.roi_cols < W) || (fp.dem_rows == H && fp.out_rows < H);
  int slot = prof_begin(is_roi ? PROF_TOPOUSM_FUSED_ROI : PROF_TOPOUSM_FUSED, s);
  auto done = [&](int rc) { prof_end(slot, s); *rc_out = rc; return 1; };
  if (cudaMemsetAsync(flags, 0, (size_t)nblk * (size_t)n8 * sizeof(int), s) != cudaSuccess)
    return done(fail(FSG_E_CUDA, "fsg_topousm_fast: clearing the block flags failed"));
  fp.v8_col0 = (int)Xa; fp.v8_xlast = (int)xlast; fp.v8_row0 = Ya; fp.v8_row1 = Yb; fp.v8_flags = flags;
  fp.col0 = 0; fp.col_end = 0; fp.tile_flags = nullptr; fp.strip0 = 0;
  // raster borders (and rows / columns outside the v8 grid): on the side stream, ahead of the interior launch
  SideLane* lane = fused_switches().no_side ? nullptr : side_lane();
  {
    cudaStream_t bs = s;
    if (lane) {
      if (cudaEventRecord(lane->fork, s) != cudaSuccess || cudaStreamWaitEvent(lane->st, lane->fork, 0) != cudaSuccess)
        return done(fail(FSG_E_CUDA, "fsg_topousm_fast: forking the border stream failed"));
      bs = lane->st;
    }
    int rc;
    if ((rc = launch_v6_rect(fp, Ya, Yb, C0, Xa, bs, true))) return done(rc);
    if ((rc = launch_v6_rect(fp, Ya, Yb, Xb, C1, bs, true))) return done(rc);
    if ((rc = launch_v6_rect(fp, R0, Ya, C0, C1, bs))) return done(rc);
    if ((rc = launch_v6_rect(fp, Yb, R1, C0, C1, bs))) return done(rc);
    if (lane && cudaEventRecord(lane->join, lane->st) != cudaSuccess)
      return done(fail(FSG_E_CUDA, "fsg_topousm_fast: joining the border stream failed"));
  }
  int64_t band_rows = v8_band_rows(Yb - Ya, n8);
  fp.band_rows = (int)band_rows;
  const int64_t bands = (Yb - Ya + band_rows - 1) / band_rows;
  if (bands > 65535) return done(fail(FSG_E_UNSUPPORTED, "fsg_topousm_fast: raster too tall"));
  void (*kern)(FusedParams) = fp.n_terms == 3 ? fused_kernel_v8<0> : (fp.n_terms == 4 ? fused_kernel_v8<1> : (fp.n_terms == 5 ? fused_kernel_v8<2> : fused_kernel_v8<3>));
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V8Smem::BYTES) != cudaSuccess)
    return done(fail(FSG_E_CUDA, "fsg_topousm_fast: shared-memory attribute of the v8 kernel"));
  kern<<<dim3((unsigned)n8, (unsigned)bands), V8_THREADS, V8Smem::BYTES, s>>>(fp);
  count_launch();
  if (cudaPeekAtLastError() != cudaSuccess)
    return done(fail(FSG_E_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(cudaPeekAtLastError()), __FILE__, __LINE__));
  // blocks v8 flagged: same tiles on v6
  {
    FusedParams tp = fp;
    tp.tile_flags = flags; tp.tile_w = V8_TW; tp.tile_rows = V8_BLK; tp.tile_row0 = Ya; tp.tile_row1 = Yb;
    tp.col0 = (int)Xa; tp.col_end = (int)Xb;
    const size_t smem = V6Geom<32, V6CfgA>::BYTES;
    if (cudaFuncSetAttribute(fused_kernel_v6<32, V6CfgA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return done(fail(FSG_E_CUDA, "fsg_topousm_fast: shared-memory attribute of the v6 kernel"));
    fused_kernel_v6<32, V6CfgA><<<dim3((unsigned)n8, (unsigned)nblk), V6Geom<32, V6CfgA>::THREADS, smem, s>>>(tp);
    count_launch();
    if (cudaPeekAtLastError() != cudaSuccess)
      return done(fail(FSG_E_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(cudaPeekAtLastError()), __FILE__, __LINE__));
  }
  if (lane && cudaStreamWaitEvent(s, lane->join, 0) != cudaSuccess)
    return done(fail(FSG_E_CUDA, "fsg_topousm_fast: joining the border stream failed"));
  return done(FSG_OK);
}

// Fills the per-launch fields of `fp` (kernel variant, grid, shared memory) and launches the fused pass
// over global output rows [fp.out_row0, +fp.out_rows).  `v8_flags` (may be NULL): scratch for the interior
// fast path.
static int launch_fused(FusedParams& fp, int fused_R, int n_levels, cudaStream_t s, int* v8_flags = nullptr,
                        size_t v8_flag_cap = 0) {
  const int64_t H = fp.H, W = fp.W;
  const FusedSwitches& sw = fused_switches();
  fp.R = fused_R;
  fp.n_lvls = n_levels;
  fp.col0 = 0; fp.col_end = 0; fp.tile_flags = nullptr;
  if (fp.out_rows <= 0) return FSG_OK;
  const bool fast_ok = H >= fused_R + 2 && W >= fused_R + 2 && W < (1 << 30) && !sw.force_generic;
  const size_t smem_cap = 227 * 1024;
  int nb = 0;   // rows per batch of the fast kernel (0: general kernel)
  int n_fused = 0;
  for (int i = 0; i < fp.n_terms; ++i) n_fused += fp.terms[i].kind == TERM_BOX_FUSED;
  const bool v6_ok = fast_ok && fused_R <= 32 && n_levels <= V6_MAXLV && n_fused <= V6_MAXF && !sw.v5;
  fp.bulk_ok = (((uintptr_t)fp.dem & 15) == 0 && fp.ld_in % 4 == 0 && !sw.no_bulk) ? 1 : 0;
  {
    const size_t esz = out_elem_size(fp.enc.kind);
    const uintptr_t need = fp.enc.kind == FSG_OUT_F32 ? 15 : 3;
    fp.out_vec_ok = (((uintptr_t)fp.out & need) == 0 && fp.ld_out % 4 == 0 && esz != 2) ? 1 : 0;
  }
  // v6 geometry B (640 threads, 12-pixel segments) needs decimation >= 4 for its coarse cell slots
  bool v6b_ok = v6_ok && sw.cfgb;
  for (int l = 0; l < n_levels; ++l) v6b_ok = v6b_ok && fp.lvl_cscale[l] <= 0.25;
  if (v6_ok && !v6b_ok && !sw.no_v8 && fused_R == 32) {
    int rc;
    if (launch_fused_v8(fp, v8_flags, v8_flag_cap, s, &rc)) return rc;
  }
  if (v6b_ok) nb = 62;
  else if (v6_ok) nb = 6;
  else if (fast_ok && fused_fast_smem_bytes<32>(fused_R, n_levels) <= smem_cap) nb = 32;
  else if (fast_ok && fused_fast_smem_bytes<16>(fused_R, n_levels) <= smem_cap) nb = 16;
  const int64_t rows = fp.out_rows;
  const int tw = nb == 62 ? V6CfgB::TW : FK_TW;
  int64_t strips = (W + tw - 1) / tw;
  fp.strip0 = 0;
  if (fp.roi_cols > 0 && fp.roi_cols < W) {   // only the strips that overlap the requested columns
    const int64_t s_lo = fp.roi_col0 / tw, s_hi = (fp.roi_col0 + fp.roi_cols + tw - 1) / tw;
    fp.strip0 = (int)s_lo;
    strips = s_hi - s_lo;
  }
  const int64_t band_rows = fused_band_rows(rows, strips, fused_R, FK_NB);
  fp.band_rows = (int)band_rows;
  int64_t bands = (rows + band_rows - 1) / band_rows;
  if (bands > 65535) return fail(FSG_E_UNSUPPORTED, "fsg_topousm_fast: raster too tall");
  dim3 grid((unsigned)strips, (unsigned)bands);
  size_t smem = nb == 32 ? fused_fast_smem_bytes<32>(fused_R, n_levels)
                         : (nb == 16 ? fused_fast_smem_bytes<16>(fused_R, n_levels) : fused_smem_bytes(fused_R));
  if (nb == 6) smem = V6Geom<32, V6CfgA>::BYTES;
  if (nb == 62) smem = V6Geom<32, V6CfgB>::BYTES;
  if (nb == 6) FSG_CUDA_OK(cudaFuncSetAttribute(fused_kernel_v6<32, V6CfgA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (nb == 62) FSG_CUDA_OK(cudaFuncSetAttribute(fused_kernel_v6<32, V6CfgB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (nb == 32) FSG_CUDA_OK(cudaFuncSetAttribute(fused_kernel_fast<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (nb == 16) FSG_CUDA_OK(cudaFuncSetAttribute(fused_kernel_fast<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else FSG_CUDA_OK(cudaFuncSetAttribute(fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool is_roi = (fp.roi_cols > 0 && fp.roi_cols < W) || (fp.dem_rows == H && fp.out_rows < H);
  int slot = prof_begin(is_roi ? PROF_TOPOUSM_FUSED_ROI : PROF_TOPOUSM_FUSED, s);
  if (nb == 6) fused_kernel_v6<32, V6CfgA><<<grid, V6Geom<32, V6CfgA>::THREADS, smem, s>>>(fp);
  else if (nb == 62) fused_kernel_v6<32, V6CfgB><<<grid, V6Geom<32, V6CfgB>::THREADS, smem, s>>>(fp);
  else if (nb == 32) fused_kernel_fast<32><<<grid, FK_THREADS, smem, s>>>(fp);
  else if (nb == 16) fused_kernel_fast<16><<<grid, FK_THREADS, smem, s>>>(fp);
  else fused_kernel<<<grid, FK_THREADS, smem, s>>>(fp);
  prof_end(slot, s);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

static void set_norm(FusedParams& fp, double norm_scale) {
  if (is_none(norm_scale)) fp.norm_mode = 0;
  else if (norm_scale > 0.0) {
    fp.norm_mode = 1; fp.norm_scale = (float)norm_scale; fp.norm_rinv = (float)(1.0 / (double)fp.norm_scale);
  } else fp.norm_mode = 2;
}

static int run_topousm(const float* dem, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out,
                       const int32_t* radii, const float* weights, int n, double pixel_size, double norm_scale,
                       const fsg_encode* enc, void* ws, size_t ws_bytes, cudaStream_t s, int64_t roi_row0 = 0,
                       int64_t roi_rows = -1, int64_t roi_col0 = 0, int64_t roi_cols = -1) {
  if (roi_rows < 0) roi_rows = H;
  if (roi_cols < 0) roi_cols = W;
  if (roi_row0 < 0 || roi_col0 < 0 || roi_row0 + roi_rows > H || roi_col0 + roi_cols > W)
    return fail(FSG_E_INVALID, "fsg_topousm_fast_roi: region of interest outside the raster");
  HostPlan plan;
  int rc = make_plan(H, W, radii, n, pixel_size, &plan);
  if (rc) return rc;
  if (!dem || !out) return fail(FSG_E_INVALID, "fsg_topousm_fast: NULL buffer");
  if (ld_in < W || ld_out < W) return fail(FSG_E_INVALID, "fsg_topousm_fast: row stride smaller than width");
  if (plan.total > 0 && (!ws || ws_bytes < plan.total))
    return fail(FSG_E_WORKSPACE, "fsg_topousm_fast: workspace of %zu bytes needed, %zu given", plan.total, ws_bytes);
  unsigned char* base = (unsigned char*)ws;
  float* tv = (float*)(base + plan.off_tv);
  float* tw = (float*)(base + plan.off_tw);
  double* taps = (double*)(base + plan.off_taps);
  int* flags = (int*)(base + plan.off_flags);

  if (plan.n_levels > 0) {
    zero_flags_kernel<<<1, 32, 0, s>>>(flags);
    FSG_LAUNCH_OK();
    int factors[MAX_LEVELS];
    float* grids[MAX_LEVELS];
    for (int k = 0; k < plan.n_levels; ++k) { factors[k] = plan.levels[k].f; grids[k] = (float*)(base + plan.levels[k].off); }
    if ((rc = launch_pyramid(dem, H, W, ld_in, plan.n_levels, factors, grids, flags, s))) return rc;
    // enclosed-void fill (runs only when the level has an all-NaN cell; device-side flag)
    for (int k = 0; k < plan.n_levels; ++k) {
      const HostLevel& l = plan.levels[k];
      double sigma = (double)(l.h < l.w ? l.h : l.w) / 64.0;
      if (sigma < 1.0) sigma = 1.0;
      int radius = (int)(4.0 * sigma + 0.5);
      float* grid_k = (float*)(base + l.off);
      Grid g{grid_k, l.h, l.w, l.w};
      if ((rc = launch_gauss_taps(sigma, radius, taps, s))) return rc;
      if ((rc = launch_gauss_axis0(g, taps, radius, tv, tw, flags + k, 0, l.h, s, base + plan.off_need))) return rc;
      if ((rc = launch_gauss_axis1(tv, tw, l.h, l.w, taps, radius, COMBINE_VOIDFILL, grid_k, flags + k, flags + 4 + k, s,
                                   base + plan.off_need)))
        return rc;
    }
  }

  FusedParams fp{};
  fp.dem = dem; fp.out = out; fp.H = H; fp.W = W; fp.ld_in = ld_in; fp.ld_out = ld_out;
  fp.dem_row0 = 0; fp.dem_rows = H; fp.out_row0 = roi_row0; fp.out_rows = roi_rows;
  fp.out = (unsigned char*)out + (size_t)roi_row0 * (size_t)ld_out * out_elem_size(enc ? enc->kind : FSG_OUT_F32);
  fp.roi_col0 = roi_col0; fp.roi_cols = roi_cols;
  fp.n_terms = n;
  fp.enc = make_encode(enc);
  set_norm(fp, norm_scale);
  for (int i = 0; i < n; ++i) {
    const HostTerm& t = plan.terms[i];
    DevTerm& d = fp.terms[i];
    d.kind = t.kind; d.r = t.radius; d.weight = weights[i]; d.grow0 = 0;
    if (t.kind == TERM_COARSE) {
      const HostLevel& l = plan.levels[t.level];
      float* mean = (float*)(base + t.grid_off);
      Grid g{(const float*)(base + l.off), l.h, l.w, l.w};
      if ((rc = mean_on_grid(g, t.size, mean, tv, tw, taps, 0, l.h, s))) return rc;
      d.grid = mean; d.gh = l.h; d.gw = l.w;
      // scipy.ndimage.zoom: zoom = (n_in - 1) / (n_out - 1)  (1.0 when n_out == 1)
      d.rscale = H > 1 ? (double)(l.h - 1) / (double)(H - 1) : 1.0;
      d.cscale = W > 1 ? (double)(l.w - 1) / (double)(W - 1) : 1.0;
      d.lvl = t.level;
      fp.lvl_cscale[t.level] = d.cscale;
      fp.lvl_gw[t.level] = (int)l.w;
    } else if (t.kind == TERM_PLANE) {
      float* mean = (float*)(base + t.grid_off);
      Grid g{dem, H, W, ld_in};
      if ((rc = mean_on_grid(g, t.size, mean, tv, tw, taps, 0, H, s))) return rc;
      d.grid = mean; d.gh = H; d.gw = W;
    }
  }
  return launch_fused(fp, plan.fused_R, plan.n_levels, s, (int*)(base + plan.off_v8flags), plan.v8flag_bytes);
}

// ---- row-band shard entry points (multi-GPU; see fujishadergpu_b200/core/sharding.py) ------------
static int run_fused_band(const float* dem, int64_t dem_row0, int64_t dem_rows, int64_t H, int64_t W, int64_t ld_in,
                          void* out, int64_t out_row0, int64_t out_rows, int64_t ld_out, const int32_t* radii,
                          const float* weights, int n, double pixel_size, const float* const* grids,
                          const int64_t* grow0, const int64_t* grows, double norm_scale, const float* norm_scale_dev,
                          const fsg_encode* enc, void* ws, size_t ws_bytes, cudaStream_t s) {
  HostPlan plan;
  int rc = make_plan(H, W, radii, n, pixel_size, &plan);
  if (rc) return rc;
  if (!dem || !out) return fail(FSG_E_INVALID, "fsg_topousm_fused_band: NULL buffer");
  if (out_row0 < 0 || out_rows < 0 || out_row0 + out_rows > H)
    return fail(FSG_E_INVALID, "fsg_topousm_fused_band: output rows outside the raster");
  const int R = plan.fused_R;
  int64_t lo = out_row0 - R < 0 ? 0 : out_row0 - R;
  int64_t hi = out_row0 + out_rows + R > H ? H : out_row0 + out_rows + R;
  if (dem_row0 > lo || dem_row0 + dem_rows < hi)
    return fail(FSG_E_INVALID, "fsg_topousm_fused_band: DEM rows [%lld,%lld) do not cover the %d-row halo [%lld,%lld)",
                (long long)dem_row0, (long long)(dem_row0 + dem_rows), R, (long long)lo, (long long)hi);
  FusedParams fp{};
  fp.dem = dem; fp.out = out; fp.H = H; fp.W = W; fp.ld_in = ld_in; fp.ld_out = ld_out;
  fp.dem_row0 = dem_row0; fp.dem_rows = dem_rows; fp.out_row0 = out_row0; fp.out_rows = out_rows;
  fp.n_terms = n;
  fp.enc = make_encode(enc);
  set_norm(fp, norm_scale);
  if (norm_scale_dev) { fp.norm_mode = 3; fp.norm_scale_dev = norm_scale_dev; }
  for (int i = 0; i < n; ++i) {
    const HostTerm& t = plan.terms[i];
    DevTerm& d = fp.terms[i];
    d.kind = t.kind; d.r = t.radius; d.weight = weights[i]; d.grow0 = 0;
    if (t.kind == TERM_COARSE || t.kind == TERM_PLANE) {
      if (!grids || !grids[i] || !grow0 || !grows)
        return fail(FSG_E_INVALID, "fsg_topousm_fused_band: term %d (radius %d) needs a precomputed mean grid", i, t.radius);
      d.grid = grids[i]; d.grow0 = grow0[i];
      if (t.kind == TERM_COARSE) {
        const HostLevel& l = plan.levels[t.level];
        d.gh = l.h; d.gw = l.w;
        d.rscale = H > 1 ? (double)(l.h - 1) / (double)(H - 1) : 1.0;
        d.cscale = W > 1 ? (double)(l.w - 1) / (double)(W - 1) : 1.0;
        d.lvl = t.level;
        fp.lvl_cscale[t.level] = d.cscale;
        fp.lvl_gw[t.level] = (int)l.w;
        // rows the bilinear taps of this band touch
        int64_t c_lo = (int64_t)floor((double)out_row0 * d.rscale);
        int64_t c_hi = (int64_t)floor((double)(out_row0 + out_rows - 1) * d.rscale) + 1;
        if (c_hi > l.h - 1) c_hi = l.h - 1;
        if (grow0[i] > c_lo || grow0[i] + grows[i] <= c_hi)
          return fail(FSG_E_INVALID, "fsg_topousm_fused_band: mean grid rows [%lld,%lld) of term %d do not cover [%lld,%lld]",
                      (long long)grow0[i], (long long)(grow0[i] + grows[i]), i, (long long)c_lo, (long long)c_hi);
      } else {
        d.gh = H; d.gw = W;
        if (grow0[i] > out_row0 || grow0[i] + grows[i] < out_row0 + out_rows)
          return fail(FSG_E_INVALID, "fsg_topousm_fused_band: mean plane of term %d does not cover the output rows", i);
      }
    }
  }
  const size_t need = v8_flag_bytes(out_rows, W);
  const bool ws_ok = ws && ws_bytes >= need && ((uintptr_t)ws & 3) == 0;
  return launch_fused(fp, plan.fused_R, plan.n_levels, s, ws_ok ? (int*)ws : nullptr, ws_ok ? need : 0);
}

// ------------------------------------------------------------------------------------------
// stand-alone helpers (exposed for parity tests and for the reference's other call sites)
// ------------------------------------------------------------------------------------------
// _upsample_to_shape (algorithms/_nan_utils.py:671-698): zoom(order=1) with align-corners mapping.
// NaN-aware branch when *has_nan != 0: zoom of filled & valid, num/max(den,1e-6) where den>1e-3.
__global__ void __launch_bounds__(256) upsample_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t h,
                                                       int64_t w, int64_t H, int64_t W, double rscale, double cscale,
                                                       const int* has_nan) {
  int64_t x = (int64_t)blockIdx.x * 256 + threadIdx.x;
  int64_t y = (int64_t)blockIdx.y + (int64_t)blockIdx.z * 32768;
  if (x >= W || y >= H) return;
  double ri = (double)y * rscale, ci = (double)x * cscale;
  int64_t r0 = (int64_t)floor(ri), c0 = (int64_t)floor(ci);
  if (r0 > h - 1) r0 = h - 1;
  if (c0 > w - 1) c0 = w - 1;
  double tr = ri - (double)r0, tc = ci - (double)c0;
  int64_t r1 = r0 + 1 < h ? r0 + 1 : h - 1, c1 = c0 + 1 < w ? c0 + 1 : w - 1;
  float a00 = in[r0 * w + c0], a01 = in[r0 * w + c1], a10 = in[r1 * w + c0], a11 = in[r1 * w + c1];
  double wr0 = 1.0 - tr, wr1 = tr, wc0 = 1.0 - tc, wc1 = tc;
  if (!*has_nan) {
    double v = ((double)a00 * wr0) * wc0;
    v += ((double)a01 * wr0) * wc1;
    v += ((double)a10 * wr1) * wc0;
    v += ((double)a11 * wr1) * wc1;
    out[y * W + x] = (float)v;
  } else {
    auto val = [](float a) { return a != a ? 0.0 : (double)a; };
    auto ok = [](float a) { return a != a ? 0.0 : 1.0; };
    double n = (val(a00) * wr0) * wc0; n += (val(a01) * wr0) * wc1; n += (val(a10) * wr1) * wc0; n += (val(a11) * wr1) * wc1;
    double d = (ok(a00) * wr0) * wc0; d += (ok(a01) * wr0) * wc1; d += (ok(a10) * wr1) * wc0; d += (ok(a11) * wr1) * wc1;
    float nf = (float)n, df = (float)d;
    out[y * W + x] = df > 1e-3f ? nf / fmaxf(df, 1e-6f) : nanf("");
  }
}

__global__ void any_nan_kernel(const float* __restrict__ in, int64_t n, int* flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  int f = 0;
  for (; i < n; i += step) f |= (in[i] != in[i]);
  if (__syncthreads_or(f) && threadIdx.x == 0) *flag = 1;
}

// _bilinear_sample_coarse (algorithms/_nan_utils.py:255-281) fused with W_large*block - up
// (algorithms/_impl_topousm_fast.py:158-186): f32 coordinates, map_coordinates(order=1, 'nearest').
__global__ void __launch_bounds__(256) large_part_kernel(const float* __restrict__ block, float* __restrict__ out, int64_t h,
                                                         int64_t w, int64_t ld_in, int64_t ld_out,
                                                         const float* __restrict__ field, int64_t ch, int64_t cw,
                                                         int64_t ld_f, int64_t off_r, int64_t off_c, float sr, float sc,
                                                         float w_large) {
  int64_t x = (int64_t)blockIdx.x * 256 + threadIdx.x;
  int64_t y = (int64_t)blockIdx.y + (int64_t)blockIdx.z * 32768;
  if (x >= w || y >= h) return;
  // (arange(r0, r1, f32) + 0.5f) * f32(ch/full_h) - 0.5f, all f32
  float rr = ((float)(off_r + y) + 0.5f) * sr - 0.5f;
  float cc = ((float)(off_c + x) + 0.5f) * sc - 0.5f;
  // map_coordinates order=1, mode='nearest': coordinates are used in f64
  double dr = (double)rr, dc = (double)cc;
  auto tap = [&](double coord, int64_t n, int64_t* i0, int64_t* i1, double* t) {
    double fl = floor(coord);
    *t = coord - fl;
    int64_t a = (int64_t)fl, b = a + 1;
    *i0 = a < 0 ? 0 : (a > n - 1 ? n - 1 : a);
    *i1 = b < 0 ? 0 : (b > n - 1 ? n - 1 : b);
  };
  int64_t r0, r1, c0, c1;
  double tr, tc;
  tap(dr, ch, &r0, &r1, &tr);
  tap(dc, cw, &c0, &c1, &tc);
  double v = ((double)field[r0 * ld_f + c0] * (1.0 - tr)) * (1.0 - tc);
  v += ((double)field[r0 * ld_f + c1] * (1.0 - tr)) * tc;
  v += ((double)field[r1 * ld_f + c0] * tr) * (1.0 - tc);
  v += ((double)field[r1 * ld_f + c1] * tr) * tc;
  float up = (float)v;
  out[y * ld_out + x] = w_large * block[y * ld_in + x] - up;
}

static int run_decimate(const float* in, float* out, int64_t H, int64_t W, int64_t ld_in, int f, void* ws,
                        size_t ws_bytes, cudaStream_t s) {
  if (f != 2 && f != 4 && f != 8 && f != 16) return fail(FSG_E_INVALID, "fsg_decimate: factor must be 2, 4, 8 or 16");
  if (!in || !out || H < 1 || W < 1 || ld_in < W) return fail(FSG_E_INVALID, "fsg_decimate: bad argument");
  int64_t h = (H + f - 1) / f, w = (W + f - 1) / f;
  double sigma = (double)(h < w ? h : w) / 64.0;
  if (sigma < 1.0) sigma = 1.0;
  int radius = (int)(4.0 * sigma + 0.5);
  size_t plane = align_up((size_t)h * w * 4, 256);
  size_t need = 2 * plane + align_up((size_t)(radius + 1) * 8, 256) + 256;
  if (!ws || ws_bytes < need) return fail(FSG_E_WORKSPACE, "fsg_decimate: workspace of %zu bytes needed", need);
  unsigned char* base = (unsigned char*)ws;
  float* tv = (float*)base;
  float* tw = (float*)(base + plane);
  double* taps = (double*)(base + 2 * plane);
  int* flags = (int*)(base + 2 * plane + align_up((size_t)(radius + 1) * 8, 256));
  zero_flags_kernel<<<1, 32, 0, s>>>(flags);
  FSG_LAUNCH_OK();
  int rc;
  float* grids1[1] = {out};
  if ((rc = launch_pyramid(in, H, W, ld_in, 1, &f, grids1, flags, s))) return rc;
  Grid g{out, h, w, w};
  if ((rc = launch_gauss_taps(sigma, radius, taps, s))) return rc;
  if ((rc = launch_gauss_axis0(g, taps, radius, tv, tw, flags, 0, h, s))) return rc;
  return launch_gauss_axis1(tv, tw, h, w, taps, radius, COMBINE_VOIDFILL, out, flags, flags + 4, s);
}

}  // namespace fsg

extern "C" {

size_t fsg_topousm_fast_workspace_bytes(int64_t H, int64_t W, const int32_t* radii_host, int n_radii, double pixel_size) {
  fsg::HostPlan plan;
  if (fsg::make_plan(H, W, radii_host, n_radii, pixel_size, &plan)) return 0;
  return plan.total;
}

int fsg_topousm_fast_roi(const float* dem, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out,
                         const int32_t* radii_host, const float* weights_host, int n_radii, double pixel_size,
                         double norm_scale, const fsg_encode* enc, void* workspace, size_t workspace_bytes,
                         int64_t roi_row0, int64_t roi_rows, int64_t roi_col0, int64_t roi_cols, void* stream) {
  if (!radii_host || !weights_host) return fsg::fail(FSG_E_INVALID, "fsg_topousm_fast_roi: radii/weights are NULL");
  return fsg::run_topousm(dem, out, H, W, ld_in, ld_out, radii_host, weights_host, n_radii, pixel_size, norm_scale,
                          enc, workspace, workspace_bytes, (cudaStream_t)stream, roi_row0, roi_rows, roi_col0, roi_cols);
}

int fsg_topousm_fast(const float* dem, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out,
                     const int32_t* radii_host, const float* weights_host, int n_radii, double pixel_size,
                     double norm_scale, const fsg_encode* enc, void* workspace, size_t workspace_bytes, void* stream) {
  if (!radii_host || !weights_host) return fsg::fail(FSG_E_INVALID, "fsg_topousm_fast: radii/weights are NULL");
  return fsg::run_topousm(dem, out, H, W, ld_in, ld_out, radii_host, weights_host, n_radii, pixel_size, norm_scale,
                          enc, workspace, workspace_bytes, (cudaStream_t)stream);
}


size_t fsg_decimate_workspace_bytes(int64_t H, int64_t W, int factor) {
  if (factor < 2 || H < 1 || W < 1) return 0;
  int64_t h = (H + factor - 1) / factor, w = (W + factor - 1) / factor;
  double sigma = (double)(h < w ? h : w) / 64.0;
  if (sigma < 1.0) sigma = 1.0;
  int radius = (int)(4.0 * sigma + 0.5);
  return 2 * fsg::align_up((size_t)h * w * 4, 256) + fsg::align_up((size_t)(radius + 1) * 8, 256) + 256;
}

int fsg_decimate(const float* in, float* out, int64_t H, int64_t W, int64_t ld_in, int factor, void* workspace,
                 size_t workspace_bytes, void* stream) {
  return fsg::run_decimate(in, out, H, W, ld_in, factor, workspace, workspace_bytes, (cudaStream_t)stream);
}

/* workspace: >= 256 bytes (NaN flag) */
int fsg_upsample(const float* in, float* out, int64_t h, int64_t w, int64_t H, int64_t W, void* workspace,
                 size_t workspace_bytes, void* stream) {
  using namespace fsg;
  if (!in || !out || h < 1 || w < 1 || H < 1 || W < 1) return fail(FSG_E_INVALID, "fsg_upsample: bad argument");
  if (!workspace || workspace_bytes < 256) return fail(FSG_E_WORKSPACE, "fsg_upsample: 256-byte workspace needed");
  cudaStream_t s = (cudaStream_t)stream;
  int* flag = (int*)workspace;
  zero_flags_kernel<<<1, 32, 0, s>>>(flag);
  FSG_LAUNCH_OK();
  int64_t n = h * w;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  any_nan_kernel<<<blocks, 256, 0, s>>>(in, n, flag);
  FSG_LAUNCH_OK();
  double rs = H > 1 ? (double)(h - 1) / (double)(H - 1) : 1.0;
  double cs = W > 1 ? (double)(w - 1) / (double)(W - 1) : 1.0;
  dim3 grid((unsigned)((W + 255) / 256), (unsigned)(H < 32768 ? H : 32768), (unsigned)((H + 32767) / 32768));
  upsample_kernel<<<grid, 256, 0, s>>>(in, out, h, w, H, W, rs, cs, flag);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_topousm_large_part(const float* block, float* out, int64_t h, int64_t w, int64_t ld_in, int64_t ld_out,
                           const float* field, int64_t ch, int64_t cw, int64_t ld_field, int64_t off_r, int64_t off_c,
                           int64_t full_h, int64_t full_w, double w_large, void* stream) {
  using namespace fsg;
  if (!block || !out || !field || h < 1 || w < 1 || ch < 1 || cw < 1 || full_h < 1 || full_w < 1)
    return fail(FSG_E_INVALID, "fsg_topousm_large_part: bad argument");
  float sr = (float)((double)ch / (double)full_h), sc = (float)((double)cw / (double)full_w);
  dim3 grid((unsigned)((w + 255) / 256), (unsigned)(h < 32768 ? h : 32768), (unsigned)((h + 32767) / 32768));
  large_part_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(block, out, h, w, ld_in, ld_out, field, ch, cw, ld_field,
                                                           off_r, off_c, sr, sc, (float)w_large);
  FSG_LAUNCH_OK();
  return FSG_OK;
}


/* ---- row-band shard entry points ---------------------------------------------------------- */
int fsg_topousm_plan(const int32_t* radii_host, int n_radii, double pixel_size, int32_t* kind_host,
                     int32_t* factor_host, int32_t* size_host, int32_t* fused_halo_host) {
  fsg::HostPlan plan;
  if (!radii_host || !kind_host || !factor_host || !size_host || !fused_halo_host)
    return fsg::fail(FSG_E_INVALID, "fsg_topousm_plan: NULL argument");
  int rc = fsg::make_plan(1 << 20, 1 << 20, radii_host, n_radii, pixel_size, &plan);  // kinds do not depend on H, W
  if (rc) return rc;
  for (int i = 0; i < n_radii; ++i) {
    kind_host[i] = plan.terms[i].kind;
    factor_host[i] = plan.terms[i].ds;
    size_host[i] = plan.terms[i].size;
  }
  *fused_halo_host = plan.fused_R;
  return FSG_OK;
}

int fsg_pyramid_band(const float* dem, int64_t rows, int64_t W, int64_t ld, const int32_t* factors_host, int n_levels,
                     float* const* grids_host, int32_t* flags_dev, void* stream) {
  using namespace fsg;
  if (!dem || !factors_host || !grids_host || !flags_dev || rows < 1 || W < 1 || ld < W)
    return fail(FSG_E_INVALID, "fsg_pyramid_band: bad argument");
  if (n_levels < 1 || n_levels > MAX_LEVELS) return fail(FSG_E_INVALID, "fsg_pyramid_band: 1..4 levels");
  int factors[MAX_LEVELS];
  for (int k = 0; k < n_levels; ++k) {
    factors[k] = factors_host[k];
    if (factors[k] != 2 && factors[k] != 4 && factors[k] != 8 && factors[k] != 16)
      return fail(FSG_E_INVALID, "fsg_pyramid_band: factor must be 2, 4, 8 or 16");
    if (!grids_host[k]) return fail(FSG_E_INVALID, "fsg_pyramid_band: NULL grid");
  }
  cudaStream_t s = (cudaStream_t)stream;
  zero_flags_kernel<<<1, 32, 0, s>>>((int*)flags_dev);
  FSG_LAUNCH_OK();
  return launch_pyramid(dem, rows, W, ld, n_levels, factors, grids_host, (int*)flags_dev, s);
}

size_t fsg_grid_mean_band_workspace_bytes(int64_t out_rows, int64_t gw) {
  if (out_rows < 1 || gw < 1) return 256;
  return 2 * fsg::align_up((size_t)out_rows * gw * 4, 256) + 256;
}

int fsg_grid_mean_band(const float* src, int64_t src_row0, int64_t src_rows, int64_t gh, int64_t gw, int64_t ld,
                       int size, int64_t out_row0, int64_t out_rows, float* out, void* workspace,
                       size_t workspace_bytes, void* stream) {
  using namespace fsg;
  if (!src || !out || gh < 1 || gw < 1 || ld < gw || src_rows < 1) return fail(FSG_E_INVALID, "fsg_grid_mean_band: bad argument");
  if (size < 0 || (size > 0 && size % 2 == 0)) return fail(FSG_E_INVALID, "fsg_grid_mean_band: size must be 0 (gaussian) or odd");
  if (out_row0 < 0 || out_rows < 0 || out_row0 + out_rows > gh) return fail(FSG_E_INVALID, "fsg_grid_mean_band: output rows outside the grid");
  if (out_rows == 0) return FSG_OK;
  // rows the vertical pass touches after mirroring / clamping at the global edges
  int64_t reach = size == 0 ? 4 : size / 2;
  int64_t lo = out_row0 - reach, hi = out_row0 + out_rows - 1 + reach;
  int64_t need_lo = lo < 0 ? 0 : lo, need_hi = hi > gh - 1 ? gh - 1 : hi;
  if (size > 0) {  // mirrored rows may reach further inside
    if (lo < 0 && -lo - 1 > need_hi) need_hi = (-lo - 1 > gh - 1) ? gh - 1 : -lo - 1;
    if (hi > gh - 1 && 2 * gh - 1 - hi < need_lo) need_lo = (2 * gh - 1 - hi < 0) ? 0 : 2 * gh - 1 - hi;
  }
  if (src_row0 > need_lo || src_row0 + src_rows - 1 < need_hi)
    return fail(FSG_E_INVALID, "fsg_grid_mean_band: source rows [%lld,%lld) do not cover [%lld,%lld]",
                (long long)src_row0, (long long)(src_row0 + src_rows), (long long)need_lo, (long long)need_hi);
  size_t plane = align_up((size_t)out_rows * gw * 4, 256);
  if (!workspace || workspace_bytes < 2 * plane + 256) return fail(FSG_E_WORKSPACE, "fsg_grid_mean_band: workspace too small");
  unsigned char* base = (unsigned char*)workspace;
  float* tv = (float*)base;
  float* tw = (float*)(base + plane);
  double* taps = (double*)(base + 2 * plane);
  Grid g{src, gh, gw, ld};
  g.row_off = src_row0;
  return mean_on_grid(g, size, out, tv, tw, taps, out_row0, out_rows, (cudaStream_t)stream);
}

int fsg_topousm_fused_band(const float* dem, int64_t dem_row0, int64_t dem_rows, int64_t H, int64_t W, int64_t ld_in,
                           void* out, int64_t out_row0, int64_t out_rows, int64_t ld_out,
                           const int32_t* radii_host, const float* weights_host, int n_radii, double pixel_size,
                           const float* const* term_grids_host, const int64_t* term_grow0_host,
                           const int64_t* term_grows_host, double norm_scale, const fsg_encode* enc, void* stream) {
  if (!radii_host || !weights_host) return fsg::fail(FSG_E_INVALID, "fsg_topousm_fused_band: radii/weights are NULL");
  if (ld_in < W || ld_out < W) return fsg::fail(FSG_E_INVALID, "fsg_topousm_fused_band: row stride smaller than width");
  return fsg::run_fused_band(dem, dem_row0, dem_rows, H, W, ld_in, out, out_row0, out_rows, ld_out, radii_host,
                             weights_host, n_radii, pixel_size, term_grids_host, term_grow0_host, term_grows_host,
                             norm_scale, nullptr, enc, nullptr, 0, (cudaStream_t)stream);
}

int64_t fsg_debug_v8_band_rows(int64_t rows, int64_t strips) { return fsg::v8_band_rows(rows, strips); }

#ifdef FSG_V8_TIMERS
int fsg_debug_v8_timers(unsigned long long* out48) {
  return cudaMemcpyFromSymbol(out48, fsg::v8_timers, sizeof(unsigned long long) * 48) == cudaSuccess ? 0 : -1;
}
#endif

size_t fsg_topousm_fused_band_workspace_bytes(int64_t out_rows, int64_t W) { return fsg::v8_flag_bytes(out_rows, W); }

int fsg_topousm_fused_band_ws(const float* dem, int64_t dem_row0, int64_t dem_rows, int64_t H, int64_t W, int64_t ld_in,
                              void* out, int64_t out_row0, int64_t out_rows, int64_t ld_out,
                              const int32_t* radii_host, const float* weights_host, int n_radii, double pixel_size,
                              const float* const* term_grids_host, const int64_t* term_grow0_host,
                              const int64_t* term_grows_host, double norm_scale, const float* norm_scale_dev,
                              const fsg_encode* enc, void* workspace, size_t workspace_bytes, void* stream) {
  if (!radii_host || !weights_host) return fsg::fail(FSG_E_INVALID, "fsg_topousm_fused_band: radii/weights are NULL");
  if (ld_in < W || ld_out < W) return fsg::fail(FSG_E_INVALID, "fsg_topousm_fused_band: row stride smaller than width");
  return fsg::run_fused_band(dem, dem_row0, dem_rows, H, W, ld_in, out, out_row0, out_rows, ld_out, radii_host,
                             weights_host, n_radii, pixel_size, term_grids_host, term_grow0_host, term_grows_host,
                             norm_scale, norm_scale_dev, enc, workspace, workspace_bytes, (cudaStream_t)stream);
}


/* Enclosed-void fill of a WHOLE decimated grid, in place (algorithms/_nan_utils.py:655-667): a no-op
 * unless the grid holds a NaN cell.  workspace: fsg_decimate_workspace_bytes(gh*factor, gw*factor, factor)
 * or simply 2*gh*gw*4 + 8*(4*sigma+2) + 1024 bytes. */
int fsg_grid_void_fill(float* grid, int64_t gh, int64_t gw, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace fsg;
  if (!grid || gh < 1 || gw < 1) return fail(FSG_E_INVALID, "fsg_grid_void_fill: bad argument");
  double sigma = (double)(gh < gw ? gh : gw) / 64.0;
  if (sigma < 1.0) sigma = 1.0;
  int radius = (int)(4.0 * sigma + 0.5);
  size_t plane = align_up((size_t)gh * gw * 4, 256);
  size_t need = 2 * plane + align_up((size_t)(radius + 1) * 8, 256) + 512;
  if (!workspace || workspace_bytes < need) return fail(FSG_E_WORKSPACE, "fsg_grid_void_fill: workspace of %zu bytes needed", need);
  unsigned char* base = (unsigned char*)workspace;
  float* tv = (float*)base;
  float* tw = (float*)(base + plane);
  double* taps = (double*)(base + 2 * plane);
  int* flags = (int*)(base + 2 * plane + align_up((size_t)(radius + 1) * 8, 256));
  cudaStream_t s = (cudaStream_t)stream;
  zero_flags_kernel<<<1, 32, 0, s>>>(flags);
  FSG_LAUNCH_OK();
  int64_t n = gh * gw;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  any_nan_kernel<<<blocks, 256, 0, s>>>(grid, n, flags);
  FSG_LAUNCH_OK();
  Grid g{grid, gh, gw, gw};
  int rc;
  if ((rc = launch_gauss_taps(sigma, radius, taps, s))) return rc;
  if ((rc = launch_gauss_axis0(g, taps, radius, tv, tw, flags, 0, gh, s))) return rc;
  return launch_gauss_axis1(tv, tw, gh, gw, taps, radius, COMBINE_VOIDFILL, grid, flags, flags + 4, s);
}

}  // extern "C"
