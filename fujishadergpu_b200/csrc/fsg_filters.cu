// Generic NaN-aware separable box / Gaussian passes (see fsg_filters.cuh for the contract).
#include "fsg_filters.cuh"

namespace fsg {

constexpr int A0_CHUNK = 128;  // rows per thread in the axis-0 running-sum pass
constexpr int A1_CW = 256;     // output columns per CTA in the axis-1 pass
constexpr int VT_W = 256;      // columns per tile of the void-fill need map

// ---------------------------------------------------------------- box, axis 0 (running sums)
__global__ void __launch_bounds__(128) box_axis0_kernel(Grid g, int size, float* __restrict__ tv, float* __restrict__ tw,
                                                        int64_t oy0, int64_t oh, int chunk) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= g.w) return;
  int64_t y0 = oy0 + (int64_t)blockIdx.y * chunk;
  int64_t y1 = y0 + chunk < oy0 + oh ? y0 + chunk : oy0 + oh;
  const int lo = size / 2, hi = size - 1 - lo;
  const double n = (double)size, inv = 1.0 / n;
  const float nf = (float)size;
  double s = 0.0;
  int cnt = 0;
  const bool interior = y0 - lo >= 0 && y1 + hi < g.h;
  if (interior) {
    const float* pk = g.src + (y0 - lo - g.row_off) * g.ld + x;
    double s1 = 0.0;
    for (int k = 0; k < size; ++k) {      // exact sums: two accumulators, association free
      float v = *pk;
      bool ok = v == v;
      if (k & 1) s1 += ok ? (double)v : 0.0;
      else s += ok ? (double)v : 0.0;
      cnt += ok;
      pk += g.ld;
    }
    s += s1;
  } else {
    for (int k = -lo; k <= hi; ++k) {
      float v = g.src[(reflect_index(y0 + k, g.h) - g.row_off) * g.ld + x];
      bool ok = v == v;
      s += ok ? (double)v : 0.0;
      cnt += ok;
    }
  }
  if (interior) {
    // interior chunk: no mirroring, plain row pointers
    const float* pin = g.src + (y0 + hi + 1 - g.row_off) * g.ld + x;
    const float* pout = g.src + (y0 - lo - g.row_off) * g.ld + x;
    float* pv = tv + (y0 - oy0) * g.w + x;
    float* pw = tw + (y0 - oy0) * g.w + x;
    int64_t y = y0;
    for (; y + 4 <= y1; y += 4) {     // four rows of loads in flight
      float vi[4], vo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { vi[k] = pin[k * g.ld]; vo[k] = pout[k * g.ld]; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pv[k * g.w] = (float)div_by_count(s, n, inv);
        pw[k * g.w] = (float)cnt / nf;
        bool oin = vi[k] == vi[k], oout = vo[k] == vo[k];
        s += (oin ? (double)vi[k] : 0.0) - (oout ? (double)vo[k] : 0.0);
        cnt += (int)oin - (int)oout;
      }
      pin += 4 * g.ld; pout += 4 * g.ld; pv += 4 * g.w; pw += 4 * g.w;
    }
    for (; y < y1; ++y) {
      *pv = (float)div_by_count(s, n, inv);
      *pw = (float)cnt / nf;
      float vin = *pin, vout = *pout;
      bool oin = vin == vin, oout = vout == vout;
      s += (oin ? (double)vin : 0.0) - (oout ? (double)vout : 0.0);
      cnt += (int)oin - (int)oout;
      pin += g.ld; pout += g.ld; pv += g.w; pw += g.w;
    }
    return;
  }
  for (int64_t y = y0; y < y1; ++y) {
    tv[(y - oy0) * g.w + x] = (float)div_by_count(s, n, inv);
    tw[(y - oy0) * g.w + x] = (float)cnt / nf;
    float vin = g.src[(reflect_index(y + hi + 1, g.h) - g.row_off) * g.ld + x];
    float vout = g.src[(reflect_index(y - lo, g.h) - g.row_off) * g.ld + x];
    bool oin = vin == vin, oout = vout == vout;
    s += (oin ? (double)vin : 0.0) - (oout ? (double)vout : 0.0);
    cnt += (int)oin - (int)oout;
  }
}

// ---------------------------------------------------------------- box, axis 1 (smem tiles)
template <int RT>
__global__ void __launch_bounds__(256) box_axis1_kernel(const float* __restrict__ tv, const float* __restrict__ tw,
                                                        int64_t h, int64_t w, int size, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int span = A1_CW + size - 1;
  const int P = span | 1;
  float* sv_p = sm;
  float* sw_p = sm + (size_t)RT * P;
  float* so_p = sw_p + (size_t)RT * P;  // RT x (A1_CW+1)
  const int64_t x0 = (int64_t)blockIdx.x * A1_CW;
  const int64_t y0 = (int64_t)blockIdx.y * RT;
  const int lo = size / 2;
  const int tid = threadIdx.x;
  const bool interior_x = x0 - lo >= 0 && x0 - lo + span <= w;
  for (int r = tid / 32; r < RT; r += 8) {       // one warp per tile row, lanes over the columns
    const int64_t gy = y0 + r;
    const bool row_ok = gy < h;
    const float* rv = tv + gy * w;
    const float* rw = tw + gy * w;
    if (row_ok && interior_x) {
      const float* pa = rv + x0 - lo;
      const float* pb = rw + x0 - lo;
      int c = tid & 31;
      for (; c + 96 < span; c += 128) {   // eight loads in flight per lane
        float a0 = pa[c], a1 = pa[c + 32], a2 = pa[c + 64], a3 = pa[c + 96];
        float b0 = pb[c], b1 = pb[c + 32], b2 = pb[c + 64], b3 = pb[c + 96];
        sv_p[r * P + c] = a0; sv_p[r * P + c + 32] = a1; sv_p[r * P + c + 64] = a2; sv_p[r * P + c + 96] = a3;
        sw_p[r * P + c] = b0; sw_p[r * P + c + 32] = b1; sw_p[r * P + c + 64] = b2; sw_p[r * P + c + 96] = b3;
      }
      for (; c < span; c += 32) { sv_p[r * P + c] = pa[c]; sw_p[r * P + c] = pb[c]; }
    } else {
      for (int c = tid & 31; c < span; c += 32) {
        float a = 0.f, b = 0.f;
        if (row_ok) {
          int64_t gx = reflect_index(x0 - lo + c, w);
          a = rv[gx];
          b = rw[gx];
        }
        sv_p[r * P + c] = a;
        sw_p[r * P + c] = b;
      }
    }
  }
  __syncthreads();
  constexpr int NSEG = 256 / RT;
  constexpr int SEGLEN = A1_CW / NSEG;
  const int r = tid % RT, gseg = tid / RT;
  const int j0 = gseg * SEGLEN;
  if (y0 + r < h && x0 + j0 < w) {
    const double n = (double)size, inv = 1.0 / n;
    const float* pv = sv_p + r * P + j0;
    const float* pw = sw_p + r * P + j0;
    double sv = 0.0, sw = 0.0;
    for (int k = 0; k < size; ++k) {
      sv += (double)pv[k];
      sw += (double)pw[k];
    }
    for (int jj = 0; jj < SEGLEN; ++jj) {
      float num = (float)div_by_count(sv, n, inv);
      float den = (float)div_by_count(sw, n, inv);
      so_p[r * (A1_CW + 1) + j0 + jj] = den > 0.f ? num / den : 0.f;
      sv += (double)pv[jj + size] - (double)pv[jj];
      sw += (double)pw[jj + size] - (double)pw[jj];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < RT * A1_CW; idx += 256) {
    int rr = idx / A1_CW, c = idx - rr * A1_CW;
    if (y0 + rr < h && x0 + c < w) out[(y0 + rr) * w + x0 + c] = so_p[rr * (A1_CW + 1) + c];
  }
}


// ---------------------------------------------------------------- box, both axes in one pass
// The two passes above move the grid five times (read, two f32 planes out, two planes in, write).  This kernel keeps
// the intermediate planes of a 16-row batch in shared memory: a CTA walks a strip of CW output columns down a band
// of rows; phase V, one thread per input column (CW + size - 1 of them), advances the exact f64 running sum / valid
// count of the column and leaves the f32-rounded axis-0 means of the batch in shared memory; phase H, thread =
// (row, segment of CW / 16 outputs), slides the axis-1 window over them.  Same values as the two passes bit for bit
// (the window sums are exact; each axis rounds once to f32; NaN-aware combine as in box_axis1_kernel).  A batch
// whose columns are all NaN-free skips the weight plane (every weight is 1.0f, their mean is 1.0f).
constexpr int B2_RB = 16;
template <int CW, int MAXSIZE>
struct Box2Geom {
  static constexpr int NT = CW + MAXSIZE - 1;          // threads = input columns
  static constexpr int P = NT + 1;                     // plane pitch (odd: rows fall on different banks)
  static constexpr int SEG = CW / 16;
  static constexpr int PO = CW + 1;
  static constexpr size_t BYTES = (size_t)(2 * B2_RB * P + B2_RB * PO) * 4;
};

template <int CW, int MAXSIZE, int MINB>
__global__ void __launch_bounds__(CW + MAXSIZE - 1, MINB)
box_mean2d_kernel(Grid g, int size, float* __restrict__ out, int64_t oy0, int64_t oh, int band_rows) {
  using G = Box2Geom<CW, MAXSIZE>;
  extern __shared__ float sm[];
  float* tvs = sm;
  float* tws = sm + B2_RB * G::P;
  float* so = tws + B2_RB * G::P;
  const int tid = threadIdx.x;
  const int lo = size / 2, hi = size - 1 - lo;
  const int64_t x0 = (int64_t)blockIdx.x * CW;
  const int64_t y0 = oy0 + (int64_t)blockIdx.y * band_rows;
  const int64_t y1 = y0 + band_rows < oy0 + oh ? y0 + band_rows : oy0 + oh;
  const double n = (double)size, inv = 1.0 / n;
  const float nf = (float)size;
  // phase V role
  const bool v_on = tid < CW + size - 1;
  const float* col = g.src + reflect_near(x0 - lo + tid, g.w) - g.row_off * g.ld;
  double s = 0.0;
  int cnt = 0;
  if (v_on) {
    double s1 = 0.0;
    if (y0 - lo >= 0 && y0 + hi < g.h) {
      const float* pk = col + (y0 - lo) * g.ld;
      for (int k = 0; k < size; ++k) {
        const float v = pk[(int64_t)k * g.ld];
        const bool ok = v == v;
        if (k & 1) s1 += ok ? (double)v : 0.0;
        else s += ok ? (double)v : 0.0;
        cnt += ok;
      }
    } else {
      for (int k = -lo; k <= hi; ++k) {
        const float v = col[reflect_near(y0 + k, g.h) * g.ld];
        const bool ok = v == v;
        if (k & 1) s1 += ok ? (double)v : 0.0;
        else s += ok ? (double)v : 0.0;
        cnt += ok;
      }
    }
    s += s1;
  }
  // phase H role: warp = two segments 4 apart (SEG = 12: 48 columns = 16 banks), lanes 0-15 / 16-31 = the 16 rows
  const bool h_on = tid < 256;
  const int hr = tid & 15;
  const int hseg = ((tid >> 5) & 3) + 8 * (tid >> 7) + 4 * ((tid >> 4) & 1);
  const int j0 = hseg * G::SEG;

  for (int64_t yb = y0; yb < y1; yb += B2_RB) {
    const int nb = y1 - yb < B2_RB ? (int)(y1 - yb) : B2_RB;
    int full = 1;
    if (v_on) {
      float vin[B2_RB], vout[B2_RB];
      if (yb - lo >= 0 && yb + B2_RB + hi < g.h) {
        const float* pin = col + (yb + hi + 1) * g.ld;
        const float* pout = col + (yb - lo) * g.ld;
#pragma unroll
        for (int i = 0; i < B2_RB; ++i) {
          vin[i] = i < nb ? pin[(int64_t)i * g.ld] : 0.f;
          vout[i] = i < nb ? pout[(int64_t)i * g.ld] : 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < B2_RB; ++i) {
          vin[i] = i < nb ? col[reflect_near(yb + i + hi + 1, g.h) * g.ld] : 0.f;
          vout[i] = i < nb ? col[reflect_near(yb + i - lo, g.h) * g.ld] : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < B2_RB; ++i) {
        tvs[i * G::P + tid] = (float)div_by_count(s, n, inv);
        tws[i * G::P + tid] = cnt == size ? 1.f : (float)cnt / nf;
        if (i < nb) full &= (int)(cnt == size);
        const bool oin = vin[i] == vin[i], oout = vout[i] == vout[i];
        s += (oin ? (double)vin[i] : 0.0) - (oout ? (double)vout[i] : 0.0);
        cnt += (int)oin - (int)oout;
      }
    }
    const int dense = __syncthreads_and(full);
    if (h_on && hr < nb) {
      const float* pv = tvs + hr * G::P + j0;
      const float* pw = tws + hr * G::P + j0;
      float* po = so + hr * G::PO + j0;
      double sv = 0.0, sv1 = 0.0, sv2 = 0.0, sv3 = 0.0;   // (exact sums: any association)
      int k = 0;
      for (; k + 3 < size; k += 4) {
        sv += (double)pv[k];
        sv1 += (double)pv[k + 1];
        sv2 += (double)pv[k + 2];
        sv3 += (double)pv[k + 3];
      }
      for (; k < size; ++k) sv += (double)pv[k];
      sv = (sv + sv1) + (sv2 + sv3);
      if (dense) {
#pragma unroll 4
        for (int j = 0; j < G::SEG; ++j) {
          po[j] = (float)div_by_count(sv, n, inv);
          sv += (double)pv[j + size] - (double)pv[j];
        }
      } else {
        double sw = 0.0;
        for (int q = 0; q < size; ++q) sw += (double)pw[q];
#pragma unroll 2
        for (int j = 0; j < G::SEG; ++j) {
          const float num = (float)div_by_count(sv, n, inv);
          const float den = (float)div_by_count(sw, n, inv);
          po[j] = den > 0.f ? num / den : 0.f;
          sv += (double)pv[j + size] - (double)pv[j];
          sw += (double)pw[j + size] - (double)pw[j];
        }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < nb * CW; idx += G::NT) {
      const int rr = idx / CW, c = idx - rr * CW;
      if (x0 + c < g.w) out[(yb - oy0 + rr) * g.w + x0 + c] = so[rr * G::PO + c];
    }
  }
}

// ---------------------------------------------------------------- gaussian passes (direct)
__global__ void gauss_taps_kernel(double sigma, int radius, double* w) {
  // one block; the normalising sum is accumulated by thread 0 in index order (deterministic)
  __shared__ double tot;
  double s2 = sigma * sigma;
  for (int x = threadIdx.x; x <= radius; x += blockDim.x) w[x] = exp(-0.5 / s2 * (double)x * (double)x);
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int x = radius; x >= 1; --x) t += w[x];   // phi(-r..-1)
    t += w[0];
    for (int x = 1; x <= radius; ++x) t += w[x];
    tot = t;
  }
  __syncthreads();
  for (int x = threadIdx.x; x <= radius; x += blockDim.x) w[x] = w[x] / tot;
}

// ---- enclosed-void fill: where the axis-0 pass is needed at all ----
// The fill only rewrites void cells, and a void cell (y, x) reads the axis-0 plane at (y, x - radius .. x + radius):
// the axis-0 pass (2 * radius + 1 taps in f64 per cell: 2049 on the level-4 grid of a 65536^2 raster) is evaluated only
// on the (row, 256-column tile) pairs within reach of a void cell.  3 % NoData: 834 -> 130 ms for the main pass.
__global__ void __launch_bounds__(256) void_tiles_kernel(const float* __restrict__ grid, int64_t h, int64_t w, int nt,
                                                         const int* run_flag, unsigned char* __restrict__ vt,
                                                         int* __restrict__ vlist, int* __restrict__ vcount) {
  if (run_flag && *run_flag == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t item0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t item = item0; item < h * nt; item += nwarps) {
    const int64_t y = item / nt;
    const int t = (int)(item - y * nt);
    const float* row = grid + y * w;
    bool any = false;
    for (int c = lane; c < VT_W; c += 32) {
      const int64_t x = (int64_t)t * VT_W + c;
      if (x < w) { const float v = row[x]; any |= !(v == v); }
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {
      vt[item] = any ? 1 : 0;
      if (any) vlist[atomicAdd(vcount, 1)] = (int)item;
    }
  }
}

__global__ void __launch_bounds__(256) need_tiles_kernel(const unsigned char* __restrict__ vt, int64_t h, int nt, int reach,
                                                         const int* run_flag, int* __restrict__ nlist,
                                                         int* __restrict__ ncount) {
  if (run_flag && *run_flag == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t total = h * nt;
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; base < total;
       base += (int64_t)gridDim.x * blockDim.x) {   // (whole warps: the append below is aggregated per warp)
    const int64_t item = base + lane;
    bool any = false;
    if (item < total) {
      const int64_t y = item / nt;
      const int t = (int)(item - y * nt);
      int lo = t - reach, hi = t + reach;
      lo = lo < 0 ? 0 : lo;
      hi = hi > nt - 1 ? nt - 1 : hi;
      for (int k = lo; k <= hi; ++k) any |= vt[y * nt + k] != 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, any);
    if (m) {   // consecutive needed items stay consecutive in the list (rows / tiles next to each other share loads)
      int pos = 0;
      if (lane == 0) pos = atomicAdd(ncount, __popc(m));
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if (any) nlist[pos + __popc(m & ((1u << lane) - 1u))] = (int)item;
    }
  }
}

__global__ void __launch_bounds__(256) gauss_axis0_kernel(Grid g, const double* __restrict__ taps, int radius,
                                                          float* __restrict__ tv, float* __restrict__ tw,
                                                          const int* run_flag, int64_t oy0, int64_t oh,
                                                          const int* __restrict__ list, const int* __restrict__ list_n,
                                                          int nt) {
  if (run_flag && *run_flag == 0) return;
  // with an item list: a 1-D grid walks the (row, tile) items that are needed (a launch the flag cancels costs ~1 200
  // empty CTAs); else rows strided per column tile
  const bool lin = list != nullptr;
  const int64_t first = lin ? blockIdx.x : blockIdx.y, step = lin ? gridDim.x : gridDim.y;
  const int64_t count = lin ? *list_n : oh;
  for (int64_t idx = first; idx < count; idx += step) {
  const int64_t it = lin ? list[idx] : idx;
  const int64_t yl = lin ? it / nt : it;
  const int64_t x = (lin ? it - yl * nt : (int64_t)blockIdx.x) * blockDim.x + threadIdx.x;   // (blockDim.x == VT_W)
  if (x >= g.w) continue;
  const int64_t y = oy0 + yl;
  double sv, sw;
  if (y - radius >= 0 && y + radius < g.h) {
    // interior rows: no clamping, both ends of a tap pair walk with a constant stride and eight pairs of loads are in
    // flight before the (strictly ordered) accumulation uses them
    const float* pc = g.src + (y - g.row_off) * g.ld + x;
    const float v0 = *pc;
    const bool g0 = v0 == v0;
    sv = (g0 ? (double)v0 : 0.0) * taps[0];
    sw = (g0 ? 1.0 : 0.0) * taps[0];
    const float* pa = pc - (int64_t)radius * g.ld;   // row y - j
    const float* pb = pc + (int64_t)radius * g.ld;   // row y + j
    int j = radius;
    for (; j >= 8; j -= 8) {
      float a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { a[u] = pa[(int64_t)u * g.ld]; b[u] = pb[-(int64_t)u * g.ld]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ga = a[u] == a[u], gb = b[u] == b[u];
        const double va = ga ? (double)a[u] : 0.0, vb = gb ? (double)b[u] : 0.0;
        const double oc = (ga ? 1.0 : 0.0) + (gb ? 1.0 : 0.0);
        const double t = taps[j - u];
        sv += (va + vb) * t;
        sw += oc * t;
      }
      pa += 8 * g.ld;
      pb -= 8 * g.ld;
    }
    for (; j >= 1; --j) {
      const float fa = *pa, fb = *pb;
      const bool ga = fa == fa, gb = fb == fb;
      sv += ((ga ? (double)fa : 0.0) + (gb ? (double)fb : 0.0)) * taps[j];
      sw += ((ga ? 1.0 : 0.0) + (gb ? 1.0 : 0.0)) * taps[j];
      pa += g.ld;
      pb -= g.ld;
    }
  } else {
    auto at = [&](int64_t yy, double* ok) {
      float v = g.src[(clamp_index(yy, g.h) - g.row_off) * g.ld + x];
      bool good = v == v;
      *ok = good ? 1.0 : 0.0;
      return good ? (double)v : 0.0;
    };
    double o0;
    double v0 = at(y, &o0);
    sv = v0 * taps[0];
    sw = o0 * taps[0];
    for (int j = radius; j >= 1; --j) {
      double oa, ob;
      double va = at(y - j, &oa), vb = at(y + j, &ob);
      sv += (va + vb) * taps[j];
      sw += (oa + ob) * taps[j];
    }
  }
  tv[yl * g.w + x] = (float)sv;
  tw[yl * g.w + x] = (float)sw;
  }
}

__global__ void __launch_bounds__(256) gauss_axis1_kernel(const float* __restrict__ tv, const float* __restrict__ tw,
                                                          int64_t h, int64_t w, const double* __restrict__ taps,
                                                          int radius, int combine, float* out, const int* run_flag,
                                                          int* still_nan, const int* __restrict__ list,
                                                          const int* __restrict__ list_n, int nt) {
  if (run_flag && *run_flag == 0) return;
  const bool lin = list != nullptr;   // the (row, tile) items that hold a void cell, 1-D grid (see gauss_axis0_kernel)
  const int64_t first = lin ? blockIdx.x : blockIdx.y, step = lin ? gridDim.x : gridDim.y;
  const int64_t count = lin ? *list_n : h;
  for (int64_t idx = first; idx < count; idx += step) {
  const int64_t it = lin ? list[idx] : idx;
  const int64_t y = lin ? it / nt : it;
  const int64_t x = (lin ? it - y * nt : (int64_t)blockIdx.x) * blockDim.x + threadIdx.x;
  if (x >= w) continue;
  if (combine == COMBINE_VOIDFILL) {
    float cur = out[y * w + x];
    if (cur == cur) continue;  // only void cells are candidates
  }
  const float* rv = tv + y * w;
  const float* rw = tw + y * w;
  double sv = (double)rv[x] * taps[0], sw = (double)rw[x] * taps[0];
  if (x - radius >= 0 && x + radius < w) {
    int j = radius;
    for (; j >= 4; j -= 4) {   // (eight pairs of loads in flight; the accumulation order is unchanged)
      float av[4], bv[4], aw[4], bw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        av[u] = rv[x - j + u]; bv[u] = rv[x + j - u];
        aw[u] = rw[x - j + u]; bw[u] = rw[x + j - u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double t = taps[j - u];
        sv += ((double)av[u] + (double)bv[u]) * t;
        sw += ((double)aw[u] + (double)bw[u]) * t;
      }
    }
    for (; j >= 1; --j) {
      sv += ((double)rv[x - j] + (double)rv[x + j]) * taps[j];
      sw += ((double)rw[x - j] + (double)rw[x + j]) * taps[j];
    }
  } else {
    for (int j = radius; j >= 1; --j) {
      int64_t xa = clamp_index(x - j, w), xb = clamp_index(x + j, w);
      sv += ((double)rv[xa] + (double)rv[xb]) * taps[j];
      sw += ((double)rw[xa] + (double)rw[xb]) * taps[j];
    }
  }
  float fv = (float)sv, fw = (float)sw;
  if (combine == COMBINE_MEAN) {
    out[y * w + x] = fw > 0.f ? fv / fw : 0.f;
  } else {
    if (fw > 0.5f) out[y * w + x] = fv / fmaxf(fw, 1e-6f);
    else if (still_nan) *still_nan = 1;
  }
  }
}

// ---------------------------------------------------------------- launchers
int launch_box_axis0(const Grid& g, int size, float* tv, float* tw, int64_t oy0, int64_t oh, cudaStream_t s) {
  if (oh <= 0) return FSG_OK;
  // rows per thread: 128 on big grids (the window start costs `size` loads per chunk), fewer on small grids so
  // that at least ~2 CTAs per SM exist
  const int64_t gx = (g.w + 127) / 128;
  int64_t want_y = (148 * 2 + gx - 1) / gx;
  int64_t chunk = (oh + want_y - 1) / want_y;
  if (chunk > A0_CHUNK) chunk = A0_CHUNK;
  if (chunk < 16) chunk = 16;
  dim3 grid((unsigned)gx, (unsigned)((oh + chunk - 1) / chunk));
  box_axis0_kernel<<<grid, 128, 0, s>>>(g, size, tv, tw, oy0, oh, (int)chunk);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

template <int RT>
static int launch_a1(const float* tv, const float* tw, int64_t h, int64_t w, int size, float* out, cudaStream_t s,
                     size_t smem) {
  FSG_CUDA_OK(cudaFuncSetAttribute(box_axis1_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((w + A1_CW - 1) / A1_CW), (unsigned)((h + RT - 1) / RT));
  box_axis1_kernel<RT><<<grid, 256, smem, s>>>(tv, tw, h, w, size, out);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int launch_box_axis1(const float* tv, const float* tw, int64_t h, int64_t w, int size, float* out, cudaStream_t s) {
  auto need = [&](int rt) {
    size_t P = (size_t)((A1_CW + size - 1) | 1);
    return (size_t)rt * P * 8 + (size_t)rt * (A1_CW + 1) * 4;
  };
  const size_t cap = 220 * 1024;
  if (need(32) <= cap) return launch_a1<32>(tv, tw, h, w, size, out, s, need(32));
  if (need(8) <= cap) return launch_a1<8>(tv, tw, h, w, size, out, s, need(8));
  if (need(1) <= cap) return launch_a1<1>(tv, tw, h, w, size, out, s, need(1));
  return fail(FSG_E_UNSUPPORTED, "box filter of %d taps exceeds the shared-memory tile budget", size);
}


template <int CW, int MAXSIZE, int MINB>
static int launch_b2(const Grid& g, int size, float* out, int64_t oy0, int64_t oh, cudaStream_t s) {
  using G = Box2Geom<CW, MAXSIZE>;
  const int64_t strips = (g.w + CW - 1) / CW;
  // bands: enough CTAs for ~2 waves of MINB CTAs per SM, at least 32 rows each (a band start costs `size` loads per
  // column), rows in multiples of the batch
  int64_t want = (148 * MINB * 2 + strips - 1) / strips;
  int64_t band = (oh + want - 1) / want;
  if (band < 32) band = 32;
  band = (band + B2_RB - 1) / B2_RB * B2_RB;
  const int64_t bands = (oh + band - 1) / band;
  if (bands > 65535 || strips > 2147483647LL) return fail(FSG_E_UNSUPPORTED, "box mean: grid too large");
  FSG_CUDA_OK(cudaFuncSetAttribute(box_mean2d_kernel<CW, MAXSIZE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BYTES));
  box_mean2d_kernel<CW, MAXSIZE, MINB><<<dim3((unsigned)strips, (unsigned)bands), G::NT, G::BYTES, s>>>(g, size, out, oy0, oh, (int)band);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

// 1 = launched, 0 = the size is outside the single-pass kernel's range (caller falls back to the two passes)
int launch_box_mean2d(const Grid& g, int size, float* out, int64_t oy0, int64_t oh, cudaStream_t s, int* rc) {
  *rc = FSG_OK;
  if (oh <= 0) return 1;
  if (size < 3 || (size & 1) == 0) return 0;
  if (size <= 65) { *rc = launch_b2<192, 65, 3>(g, size, out, oy0, oh, s); return 1; }
  if (size <= 257) { *rc = launch_b2<512, 257, 1>(g, size, out, oy0, oh, s); return 1; }
  return 0;
}

int launch_gauss_taps(double sigma, int radius, double* taps_dev, cudaStream_t s) {
  gauss_taps_kernel<<<1, 256, 0, s>>>(sigma, radius, taps_dev);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

// scratch of the whole-grid void fill: tile map (h * nt bytes, padded), two item lists (h * nt ints each), two counters
struct VoidScratch {
  unsigned char* vt;
  int* vlist;
  int* nlist;
  int* counts;   // [0] void items, [1] needed items
};
static VoidScratch void_scratch(unsigned char* p, int64_t h, int64_t nt) {
  const size_t items = (size_t)h * (size_t)nt;
  const size_t a = (items + 255) / 256 * 256;
  VoidScratch v;
  v.vt = p;
  v.vlist = reinterpret_cast<int*>(p + a);
  v.nlist = v.vlist + items;
  v.counts = v.nlist + items;
  return v;
}
size_t void_fill_need_bytes(int64_t h, int64_t w) {
  const size_t items = (size_t)h * (size_t)((w + VT_W - 1) / VT_W);
  return (items + 255) / 256 * 256 + 2 * items * 4 + 256;
}

int launch_gauss_axis0(const Grid& g, const double* taps_dev, int radius, float* tv, float* tw, const int* run_flag,
                       int64_t oy0, int64_t oh, cudaStream_t s, unsigned char* need_scratch) {
  if (oh <= 0) return FSG_OK;
  const int64_t gx = (g.w + 255) / 256;
  int64_t gy = (148 * 16 + gx - 1) / gx;   // ~16 CTAs per SM in total, rows are strided over
  if (gy > oh) gy = oh;
  dim3 grid((unsigned)gx, (unsigned)gy);
  const int nt = (int)gx;
  const int* list = nullptr;
  const int* list_n = nullptr;
  if (need_scratch && oy0 == 0 && oh == g.h && g.row_off == 0 && g.ld == g.w && (int64_t)g.h * nt < 2147483647LL) {
    // whole-grid void fill: lists of the (row, tile) items that hold a void cell / lie within reach of one
    VoidScratch v = void_scratch(need_scratch, g.h, nt);
    FSG_CUDA_OK(cudaMemsetAsync(v.counts, 0, 2 * sizeof(int), s));
    void_tiles_kernel<<<148 * 4, 256, 0, s>>>(g.src, g.h, g.w, nt, run_flag, v.vt, v.vlist, v.counts);
    FSG_LAUNCH_OK();
    need_tiles_kernel<<<148 * 2, 256, 0, s>>>(v.vt, g.h, nt, radius / VT_W + 1, run_flag, v.nlist, v.counts + 1);
    FSG_LAUNCH_OK();
    list = v.nlist;
    list_n = v.counts + 1;
    grid = dim3(148 * 8);
  }
  gauss_axis0_kernel<<<grid, 256, 0, s>>>(g, taps_dev, radius, tv, tw, run_flag, oy0, oh, list, list_n, nt);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int launch_gauss_axis1(const float* tv, const float* tw, int64_t h, int64_t w, const double* taps_dev, int radius,
                       int combine, float* out, const int* run_flag, int* still_nan, cudaStream_t s,
                       const unsigned char* void_tiles) {
  const int64_t gx = (w + 255) / 256;
  int64_t gy = (148 * 16 + gx - 1) / gx;
  if (gy > h) gy = h;
  dim3 grid((unsigned)gx, (unsigned)gy);
  const int* list = nullptr;
  const int* list_n = nullptr;
  if (void_tiles && combine == COMBINE_VOIDFILL) {   // the items launch_gauss_axis0 listed (same h, w)
    VoidScratch v = void_scratch(const_cast<unsigned char*>(void_tiles), h, gx);
    list = v.vlist;
    list_n = v.counts;
    grid = dim3(148 * 8);
  }
  gauss_axis1_kernel<<<grid, 256, 0, s>>>(tv, tw, h, w, taps_dev, radius, combine, out, run_flag, still_nan, list, list_n,
                                          (int)gx);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

}  // namespace fsg
