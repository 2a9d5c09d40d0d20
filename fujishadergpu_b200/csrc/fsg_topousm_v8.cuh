// fused_kernel_v8: interior / dense fast path of the topousm_fast full-resolution pass for the reference's
// default scale construction (fused radii 2, 8, 32 followed by decimated terms).  Same arithmetic as
// fused_kernel_v6 (reference: compute_topousm_fast_efficient_block, _impl_topousm_fast.py:49-100; box means:
// handle_nan_with_uniform, _nan_utils.py:34-47; zoom: _upsample_to_shape, :671-698) -- bit-identical output.
//
// What is different from v6 (ncu, round 1: 212 thread-instructions per pixel, 37 % issue, ~250 B/px of
// shared-memory traffic at 128 B/clk/SM):
//   * only interior work: every strip column and every window row exists in the raster, the rows are 16-byte
//     aligned, no NaN in reach.  Raster borders are launched on fused_kernel_v6 by the host; a batch that meets a
//     NaN (or Inf) flags its 256-row block, the CTA restarts at the next block, and one more v6 launch redoes the
//     flagged blocks.  No edge, reflect or NaN code in the hot loops.
//   * vertical pass, thread = 4 (radius 32) or 2 (radii 2 and 8) adjacent columns, 16 rows fully unrolled:
//     128-/64-bit ring loads, independent running sums per column, and for the radii 2 / 8 every ring row of
//     the batch is loaded and widened once and serves as entering and leaving tap of both windows (33 loads
//     instead of 64 per column pair and batch).  For the f32-grid rounding the binade of s/n follows from the
//     bits of s (n/2^k is exactly representable), so the magic constant does not wait for the product.
//   * three planes (one per radius, f64) are live at once: the horizontal pass has no barrier between radii.
//     Its tasks are (row, 16-pixel segment) x {radii 2 + 8 | radius 32}; window taps come in as 128-bit loads
//     that stay in registers (each tap of the radii 2 / 8 is loaded once), the 65-tap window of radius 32
//     starts from 16 four-column block sums the vertical pass leaves next to the plane.
//   * decimated terms, normalisation and the store run on thread = output column marching down the batch: the
//     zoom fraction and the coarse columns are per-thread constants, the row fraction comes from a 16-entry
//     table, the four coarse taps are reloaded only when the coarse row changes; the staged fused partial sums
//     are read back transposed (what v6's store staging did anyway).  v6's A0 + t*dA fast form with the
//     rounding-boundary guard and the exact four-tap fallback is kept, evaluated columns first.
//   * two CTA barriers per 16-row batch: [horizontal pass of batch b] | [vertical pass of batch b+1 on warps
//     0-5 || decimated terms + store of batch b on warps 6-11].
//   * the ring holds six groups of 16 rows (group = batch-aligned); every access of a pass stays inside one
//     group (immediate-offset addressing, no wrap tests), a vertical pass touches five groups, and the sixth is
//     being filled by bulk async copies for the batch after the next one: a whole iteration of latency slack.
// Included by fsg_topousm.cu after fsg_topousm_v6.cuh (uses its mbarrier / bulk-copy helpers).
#pragma once

namespace fsg {

constexpr int V8_NB = 16;                         // rows per batch
constexpr int V8_RH = 32;                         // ring halo (largest fused radius)
constexpr int V8_TW = 192;                        // output columns per strip
constexpr int V8_SW = V8_TW + 2 * V8_RH;          // 256 strip columns incl. halo
constexpr int V8_THREADS = 384;                   // 12 warps
constexpr int V8_NGRP = 6;                        // ring = 6 groups of 16 rows: 5 in use by a vertical pass, 1 being filled
constexpr int V8_RROWS = V8_NGRP * V8_NB;         // 96
constexpr int V8_RS = V8_SW + 4;                  // ring row stride (floats): RS/4 odd -> LDS.128 by rows conflict-free
constexpr int V8_PS32 = V8_SW + 2;                // plane row strides (doubles): PS/2 odd -> LDS.128 by rows conflict-free
constexpr int V8_W8 = V8_TW + 16, V8_PS8 = V8_W8 + 2;
constexpr int V8_W2 = V8_TW + 4, V8_PS2 = V8_W2 + 2;
constexpr int V8_PSB = V8_SW / 4 + 2;             // four-column block sums of the radius-32 plane
constexpr int V8_SS = V8_TW + 4;                  // staging row stride (floats): SS/4 odd
constexpr int V8_BLK = 256;                       // rows per NaN fix-up block
constexpr int V8_MAXC = 3;                        // decimated terms
constexpr int V8_NSEG = V8_TW / 16;               // 12 segments of 16 pixels
constexpr int V8_CROWS = 8, V8_CCOLS = 52;        // staged coarse cells per term (decimation >= 4: <= 6 rows, <= 50 columns)


struct V8Smem {
  static constexpr size_t OFF_RING = 0;
  static constexpr size_t OFF_P32 = OFF_RING + (size_t)V8_RROWS * V8_RS * 4;
  static constexpr size_t OFF_P8 = OFF_P32 + (size_t)V8_NB * V8_PS32 * 8;
  static constexpr size_t OFF_P2 = OFF_P8 + (size_t)V8_NB * V8_PS8 * 8;
  static constexpr size_t OFF_B4 = OFF_P2 + (size_t)V8_NB * V8_PS2 * 8;
  static constexpr size_t OFF_STA = OFF_B4 + (size_t)V8_NB * V8_PSB * 8;
  static constexpr size_t OFF_STB = OFF_STA + (size_t)V8_NB * V8_SS * 4;
  static constexpr size_t OFF_TAB = OFF_STB + (size_t)V8_NB * V8_SS * 4;
  // per (decimated term, batch row): zoom row fraction (f64), upper coarse row (int); per term: change mask
  static constexpr size_t OFF_TAB_R0 = OFF_TAB + (size_t)V8_MAXC * V8_NB * 8;
  static constexpr size_t OFF_TAB_CHG = OFF_TAB_R0 + (size_t)V8_MAXC * V8_NB * 4;
  // decimated mean grids around the batch: V8_CROWS rows x V8_CCOLS columns per term (f32), clamped at the grid edges
  static constexpr size_t OFF_CELLS = OFF_TAB_CHG + 16;
  static constexpr size_t OFF_BAR = OFF_CELLS + (size_t)V8_MAXC * V8_CROWS * V8_CCOLS * 4;
  static constexpr size_t BYTES = OFF_BAR + 16;
  static_assert(OFF_P32 % 16 == 0 && OFF_P8 % 16 == 0 && OFF_P2 % 16 == 0 && OFF_B4 % 16 == 0 && OFF_STA % 16 == 0 &&
                    OFF_STB % 16 == 0 && OFF_TAB % 16 == 0 && OFF_BAR % 16 == 0,
                "128-bit shared-memory accesses need 16-byte aligned regions");
  static_assert(BYTES <= 227 * 1024, "v8 shared memory");
};
static_assert((V8_RS / 4) % 2 == 1 && (V8_PS32 / 2) % 2 == 1 && (V8_PS8 / 2) % 2 == 1 && (V8_PS2 / 2) % 2 == 1 &&
                  (V8_PSB / 2) % 2 == 1 && (V8_SS / 4) % 2 == 1,
              "row strides must be odd multiples of 16 bytes");
static_assert(V8_SW == 256 && V8_THREADS == 2 * V8_NB * V8_NSEG && V8_TW == (V8_THREADS / 32 - 6) * 32, "v8 role layout");

// f32-rounded value (kept as f64) of the f64 mean s * (1 / N) for N = 5, 17, 65: (q + M) - M with M = +-2^(e_q + 29)
// rounds q to the f32 grid of its binade, ties to even (SciPy's f64 mean followed by the f32 store; same two
// roundings as fused_kernel_v6, so exact ties -- windows that span a binade -- break the same way).  N / 2^K is
// exactly representable, so whether the mantissa of s is >= N / 2^K (one integer add on the high word) tells the
// binade of the quotient and M does not depend on the product.
template <int N>
__device__ __forceinline__ double v8_mean_f32grid(double s) {
  constexpr int K = N == 5 ? 2 : (N == 17 ? 4 : 6);
  static_assert(N == 5 || N == 17 || N == 65, "window sizes of the radii 2, 8, 32");
  constexpr unsigned FR = ((unsigned)(N - (1 << K)) << 20) >> K;   // (N / 2^K - 1) * 2^20
  constexpr unsigned ADD = (0x100000u - FR) + ((unsigned)(29 - (K + 1)) << 20);
  const unsigned hi = (unsigned)__double2hiint(s);
  const double M = __hiloint2double((int)((hi + ADD) & 0xfff00000u), 0);
  const double q = s * (1.0 / (double)N);
  return (q + M) - M;
}

__device__ __forceinline__ bool v8_finite(double s) { return fabs(s) < __longlong_as_double(0x7ff0000000000000LL); }

#ifdef FSG_V8_TIMERS
// per-warp cycle counters of CTA (0, 0): [warp][0..3] = busy phase A, wait at barrier A, busy phase B, wait at barrier B
__device__ unsigned long long v8_timers[V8_THREADS / 32][4];
#define V8_TICK(slot) do { if (dbg_on) { const long long t_ = clock64(); dbg_acc[slot] += t_ - dbg_t; dbg_t = t_; } } while (0)
#else
#define V8_TICK(slot) do { } while (0)
#endif

template <int NCO>
__global__ void __launch_bounds__(V8_THREADS, 1) fused_kernel_v8(FusedParams p) {
  static_assert(NCO >= 0 && NCO <= V8_MAXC, "decimated terms");
  extern __shared__ __align__(16) unsigned char smraw[];
  const NormDev nd = resolve_norm(p);
  float* ring = reinterpret_cast<float*>(smraw + V8Smem::OFF_RING);
  double* P32 = reinterpret_cast<double*>(smraw + V8Smem::OFF_P32);
  double* P8 = reinterpret_cast<double*>(smraw + V8Smem::OFF_P8);
  double* P2 = reinterpret_cast<double*>(smraw + V8Smem::OFF_P2);
  double* B4 = reinterpret_cast<double*>(smraw + V8Smem::OFF_B4);
  float* stA = reinterpret_cast<float*>(smraw + V8Smem::OFF_STA);
  float* stB = reinterpret_cast<float*>(smraw + V8Smem::OFF_STB);
  double* rt_tr = reinterpret_cast<double*>(smraw + V8Smem::OFF_TAB);
  int* rt_off = reinterpret_cast<int*>(smraw + V8Smem::OFF_TAB_R0);   // (coarse row - first staged row) * V8_CCOLS
  float* cells = reinterpret_cast<float*>(smraw + V8Smem::OFF_CELLS);
  unsigned char* rt_act = smraw + V8Smem::OFF_TAB_CHG;   // per batch row: bit k = term k changes its coarse row here
  const unsigned bar = smem_u32(smraw + V8Smem::OFF_BAR);
  const unsigned ring_sa = smem_u32(ring);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int x0r = p.v8_col0 + (int)blockIdx.x * V8_TW;
  const int x0 = x0r < p.v8_xlast ? x0r : p.v8_xlast;
  const int cs0 = x0 - V8_RH;
  const int64_t yb0 = p.v8_row0 + (int64_t)blockIdx.y * p.band_rows;
  const int64_t yb1 = (yb0 + p.band_rows < p.v8_row1) ? yb0 + p.band_rows : p.v8_row1;
  const float w2 = p.terms[0].weight, w8 = p.terms[1].weight, w32 = p.terms[2].weight;
  constexpr int nco = NCO;

  // ---- roles of the second phase ----
  const bool is_v32 = warp < 2;                       // thread = strip columns 4 t .. 4 t + 3
  const bool is_st = warp >= 2 && warp < 6;           // thread = strip columns 24 + 2 ts, + 1 (radii 8 and 2)
  const int ts = tid - 64;
  const bool st_on = is_st && ts < V8_W8 / 2;
  const bool st_r2 = st_on && ts >= 3 && ts < 3 + V8_W2 / 2;
  const int tc = tid - 192;                           // warps 6-11: output column x0 + tc

  // ---- first-phase task: (group, row, segment); lanes 0-15 / 16-31 of a warp = the 16 rows of two segments ----
  const int hgrp = tid >= V8_THREADS / 2 ? 1 : 0;     // 0: radii 2 + 8, 1: radius 32
  const int hu = tid - hgrp * (V8_THREADS / 2);
  const int hrow = hu & 15, hseg = hu >> 4;

  // ---- decimated terms: per-thread column constants (warps 6-11) ----
  double c_tc[V8_MAXC];
  int c_cc[V8_MAXC];      // staged column of the left coarse tap
  int c_base[V8_MAXC];    // first staged coarse column of the strip (every thread)
  double c_b0[V8_MAXC], c_db[V8_MAXC];
#pragma unroll
  for (int k = 0; k < V8_MAXC; ++k) {
    c_tc[k] = 0.0; c_cc[k] = 0; c_base[k] = 0; c_b0[k] = 0.0; c_db[k] = 0.0;
    if (k < nco) {
      const DevTerm& T = p.terms[3 + k];
      const int gwm1 = (int)T.gw - 1;
      int cb = (int)floor((double)x0 * T.cscale);
      c_base[k] = cb > gwm1 ? gwm1 : cb;
      if (tc >= 0) {
        double ci = (double)(x0 + tc) * T.cscale;
        double fl = floor(ci);
        if (fl > (double)gwm1) fl = (double)gwm1;
        c_tc[k] = ci - fl;
        c_cc[k] = (int)fl - c_base[k];
      }
    }
  }

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // Fills alternate between the two mbarriers; every thread waits for every fill exactly once, in order.
  unsigned n_issued = 0, n_waited = 0;   // (CTA-uniform)
  auto wait_fill = [&]() {
    mbar_wait(bar + 8u * (n_waited & 1u), (n_waited >> 1) & 1u);
    ++n_waited;
  };
  // bulk copies of DEM rows [row_first, row_first + 16 * ngroups) into ring groups grp0, grp0 + 1, ...  Warp 0 only
  // issues; every thread counts.
  auto issue_fill = [&](int64_t row_first, int ngroups, int grp0) {
    const unsigned b = bar + 8u * (n_issued & 1u);
    ++n_issued;
    if (warp != 0) return;
    const int n = ngroups * V8_NB;
    if (lane == 0) mbar_expect_tx(b, (unsigned)n * (unsigned)(V8_SW * 4));
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      const float* src = p.dem + (row_first + k - p.dem_row0) * p.ld_in + cs0;
      bulk_copy_g2s(ring_sa + (unsigned)(grp0 * V8_NB + k) * (unsigned)(V8_RS * 4), src, (unsigned)(V8_SW * 4), b);
    }
  };
  // first ring row of group (batch + d) when batch - 2 sits in group gb
  auto grp_row = [](int gb, int d) {
    int g = gb + d + 2;
    return (g >= V8_NGRP ? g - V8_NGRP : g) * V8_NB;
  };

  // running window sums (f64, exact): radius 32 on four columns, radii 8 / 2 on two columns
  double s32[4] = {0.0, 0.0, 0.0, 0.0};
  double s8[2] = {0.0, 0.0}, s2[2] = {0.0, 0.0};

  // ---- vertical pass of batch b (first row Y0) when the rows Y0 - 32 .. Y0 - 17 sit in group gb ----
  // s32 enters as (window sum of row Y0) - x[Y0 + 32] and leaves in the same state for row Y0 + 16, so the pass
  // touches the five groups b - 2 .. b + 2 only.
  auto vpass = [&](int gb) -> int {
    if (is_v32) {
      const float4* pin = reinterpret_cast<const float4*>(ring + grp_row(gb, 2) * V8_RS) + tid;
      const float4* pout = reinterpret_cast<const float4*>(ring + grp_row(gb, -2) * V8_RS) + tid;
      double2* dst = reinterpret_cast<double2*>(P32) + 2 * tid;
      double* b4 = B4 + tid;
      // The ring loads and the plane stores both go to shared memory, so the compiler keeps their order: the rows
      // are loaded four at a time, one chunk ahead of the chunk being computed and stored.
      float4 cin[2][4], cout[2][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cin[0][u] = pin[u * (V8_RS / 4)];
        cout[0][u] = pout[u * (V8_RS / 4)];
      }
      s32[0] += (double)cin[0][0].x; s32[1] += (double)cin[0][0].y;
      s32[2] += (double)cin[0][0].z; s32[3] += (double)cin[0][0].w;
#pragma unroll
      for (int c = 0; c < V8_NB / 4; ++c) {
        if (c + 1 < V8_NB / 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            cin[(c + 1) & 1][u] = pin[(4 * (c + 1) + u) * (V8_RS / 4)];
            cout[(c + 1) & 1][u] = pout[(4 * (c + 1) + u) * (V8_RS / 4)];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = 4 * c + u;
          const double q0 = v8_mean_f32grid<65>(s32[0]), q1 = v8_mean_f32grid<65>(s32[1]);
          const double q2 = v8_mean_f32grid<65>(s32[2]), q3 = v8_mean_f32grid<65>(s32[3]);
          dst[i * (V8_PS32 / 2)] = make_double2(q0, q1);
          dst[i * (V8_PS32 / 2) + 1] = make_double2(q2, q3);
          b4[i * V8_PSB] = (q0 + q1) + (q2 + q3);
          const float4 b = cout[c & 1][u];
          if (i + 1 < V8_NB) {
            const float4 a = u + 1 < 4 ? cin[c & 1][u + 1] : cin[(c + 1) & 1][0];
            s32[0] += (double)a.x - (double)b.x;
            s32[1] += (double)a.y - (double)b.y;
            s32[2] += (double)a.z - (double)b.z;
            s32[3] += (double)a.w - (double)b.w;
          } else {
            s32[0] -= (double)b.x; s32[1] -= (double)b.y; s32[2] -= (double)b.z; s32[3] -= (double)b.w;
          }
        }
      }
      return v8_finite((fabs(s32[0]) + fabs(s32[1])) + (fabs(s32[2]) + fabs(s32[3]))) ? 0 : 1;
    }
    if (st_on) {
      // rows Y0 - 8 + k, k = 0..32: the second half of group b - 1, group b, the first nine rows of group b + 1
      const float2* pa = reinterpret_cast<const float2*>(ring + (grp_row(gb, -1) + 8) * V8_RS) + 12 + ts;
      const float2* pb = reinterpret_cast<const float2*>(ring + grp_row(gb, 0) * V8_RS) + 12 + ts;
      const float2* pc = reinterpret_cast<const float2*>(ring + grp_row(gb, 1) * V8_RS) + 12 + ts;
      double wx[33], wy[33];
#pragma unroll
      for (int k = 0; k < 33; ++k) {
        const float2 v = k < 8 ? pa[k * (V8_RS / 2)] : (k < 24 ? pb[(k - 8) * (V8_RS / 2)] : pc[(k - 24) * (V8_RS / 2)]);
        wx[k] = (double)v.x;
        wy[k] = (double)v.y;
      }
      double2* d8 = reinterpret_cast<double2*>(P8) + ts;
      double2* d2 = reinterpret_cast<double2*>(P2) + (ts - 3);
#pragma unroll
      for (int i = 0; i < V8_NB; ++i) {
        d8[i * (V8_PS8 / 2)] = make_double2(v8_mean_f32grid<17>(s8[0]), v8_mean_f32grid<17>(s8[1]));
        const double2 m2 = make_double2(v8_mean_f32grid<5>(s2[0]), v8_mean_f32grid<5>(s2[1]));
        if (st_r2) d2[i * (V8_PS2 / 2)] = m2;
        s8[0] += wx[i + 17] - wx[i];
        s8[1] += wy[i + 17] - wy[i];
        s2[0] += wx[i + 11] - wx[i + 6];
        s2[1] += wy[i + 11] - wy[i + 6];
      }
    }
    return 0;
  };

  // ---- horizontal pass of the batch (all 12 warps) ----
  auto hpass = [&](int gb) {
    const int sx = grp_row(gb, 0);
    const float4* xp = reinterpret_cast<const float4*>(ring + (sx + hrow) * V8_RS) + (V8_RH / 4 + 4 * hseg);
    float x[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = xp[q];
      x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
    float o[16];
    if (hgrp == 0) {
      const double2* p2 = reinterpret_cast<const double2*>(P2) + hrow * (V8_PS2 / 2) + 8 * hseg;
      const double2* p8 = reinterpret_cast<const double2*>(P8) + hrow * (V8_PS8 / 2) + 8 * hseg;
      double a[20], b[32];
#pragma unroll
      for (int q = 0; q < 10; ++q) { const double2 v = p2[q]; a[2 * q] = v.x; a[2 * q + 1] = v.y; }
#pragma unroll
      for (int q = 0; q < 16; ++q) { const double2 v = p8[q]; b[2 * q] = v.x; b[2 * q + 1] = v.y; }
      double W2 = ((a[0] + a[1]) + (a[2] + a[3])) + a[4];
      double W8 = (((b[0] + b[1]) + (b[2] + b[3])) + ((b[4] + b[5]) + (b[6] + b[7]))) +
                  (((b[8] + b[9]) + (b[10] + b[11])) + ((b[12] + b[13]) + (b[14] + b[15]))) + b[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float m2 = (float)(W2 * (1.0 / 5.0));
        const float m8 = (float)(W8 * (1.0 / 17.0));
        float acc = 0.f + w2 * (x[j] - m2);
        acc = acc + w8 * (x[j] - m8);
        o[j] = acc;
        if (j < 15) {
          W2 += a[j + 5] - a[j];
          W8 += b[j + 17] - b[j];
        }
      }
    } else {
      const double2* pl = reinterpret_cast<const double2*>(P32) + hrow * (V8_PS32 / 2) + 8 * hseg;
      const double2* pb = reinterpret_cast<const double2*>(B4) + hrow * (V8_PSB / 2) + 2 * hseg;
      double bs[16], lv[16], en[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) { const double2 v = pb[q]; bs[2 * q] = v.x; bs[2 * q + 1] = v.y; }
#pragma unroll
      for (int q = 0; q < 8; ++q) { const double2 v = pl[q]; lv[2 * q] = v.x; lv[2 * q + 1] = v.y; }
#pragma unroll
      for (int q = 0; q < 8; ++q) { const double2 v = pl[32 + q]; en[2 * q] = v.x; en[2 * q + 1] = v.y; }
      double Wv = (((bs[0] + bs[1]) + (bs[2] + bs[3])) + ((bs[4] + bs[5]) + (bs[6] + bs[7]))) +
                  (((bs[8] + bs[9]) + (bs[10] + bs[11])) + ((bs[12] + bs[13]) + (bs[14] + bs[15]))) + en[0];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float m = (float)(Wv * (1.0 / 65.0));
        o[j] = w32 * (x[j] - m);
        if (j < 15) Wv += en[j + 1] - lv[j];
      }
    }
    float4* sp = reinterpret_cast<float4*>(hgrp == 0 ? stA : stB) + hrow * (V8_SS / 4) + 4 * hseg;
#pragma unroll
    for (int q = 0; q < 4; ++q) sp[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  };

  // ---- decimated terms, normalisation, store: warps 6-11, thread = output column, 16 rows ----
  // normalisation as one branch-free form: v = acc / sc correctly rounded (q = acc * rinv and one FMA residual
  // step, as in v6); (sc, rinv) = (1, 1) leaves acc untouched, (1, 0) gives the zeros of a non-positive scale
  // (v8 never sees a NaN).
  const float n_sc = nd.mode == 1 ? nd.sc : 1.f;
  const float n_rinv = nd.mode == 1 ? nd.rinv : (nd.mode == 2 ? 0.f : 1.f);
  float cw[V8_MAXC];
#pragma unroll
  for (int k = 0; k < V8_MAXC; ++k) cw[k] = k < nco ? p.terms[3 + k].weight : 0.f;
  // the four taps of a column thread: staged rows roff, roff + 1 row, staged columns c_cc, c_cc + 1 (the staging
  // clamps at the grid edges, so the "+ 1" taps always exist)
  auto reload_cell = [&](int k, int roff) {
    const float* g = cells + k * (V8_CROWS * V8_CCOLS) + roff + c_cc[k];
    const double a00 = (double)g[0], a01 = (double)g[1];
    const double a10 = (double)g[V8_CCOLS], a11 = (double)g[V8_CCOLS + 1];
    const double b0 = fma(c_tc[k], a01 - a00, a00), b1 = fma(c_tc[k], a11 - a10, a10);
    c_b0[k] = b0;
    c_db[k] = b1 - b0;
  };
  auto exact_taps = [&](int k, int roff, double tr) -> double {   // scipy's four-tap sum, op for op (as in v6)
    const float* g = cells + k * (V8_CROWS * V8_CCOLS) + roff + c_cc[k];
    const double wr0 = 1.0 - tr, wc0 = 1.0 - c_tc[k];
    const double p00 = (double)g[0] * wr0, p01 = (double)g[1] * wr0;
    const double p10 = (double)g[V8_CCOLS] * tr, p11 = (double)g[V8_CCOLS + 1] * tr;
    double v = p00 * wc0;
    v += p01 * c_tc[k];
    v += p10 * wc0;
    v += p11 * c_tc[k];
    return v;
  };
  auto cpass = [&](int64_t Y0, int gb) {
    const float* xc = ring + grp_row(gb, 0) * V8_RS + V8_RH + tc;
    const float* pa = stA + tc;
    const float* pb = stB + tc;
    const size_t esz = p.enc.kind == FSG_OUT_F32 ? 4 : (p.enc.kind == FSG_OUT_U8 ? 1 : 2);
    unsigned char* op = (unsigned char*)p.out + ((Y0 - p.out_row0) * p.ld_out + x0 + tc) * (int64_t)esz;
    const size_t ostep = (size_t)p.ld_out * esz;
#pragma unroll 1
    for (int g = 0; g < V8_NB / 4; ++g) {
      float xv[4], acc[4];
      double m64[NCO > 0 ? NCO : 1][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xv[u] = xc[(4 * g + u) * V8_RS];
        acc[u] = pa[(4 * g + u) * V8_SS] + pb[(4 * g + u) * V8_SS];
      }
      if (NCO > 0) {
        // rows at which a term moves to the next coarse row: one byte per row, warp-uniform (REDUX -> uniform register)
        const unsigned act4 = __reduce_or_sync(0xffffffffu, reinterpret_cast<const unsigned*>(rt_act)[g]);
        unsigned risk = 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = 4 * g + u;
          const unsigned act = (act4 >> (8 * u)) & 7u;
          if (act != 0u) {
#pragma unroll
            for (int k = 0; k < NCO; ++k)
              if ((act >> k) & 1u) reload_cell(k, rt_off[k * V8_NB + i]);
          }
#pragma unroll
          for (int k = 0; k < NCO; ++k) {
            const double m = fma(rt_tr[k * V8_NB + i], c_db[k], c_b0[k]);
            m64[k][u] = m;
            // the fast form agrees with scipy's four-tap sum to a few f64 ulps: within V6_GUARD ulps of an f32
            // rounding boundary the four-tap form decides (see fused_kernel_v6)
            const unsigned rk = ((unsigned)__double2loint(m) + (V6_GUARD - 0x10000000u)) << 3;
            risk = rk < risk ? rk : risk;
          }
        }
        if (risk < (2u * V6_GUARD) << 3) {
#pragma unroll
          for (int k = 0; k < NCO; ++k) {
#pragma unroll 1
            for (int u = 0; u < 4; ++u) {
              const double v = exact_taps(k, rt_off[k * V8_NB + 4 * g + u], rt_tr[k * V8_NB + 4 * g + u]);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q == u) m64[k][q] = v;
            }
          }
        }
      }
      float vo[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float a = acc[u];
#pragma unroll
        for (int k = 0; k < NCO; ++k) a = a + cw[k] * (xv[u] - (float)m64[k][u]);
        const float q = a * n_rinv;
        const float rem = fmaf(-q, n_sc, a);
        vo[u] = fmaf(rem, n_rinv, q);
      }
      if (p.enc.kind == FSG_OUT_F32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<float*>(op + u * ostep) = vo[u];
      } else if (p.enc.kind == FSG_OUT_U8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<uint8_t*>(op + u * ostep) = (uint8_t)(int)encode_dn(vo[u], p.enc);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<int16_t*>(op + u * ostep) = (int16_t)(int)encode_dn(vo[u], p.enc);
      }
      op += 4 * ostep;
    }
  };

#ifdef FSG_V8_TIMERS
  const bool dbg_on = blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && lane == 0;
  long long dbg_acc[4] = {0, 0, 0, 0}, dbg_t = clock64();
#endif
  int64_t y = yb0;
  while (y < yb1) {
    // ---- (re)start at row y: groups -2 .. 2 (rows y - 32 .. y + 47) into ring groups 0 .. 4, window sums ----
    __syncthreads();
    issue_fill(y - V8_RH, 5, 0);
    wait_fill();
    if (is_v32) {   // rows y - 32 .. y + 31: the window sum of row y without its last row
      const float4* pr = reinterpret_cast<const float4*>(ring) + tid;
      double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll 4
      for (int d = 0; d < 2 * V8_RH; ++d) {
        const float4 v = pr[d * (V8_RS / 4)];
        t0 += (double)v.x; t1 += (double)v.y; t2 += (double)v.z; t3 += (double)v.w;
      }
      s32[0] = t0; s32[1] = t1; s32[2] = t2; s32[3] = t3;
    } else if (st_on) {
      const float2* pr = reinterpret_cast<const float2*>(ring + (V8_RH - 8) * V8_RS) + 12 + ts;
      double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
#pragma unroll
      for (int d = 0; d <= 16; ++d) {
        const float2 v = pr[d * (V8_RS / 2)];
        t0 += (double)v.x; t1 += (double)v.y;
        if (d >= 6 && d <= 10) { u0 += (double)v.x; u1 += (double)v.y; }
      }
      s8[0] = t0; s8[1] = t1; s2[0] = u0; s2[1] = u1;
    }
    // software pipeline over batches: phase A = vertical pass of batch yv (warps 0-5) || decimated terms + store of
    // batch yc (warps 6-11); phase B = horizontal pass of batch yv.  The group batch yv + 16 adds to the ring is
    // requested at the start of phase A of batch yv.
    int64_t yv = y, yc = -1;
    int gbv = 0, gbc = 0;
    bool first = true;
    for (;;) {
      const bool do_v = yv < yb1, do_c = yc >= 0;
      const bool pre = do_v && yv + V8_NB < yb1;      // batch yv + 16 exists: its last group is fetched now
      const bool wt = do_v && !first;                 // group yv + 2 was requested one iteration ago
      int vbad = 0;
      if (warp < 6) {
        if (wt) wait_fill();
        if (pre) issue_fill(yv + 3 * V8_NB, 1, grp_row(gbv, 3) / V8_NB);
        if (do_v) vbad = vpass(gbv);
      } else {
        if (pre) issue_fill(0, 0, 0);   // (counts the fill)
        if (do_c) cpass(yc, gbc);
        if (wt) wait_fill();
      }
      V8_TICK(0);
      const int bad = __syncthreads_or(vbad);
      V8_TICK(1);
      if (bad) {   // NaN / Inf in reach of batch yv: its block goes to the general kernel, restart behind it
        if (pre) wait_fill();
        const int64_t blk = (yv - p.v8_row0) / V8_BLK;
        if (tid == 0) p.v8_flags[blk * gridDim.x + blockIdx.x] = 1;
        y = p.v8_row0 + (blk + 1) * V8_BLK;
        break;
      }
      if (!do_v) {
        y = yb1;
        break;
      }
      // ---- phase B ----
      if (warp == 6 && NCO > 0) {   // zoom row table of batch yv: lanes 0-15 = rows of terms 0 and 2, lanes 16-31 = term 1
        const int i = lane & 15;
        unsigned bal[2] = {0u, 0u};
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const int k = pass == 0 ? (lane >> 4) : 2;
          int r0i = 0, rprev = -1;
          if (k < NCO && (pass == 0 || lane < 16)) {
            const DevTerm& T = p.terms[3 + k];
            const double ri = (double)(yv + i) * T.rscale;
            int64_t r0 = (int64_t)floor(ri);
            if (r0 > T.gh - 1) r0 = T.gh - 1;
            int64_t rb = (int64_t)floor((double)yv * T.rscale);   // first staged row
            if (rb > T.gh - 1) rb = T.gh - 1;
            rt_tr[k * V8_NB + i] = ri - (double)r0;
            rt_off[k * V8_NB + i] = (int)(r0 - rb) * V8_CCOLS;
            r0i = (int)r0;
            if (i == 0 && !first) {   // coarse row of the row above the batch (the cell a column thread still holds)
              int64_t rp = (int64_t)floor((double)(yv - 1) * T.rscale);
              if (rp > T.gh - 1) rp = T.gh - 1;
              rprev = (int)rp;
            }
          }
          const int up = __shfl_up_sync(0xffffffffu, r0i, 1, 16);
          bal[pass] = __ballot_sync(0xffffffffu, r0i != (i == 0 ? rprev : up));
        }
        if (lane < 16) {
          unsigned a = (bal[0] >> i) & 1u;
          if (NCO > 1) a |= ((bal[0] >> (16 + i)) & 1u) << 1;
          if (NCO > 2) a |= ((bal[1] >> i) & 1u) << 2;
          rt_act[i] = (unsigned char)a;
        }
      }
      // warps 8-11: the coarse cells the batch touches (rows / columns clamped at the grid edges) are loaded now and
      // stored behind the horizontal pass (two (term, column) items per thread, up to V8_CROWS rows each)
      float cv[2][V8_CROWS];
      int cv_dst[2] = {-1, -1}, cv_n[2] = {0, 0};
      if (warp >= 8) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int idx = tid - 256 + 128 * it;
          const int k = idx / V8_CCOLS, cc = idx - k * V8_CCOLS;
          if (k < nco) {
            const DevTerm& T = p.terms[3 + k];
            int64_t rb = (int64_t)floor((double)yv * T.rscale);
            if (rb > T.gh - 1) rb = T.gh - 1;
            int64_t rl = (int64_t)floor((double)(yv + V8_NB - 1) * T.rscale) + 1;   // last row a tap can touch
            if (rl > T.gh - 1) rl = T.gh - 1;
            const int gwm1 = (int)T.gw - 1;
            int gc = c_base[k] + cc;
            if (gc > gwm1) gc = gwm1;
            const float* src = T.grid + (rb - T.grow0) * T.gw + gc;
            const int nr = (int)(rl - rb) + 1;
            cv_n[it] = nr < V8_CROWS ? nr : V8_CROWS;
            cv_dst[it] = k * (V8_CROWS * V8_CCOLS) + cc;
#pragma unroll
            for (int rr = 0; rr < V8_CROWS; ++rr)
              cv[it][rr] = rr < cv_n[it] ? __ldg(src + (int64_t)rr * T.gw) : 0.f;
          }
        }
      }
      hpass(gbv);
      if (warp >= 8) {
#pragma unroll
        for (int it = 0; it < 2; ++it)
          if (cv_dst[it] >= 0) {
#pragma unroll
            for (int rr = 0; rr < V8_CROWS; ++rr)   // (rows behind the last one are never read: taps stop at rl)
              if (rr < cv_n[it]) cells[cv_dst[it] + rr * V8_CCOLS] = cv[it][rr];
          }
      }
      V8_TICK(2);
      __syncthreads();
      V8_TICK(3);
      yc = yv;
      gbc = gbv;
      yv += V8_NB;
      gbv = gbv + 1 == V8_NGRP ? 0 : gbv + 1;
      first = false;
    }
  }
#ifdef FSG_V8_TIMERS
  if (dbg_on)
    for (int q = 0; q < 4; ++q) v8_timers[warp][q] = (unsigned long long)dbg_acc[q];
#endif
}

}  // namespace fsg
