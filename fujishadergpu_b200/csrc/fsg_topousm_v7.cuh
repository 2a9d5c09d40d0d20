// fused_kernel_v7: role-split, software-pipelined full-resolution pass of topousm_fast.
// Same arithmetic as fused_kernel / fused_kernel_v6 (reference: compute_topousm_fast_efficient_block,
// _impl_topousm_fast.py:49-100; handle_nan_with_uniform, _nan_utils.py:34-47; _upsample_to_shape,
// :671-698), bit-identical results.
//
// Why: the ncu profiles of v5/v6 show two different bottlenecks that never overlap because the phases
// are separated by CTA barriers -- the vertical pass is bound by f32->f64 conversions (XU pipe,
// 16 lanes/clk/SM) and the horizontal / coarse passes by shared-memory bandwidth (128 B/clk/SM).
// Here the CTA is split by role: warps 5-9 run the vertical pass of fused term k+1 (two adjacent
// columns per thread) WHILE warps 0-4 run the horizontal pass of term k, the coarse terms, the epilogue
// and the stores of the previous batch.  Two vertical-mean planes alternate between the roles; one
// __syncthreads per step keeps them in lock step (no mbarrier between the roles).  Batches are 16 rows;
// the ring holds one batch of rows ahead of the vertical pass (bulk async copies, two in flight).
// Output goes straight from registers to global memory (6 x 16-byte stores per thread).
#pragma once

namespace fsg {

constexpr int V7_THREADS = 320;
constexpr int V7_HT = 160;        // threads of the horizontal role (warps 0-4); the rest run the vertical role
constexpr int V7_NB = 16;
constexpr int V7_TW = 240;
constexpr int V7_SEG = 24;
constexpr int V7_NSEG = V7_TW / V7_SEG;   // 10
constexpr int V7_KMAX = 7;        // coarse cells a 24-pixel segment can touch at decimation >= 4
constexpr int V7_MAXLV = 2;
constexpr int V7_MAXF = 4;
static_assert(V7_NSEG * V7_NB == V7_HT, "one horizontal thread per (row, segment)");

template <int RH>
struct V7Geom {
  static constexpr int SW = V7_TW + 2 * RH;
  static constexpr int RS = ((SW / 4) % 2 == 1) ? SW : SW + 4;
  static constexpr int PS = SW | 1;
  static constexpr int NRING = 2 * V7_NB + 2 * RH + 1;
  static constexpr int NSLOT = V7_NB + 2 * RH + 2;     // slot table entries, d in [-RH-1, NB+RH]
  static constexpr size_t PLANE_BYTES = ((size_t)V7_NB * PS * 8 + 15) / 16 * 16;
  static constexpr size_t OFF_PLANE = ((size_t)NRING * RS * 4 + 15) / 16 * 16;
  static constexpr size_t OFF_CELL = OFF_PLANE + 2 * PLANE_BYTES;
  static constexpr size_t CELL_BYTES = (size_t)V7_NSEG * V7_KMAX * V7_NB * 16;
  static constexpr size_t OFF_SLOT = OFF_CELL + CELL_BYTES;
  static constexpr size_t OFF_TC = (OFF_SLOT + (size_t)NSLOT * 4 + 15) / 16 * 16;
  static constexpr size_t OFF_BAR = OFF_TC + (size_t)V7_MAXLV * V7_TW * sizeof(V6ColEntry);
  static constexpr size_t BYTES = OFF_BAR + 16 + 128;  // two mbarriers + step flags (4 slots x 4) + start flags
  static_assert(SW % 2 == 0 && SW / 2 <= V7_THREADS - V7_HT, "two columns per vertical thread");
  static_assert(RH >= V7_NB, "the ring keeps the centre rows of the previous batch for the horizontal role");
};

__device__ __forceinline__ void v7_bar_vrole() { asm volatile("bar.sync 1, %0;\n" ::"n"(V7_THREADS - V7_HT) : "memory"); }

template <int RH>
__global__ void __launch_bounds__(V7_THREADS, 1) fused_kernel_v7(FusedParams p) {
  using G = V7Geom<RH>;
  constexpr int NB = V7_NB, SEG = V7_SEG, TW = V7_TW, SW = G::SW, RS = G::RS, PS = G::PS, NRING = G::NRING;
  extern __shared__ __align__(16) unsigned char smraw[];
  float* ring = reinterpret_cast<float*>(smraw);
  int* slot_tab = reinterpret_cast<int*>(smraw + G::OFF_SLOT);
  V6ColEntry* coltab = reinterpret_cast<V6ColEntry*>(smraw + G::OFF_TC);
  const unsigned bar0 = smem_u32(smraw + G::OFF_BAR);
  volatile int* sflag = reinterpret_cast<volatile int*>(smraw + G::OFF_BAR + 16);   // [4 slots][bad, notfull, mode, spare]

  const int tid = threadIdx.x;
  const bool is_h = tid < V7_HT;
  const int W = (int)p.W;
  const int64_t H = p.H;
  const int x0 = ((int)blockIdx.x + p.strip0) * TW;
  const int cs0 = x0 - RH;
  const int64_t out_end = p.out_row0 + p.out_rows;
  const int64_t yb0 = p.out_row0 + (int64_t)blockIdx.y * p.band_rows;
  const int64_t yb1 = (yb0 + p.band_rows < out_end) ? yb0 + p.band_rows : out_end;
  const bool edge_strip = (cs0 < 0) || (cs0 + SW > W);
  const bool bulk = p.bulk_ok && !edge_strip;

  // ---- term bookkeeping (uniform) ----
  int fidx[V7_MAXF];
  int nf = 0;
#pragma unroll
  for (int k = 0; k < V7_MAXF; ++k) fidx[k] = -1;
  for (int t = 0; t < p.n_terms; ++t) {
    if (p.terms[t].kind == TERM_BOX_FUSED) {
#pragma unroll
      for (int k = 0; k < V7_MAXF; ++k) if (k == nf) fidx[k] = t;
      ++nf;
    }
  }
  auto fused_at = [&](int j) {
    int v = -1;
#pragma unroll
    for (int k = 0; k < V7_MAXF; ++k) if (k == j) v = fidx[k];
    return v;
  };

  // ---- vertical role: two adjacent columns per thread ----
  const int vt = tid - V7_HT;
  const int vc = 2 * vt;                      // strip slot of the first column
  const bool vpair = !is_h && vc < SW;
  const bool vok0 = vpair && cs0 + vc >= 0 && cs0 + vc < W;
  const bool vok1 = vpair && cs0 + vc + 1 >= 0 && cs0 + vc + 1 < W;
  double rs0[V7_MAXF], rs1[V7_MAXF];
  int rc0[V7_MAXF], rc1[V7_MAXF];
#pragma unroll
  for (int k = 0; k < V7_MAXF; ++k) { rs0[k] = rs1[k] = 0.0; rc0[k] = rc1[k] = 0; }
  unsigned nanmask = 0;   // bit j: fused term j is in the NaN-aware form (uniform)

  // ---- horizontal role: (row, segment) ----
  const int hi = tid % NB;
  const int hg = tid / NB;
  const int hj0 = hg * SEG;
  int hjn = SEG;
  if (x0 + hj0 + hjn > W) hjn = (W - x0 - hj0 > 0) ? (W - x0 - hj0) : 0;
  if (!is_h) hjn = 0;
  float acc[SEG], xr[SEG];
#pragma unroll
  for (int jj = 0; jj < SEG; ++jj) { acc[jj] = 0.f; xr[jj] = 0.f; }

  int lv_c0[V7_MAXLV];
#pragma unroll
  for (int l = 0; l < V7_MAXLV; ++l) {
    lv_c0[l] = 0;
    if (l < p.n_lvls) {
      const double cs = p.lvl_cscale[l];
      const int gwm1 = p.lvl_gw[l] - 1;
      for (int j = tid; j < TW; j += V7_THREADS) {
        double ci = (double)(x0 + j) * cs;
        double fl = floor(ci);
        if (fl > (double)gwm1) fl = (double)gwm1;
        int cseg = (int)floor((double)(x0 + (j / SEG) * SEG) * cs);
        if (cseg > gwm1) cseg = gwm1;
        V6ColEntry e;
        e.tc = ci - fl;
        e.koff = ((int)fl - cseg) * (NB * 16);
        e.pad = 0;
        coltab[l * TW + j] = e;
      }
      int c = (int)floor((double)(x0 + hj0) * cs);
      lv_c0[l] = c > gwm1 ? gwm1 : c;
    }
  }

  // ---- ring: rows live at slot (row - row_org) mod NRING; fill(b) brings the rows batch b adds ----
  int64_t row_org = yb0 - RH < 0 ? 0 : yb0 - RH;
  if (row_org < p.dem_row0) row_org = p.dem_row0;
  const int64_t dem_last = p.dem_row0 + p.dem_rows - 1;
  auto need_hi_of = [&](int64_t yy) {
    int64_t v = yy + NB + RH >= H ? H - 1 : yy + NB + RH;
    return v > dem_last ? dem_last : v;
  };
  const int nbatches = (int)((yb1 - yb0 + NB - 1) / NB);
  const unsigned ring_sa = smem_u32(ring);
  unsigned fill_par[2] = {0u, 0u};
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    for (int k = 0; k < 32; ++k) sflag[k] = 0;
  }
  __syncthreads();
  auto issue_fill = [&](int b) {   // vertical role only; rows (hi(b-1), hi(b)], empty past the last batch
    int64_t from = 0, to = -1;
    if (b < nbatches) {
      from = b == 0 ? row_org : need_hi_of(yb0 + (int64_t)(b - 1) * NB) + 1;
      to = need_hi_of(yb0 + (int64_t)b * NB);
    }
    const int n = to >= from ? (int)(to - from + 1) : 0;
    if (bulk) {
      if (vt >= 0 && vt < 32) {
        const unsigned bar = bar0 + 8u * (unsigned)(b & 1);
        if (vt == 0) mbar_expect_tx(bar, (unsigned)n * (unsigned)(SW * 4));
        __syncwarp();
        for (int k = vt; k < n; k += 32) {
          int sl = (int)((from + k - row_org) % NRING);
          bulk_copy_g2s(ring_sa + (unsigned)sl * (unsigned)(RS * 4), p.dem + (from + k - p.dem_row0) * p.ld_in + cs0,
                        (unsigned)(SW * 4), bar);
        }
      }
    } else {
      if (!is_h) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const bool ok = q ? vok1 : vok0;
          if (ok && n > 0) {
            const float* src = p.dem + (from - p.dem_row0) * p.ld_in + (cs0 + vc + q);
            for (int k = 0; k < n; ++k) {
              int sl = (int)((from + k - row_org) % NRING);
              unsigned sa = ring_sa + 4u * (unsigned)(vc + q) + (unsigned)sl * (unsigned)(RS * 4);
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src) : "memory");
              src += p.ld_in;
            }
          }
        }
        cp_async_commit();
      }
    }
  };
  auto wait_fill = [&](int b) {   // all threads; at most one younger fill is outstanding
    if (bulk) {
      mbar_wait(bar0 + 8u * (unsigned)(b & 1), fill_par[b & 1]);
      fill_par[b & 1] ^= 1u;
    } else if (!is_h) {
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    }
  };
  issue_fill(0);
  issue_fill(1);

  const int total = nbatches * nf;
  for (int s = 0; s <= total; ++s) {
    const int bv = s < total ? s / nf : 0, jv = s < total ? s % nf : 0;
    const int slot = s & 3;
    if (s < total && jv == 0) wait_fill(bv);
    // ======================= vertical role: pass s =======================
    if (!is_h && s < total) {
      const int64_t y = yb0 + (int64_t)bv * NB;
      const int base = (int)((y - row_org) % NRING);
      const int nrows_b = (int)((yb1 - y) < NB ? (yb1 - y) : NB);
      const bool interior_rows = (y - RH - 1 >= 0) && (y + NB + RH + 1 < H);
      if (jv == 0) {
        for (int e = vt; e < G::NSLOT; e += V7_THREADS - V7_HT) {
          int d = e - RH - 1;
          int dd = reflect1(y + d, H) - (int)y;
          int sl = (base + dd) % NRING;
          if (sl < 0) sl += NRING;
          slot_tab[e] = sl * RS;
        }
        v7_bar_vrole();
      }
      auto init_col = [&](int r, int col, bool ok, double& sum, int& cnt) {
        sum = 0.0;
        cnt = 0;
        if (ok) {
          double a0 = 0.0, a1 = 0.0;
          for (int d = -r; d <= r; ++d) {
            float v = ring[slot_tab[d + RH + 1] + col];
            bool fin = v == v;
            if (d & 1) a1 += fin ? (double)v : 0.0;
            else a0 += fin ? (double)v : 0.0;
            cnt += fin;
          }
          sum = a0 + a1;
        }
      };
      if (bv == 0 && jv == 0) {
        // window sums at the first row of the band; an incomplete window starts the term in NaN mode
        for (int k = 0; k < nf; ++k) {
          const int r = p.terms[fused_at(k)].r;
          double a, b;
          int ca, cb;
          init_col(r, vc, vok0, a, ca);
          init_col(r, vc + 1, vok1, b, cb);
#pragma unroll
          for (int q = 0; q < V7_MAXF; ++q) if (q == k) { rs0[q] = a; rc0[q] = ca; rs1[q] = b; rc1[q] = cb; }
          if ((vok0 && ca != 2 * r + 1) || (vok1 && cb != 2 * r + 1)) sflag[16 + k] = 1;
        }
        v7_bar_vrole();
        nanmask = 0;
        for (int k = 0; k < nf; ++k) nanmask |= sflag[16 + k] ? (1u << k) : 0u;
      }
      const DevTerm& T = p.terms[fused_at(jv)];
      const int r = T.r;
      const double n = (double)(2 * r + 1), inv = 1.0 / n;
      const bool vactive = vpair && vc + 1 >= RH - r && vc < RH + TW + r;
      const bool nanmode = (nanmask >> jv) & 1u;
      unsigned char* plane_b = smraw + G::OFF_PLANE + (size_t)(s & 1) * G::PLANE_BYTES;
      double* plane64 = reinterpret_cast<double*>(plane_b);
      float* plane32 = reinterpret_cast<float*>(plane_b);
      unsigned char* cplane = plane_b + (size_t)NB * PS * 4;
      double sa = 0.0, sb = 0.0;
      int ca = 0, cb = 0;
#pragma unroll
      for (int q = 0; q < V7_MAXF; ++q) if (q == jv) { sa = rs0[q]; ca = rc0[q]; sb = rs1[q]; cb = rc1[q]; }
      auto nan_pass = [&](int col, bool ok, double& sum, int& cnt, bool& full) {
        if (!ok) return;
        const int* tin = slot_tab + (r + 1 + RH + 1);
        const int* tout = slot_tab + (-r + RH + 1);
        for (int i = 0; i < nrows_b; ++i) {
          plane32[i * PS + col] = (float)div_by_count(sum, n, inv);
          cplane[i * PS + col] = (unsigned char)cnt;
          float vin = ring[tin[i] + col];
          float vout = ring[tout[i] + col];
          bool oin = vin == vin, oout = vout == vout;
          sum += (oin ? (double)vin : 0.0) - (oout ? (double)vout : 0.0);
          cnt += (int)oin - (int)oout;
        }
        full = full && (cnt == 2 * r + 1);
      };
      if (tid == V7_HT) sflag[slot * 4 + 2] = nanmode ? 1 : 0;
      if (!nanmode) {
        if (vactive) {
          if (interior_rows && vok0 && vok1) {
            int sin = base + r + 1; if (sin >= NRING) sin -= NRING;
            int sout = base - r; if (sout < 0) sout += NRING;
            int i = 0;
            while (i < nrows_b) {
              int run = nrows_b - i;
              if (NRING - sin < run) run = NRING - sin;
              if (NRING - sout < run) run = NRING - sout;
              const float* pin = ring + sin * RS + vc;
              const float* pout = ring + sout * RS + vc;
              double* pv = plane64 + i * PS + vc;
              int k = 0;
              for (; k + 2 <= run; k += 2) {
                const float2 i0 = *reinterpret_cast<const float2*>(pin), i1 = *reinterpret_cast<const float2*>(pin + RS);
                const float2 o0 = *reinterpret_cast<const float2*>(pout), o1 = *reinterpret_cast<const float2*>(pout + RS);
                const double da0 = (double)i0.x - (double)o0.x, db0 = (double)i0.y - (double)o0.y;
                const double da1 = (double)i1.x - (double)o1.x, db1 = (double)i1.y - (double)o1.y;
                const double sa1 = sa + da0, sb1 = sb + db0;
                pv[0] = round_to_f32_grid(sa * inv);
                pv[1] = round_to_f32_grid(sb * inv);
                pv[PS] = round_to_f32_grid(sa1 * inv);
                pv[PS + 1] = round_to_f32_grid(sb1 * inv);
                sa = sa1 + da1;
                sb = sb1 + db1;
                pin += 2 * RS; pout += 2 * RS; pv += 2 * PS;
              }
              for (; k < run; ++k) {
                const float2 i0 = *reinterpret_cast<const float2*>(pin);
                const float2 o0 = *reinterpret_cast<const float2*>(pout);
                pv[0] = round_to_f32_grid(sa * inv);
                pv[1] = round_to_f32_grid(sb * inv);
                sa += (double)i0.x - (double)o0.x;
                sb += (double)i0.y - (double)o0.y;
                pin += RS; pout += RS; pv += PS;
              }
              i += run;
              sin += run; if (sin >= NRING) sin -= NRING;
              sout += run; if (sout >= NRING) sout -= NRING;
            }
          } else {
            const int* tin = slot_tab + (r + 1 + RH + 1);
            const int* tout = slot_tab + (-r + RH + 1);
            if (vok0) {
              for (int i = 0; i < nrows_b; ++i) {
                plane64[i * PS + vc] = round_to_f32_grid(sa * inv);
                sa += (double)ring[tin[i] + vc] - (double)ring[tout[i] + vc];
              }
            }
            if (vok1) {
              for (int i = 0; i < nrows_b; ++i) {
                plane64[i * PS + vc + 1] = round_to_f32_grid(sb * inv);
                sb += (double)ring[tin[i] + vc + 1] - (double)ring[tout[i] + vc + 1];
              }
            }
          }
          if ((vok0 && sa != sa) || (vok1 && sb != sb)) sflag[slot * 4 + 0] = 1;   // a NaN entered a window
        }
      } else {
        bool full = true;
        if (vactive) {
          nan_pass(vc, vok0, sa, ca, full);
          nan_pass(vc + 1, vok1, sb, cb, full);
        }
        if (!full) sflag[slot * 4 + 1] = 1;
      }
#pragma unroll
      for (int q = 0; q < V7_MAXF; ++q) if (q == jv) { rs0[q] = sa; rc0[q] = ca; rs1[q] = sb; rc1[q] = cb; }
    }
    // ======================= horizontal role: pass s-1 =======================
    if (is_h && s >= 1) {
      const int bh = (s - 1) / nf, jh = (s - 1) % nf;
      const int64_t y = yb0 + (int64_t)bh * NB;
      const int64_t orow = y + hi;
      const bool hrow_ok = orow < yb1;
      const bool hfull = hrow_ok && !edge_strip && hjn == SEG;
      const int t_this = fused_at(jh);
      const int t_prev = jh == 0 ? -1 : fused_at(jh - 1);
      const int t_next = jh == nf - 1 ? p.n_terms : fused_at(jh + 1);
      if (jh == 0) {
        const float* xrow = ring + (int)((orow - row_org) % NRING) * RS + RH + hj0;
        if (hfull) {
#pragma unroll
          for (int q = 0; q < SEG / 4; ++q) {
            float4 v = *reinterpret_cast<const float4*>(xrow + 4 * q);
            xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < SEG; ++jj) xr[jj] = (hrow_ok && jj < hjn) ? xrow[jj] : 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < SEG; ++jj) acc[jj] = 0.f;
      }
      // terms that are not fused boxes, in list order around the fused term of this pass
      auto other_terms = [&](int t_lo, int t_hi) {
        for (int t = t_lo; t < t_hi; ++t) {
          const DevTerm& T = p.terms[t];
          if (T.kind == TERM_COARSE) {
            if (hrow_ok && hjn > 0) {
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
#pragma unroll
              for (int l = 0; l < V7_MAXLV; ++l) if (l == T.lvl) c0 = lv_c0[l];
              const float wgt = T.weight;
              const V6ColEntry* ce = coltab + T.lvl * TW + hj0;
              if (hjn == SEG) {
                double2* cell = reinterpret_cast<double2*>(smraw + G::OFF_CELL) + (size_t)hg * (V7_KMAX * NB) + hi;
                const int npair = ce[SEG - 1].koff / (NB * 16) + 1;
                int ca = c0;
                double Ak = (double)__ldg(g0 + ca) * wr0 + (double)__ldg(g1 + ca) * tr;
                if (npair <= 3) {
#pragma unroll
                  for (int k = 0; k < 3; ++k) {
                    if (k < npair) {
                      int cb = ca + 1 < gwm1 ? ca + 1 : gwm1;
                      double An = (double)__ldg(g0 + cb) * wr0 + (double)__ldg(g1 + cb) * tr;
                      cell[k * NB] = make_double2(Ak, An - Ak);
                      Ak = An;
                      ca = cb;
                    }
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < V7_KMAX; ++k) {
                    if (k < npair) {
                      int cb = ca + 1 < gwm1 ? ca + 1 : gwm1;
                      double An = (double)__ldg(g0 + cb) * wr0 + (double)__ldg(g1 + cb) * tr;
                      cell[k * NB] = make_double2(Ak, An - Ak);
                      Ak = An;
                      ca = cb;
                    }
                  }
                }
                const unsigned char* cellb = reinterpret_cast<const unsigned char*>(cell);
                // A0 + tc*dA vs scipy's four-tap sum: see fused_kernel_v6 (guard band V6_GUARD f64 ulps)
#pragma unroll
                for (int g = 0; g < SEG; g += V6_CG) {
                  double m64[V6_CG];
                  unsigned risk = 0xffffffffu;
#pragma unroll
                  for (int u = 0; u < V6_CG; ++u) {
                    const V6ColEntry e = ce[g + u];
                    const double2 ad = *reinterpret_cast<const double2*>(cellb + e.koff);
                    m64[u] = fma(e.tc, ad.y, ad.x);
                    const unsigned rk = ((unsigned)__double2loint(m64[u]) + (V6_GUARD - 0x10000000u)) << 3;
                    risk = rk < risk ? rk : risk;
                  }
                  if (risk < (2u * V6_GUARD) << 3) {
#pragma unroll
                    for (int u = 0; u < V6_CG; ++u) {
                      const V6ColEntry e = ce[g + u];
                      int cA = c0 + e.koff / (NB * 16);
                      int cB = cA + 1 < gwm1 ? cA + 1 : gwm1;
                      double p00 = (double)__ldg(g0 + cA) * wr0, p01 = (double)__ldg(g0 + cB) * wr0;
                      double p10 = (double)__ldg(g1 + cA) * tr, p11 = (double)__ldg(g1 + cB) * tr;
                      double wc0 = 1.0 - e.tc;
                      double v = p00 * wc0;
                      v += p01 * e.tc;
                      v += p10 * wc0;
                      v += p11 * e.tc;
                      m64[u] = v;
                    }
                  }
#pragma unroll
                  for (int u = 0; u < V6_CG; ++u) {
                    float mean = (float)m64[u];
                    acc[g + u] = acc[g + u] + wgt * (xr[g + u] - mean);
                  }
                }
              } else {
#pragma unroll 1
                for (int jj = 0; jj < hjn; ++jj) {
                  const V6ColEntry e = ce[jj];
                  int cA = c0 + e.koff / (NB * 16);
                  int cB = cA + 1 < gwm1 ? cA + 1 : gwm1;
                  double p00 = (double)__ldg(g0 + cA) * wr0, p01 = (double)__ldg(g0 + cB) * wr0;
                  double p10 = (double)__ldg(g1 + cA) * tr, p11 = (double)__ldg(g1 + cB) * tr;
                  double wc0 = 1.0 - e.tc;
                  double v = p00 * wc0;
                  v += p01 * e.tc;
                  v += p10 * wc0;
                  v += p11 * e.tc;
                  float mean = (float)v;
#pragma unroll
                  for (int q = 0; q < SEG; ++q)
                    if (q == jj) acc[q] = acc[q] + wgt * (xr[q] - mean);
                }
              }
            }
          } else if (T.kind == TERM_PLANE) {
            if (hrow_ok && hjn > 0) {
              const float* prow = T.grid + (orow - T.grow0) * W;
#pragma unroll
              for (int jj = 0; jj < SEG; ++jj)
                if (jj < hjn) acc[jj] = acc[jj] + T.weight * (xr[jj] - __ldg(prow + x0 + hj0 + jj));
            }
          }
        }
      };
      other_terms(t_prev + 1, t_this);
      {
        const DevTerm& T = p.terms[t_this];
        const int r = T.r;
        const double n = (double)(2 * r + 1), inv = 1.0 / n;
        const bool nanmode = sflag[((s - 1) & 3) * 4 + 2] != 0;
        const unsigned char* plane_b = smraw + G::OFF_PLANE + (size_t)((s - 1) & 1) * G::PLANE_BYTES;
        const double* plane64 = reinterpret_cast<const double*>(plane_b);
        const float* plane32 = reinterpret_cast<const float*>(plane_b);
        const unsigned char* cplane = plane_b + (size_t)NB * PS * 4;
        if (hfull && !nanmode) {
          const double* pl = plane64 + hi * PS + RH + hj0 - r;
          const double* pr = pl + 2 * r + 1;
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          int d = 0;
          for (; d + 4 <= 2 * r + 1; d += 4) { s0 += pl[d]; s1 += pl[d + 1]; s2 += pl[d + 2]; s3 += pl[d + 3]; }
          for (; d < 2 * r + 1; ++d) s0 += pl[d];
          double sv = (s0 + s1) + (s2 + s3);
          const float wgt = T.weight;
#pragma unroll
          for (int jj = 0; jj < SEG; ++jj) {
            float mean = (float)(sv * inv);
            acc[jj] = acc[jj] + wgt * (xr[jj] - mean);
            if (jj + 1 < SEG) sv += pr[jj] - pl[jj];
          }
        } else if (hrow_ok && hjn > 0) {
          const unsigned char* crow = cplane + hi * PS;
          if (nanmode) {
            const float* vrow = plane32 + hi * PS;
            if (edge_strip) hphase_generic<true, SEG, float>(vrow, crow, xr, r, hj0, hjn, x0, cs0, W, true, T.weight, acc);
            else hphase_generic<false, SEG, float>(vrow, crow, xr, r, hj0, hjn, x0, cs0, W, true, T.weight, acc);
          } else {
            const double* vrow = plane64 + hi * PS;
            if (edge_strip) hphase_generic<true, SEG, double>(vrow, crow, xr, r, hj0, hjn, x0, cs0, W, false, T.weight, acc);
            else hphase_generic<false, SEG, double>(vrow, crow, xr, r, hj0, hjn, x0, cs0, W, false, T.weight, acc);
          }
        }
      }
      other_terms(t_this + 1, t_next);
      if (jh == nf - 1 && hrow_ok && hjn > 0) {
        // ---- epilogue: normalise (correctly rounded v / s), encode, store from registers ----
        if (p.norm_mode == 1) {
          const float sc = p.norm_scale, rinv = p.norm_rinv;
#pragma unroll
          for (int jj = 0; jj < SEG; ++jj) {
            float q = acc[jj] * rinv;
            float rem = fmaf(-q, sc, acc[jj]);
            acc[jj] = fmaf(rem, rinv, q);
          }
        } else if (p.norm_mode == 2) {
#pragma unroll
          for (int jj = 0; jj < SEG; ++jj) acc[jj] = (acc[jj] != acc[jj]) ? acc[jj] : 0.f;
        }
        const int64_t obase = (orow - p.out_row0) * p.ld_out + x0 + hj0;
        if (hjn == SEG && p.out_vec_ok && p.enc.kind == FSG_OUT_F32) {
          float4* o = reinterpret_cast<float4*>((float*)p.out + obase);
#pragma unroll
          for (int q = 0; q < SEG / 4; ++q) o[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        } else if (hjn == SEG && p.out_vec_ok && p.enc.kind == FSG_OUT_U8) {
          unsigned* o = reinterpret_cast<unsigned*>((uint8_t*)p.out + obase);
#pragma unroll
          for (int q = 0; q < SEG / 4; ++q) {
            unsigned b0 = (unsigned)(int)encode_dn(acc[4 * q], p.enc), b1 = (unsigned)(int)encode_dn(acc[4 * q + 1], p.enc);
            unsigned b2 = (unsigned)(int)encode_dn(acc[4 * q + 2], p.enc), b3 = (unsigned)(int)encode_dn(acc[4 * q + 3], p.enc);
            o[q] = (b0 & 255u) | ((b1 & 255u) << 8) | ((b2 & 255u) << 16) | (b3 << 24);
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < SEG; ++jj)
            if (jj < hjn) store_out(p.out, obase + jj, acc[jj], p.enc);
        }
      }
    }
    __syncthreads();
    // ======================= step end (uniform) =======================
    if (s < total) {
      const bool bad = sflag[slot * 4 + 0] != 0;
      if (bad) {
        // some window of this radius met a NaN during the optimistic dense pass: redo it NaN-aware
        nanmask |= 1u << jv;
        if (!is_h) {
          const int64_t y = yb0 + (int64_t)bv * NB;
          const int nrows_b = (int)((yb1 - y) < NB ? (yb1 - y) : NB);
          const DevTerm& T = p.terms[fused_at(jv)];
          const int r = T.r;
          const double n = (double)(2 * r + 1), inv = 1.0 / n;
          const bool vactive = vpair && vc + 1 >= RH - r && vc < RH + TW + r;
          unsigned char* plane_b = smraw + G::OFF_PLANE + (size_t)(s & 1) * G::PLANE_BYTES;
          float* plane32 = reinterpret_cast<float*>(plane_b);
          unsigned char* cplane = plane_b + (size_t)NB * PS * 4;
          double sum[2];
          int cnt[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const bool ok = q ? vok1 : vok0;
            const int col = vc + q;
            sum[q] = 0.0;
            cnt[q] = 0;
            if (ok) {
              double a0 = 0.0, a1 = 0.0;
              for (int d = -r; d <= r; ++d) {
                float v = ring[slot_tab[d + RH + 1] + col];
                bool fin = v == v;
                if (d & 1) a1 += fin ? (double)v : 0.0;
                else a0 += fin ? (double)v : 0.0;
                cnt[q] += fin;
              }
              sum[q] = a0 + a1;
              if (vactive) {
                const int* tin = slot_tab + (r + 1 + RH + 1);
                const int* tout = slot_tab + (-r + RH + 1);
                for (int i = 0; i < nrows_b; ++i) {
                  plane32[i * PS + col] = (float)div_by_count(sum[q], n, inv);
                  cplane[i * PS + col] = (unsigned char)cnt[q];
                  float vin = ring[tin[i] + col];
                  float vout = ring[tout[i] + col];
                  bool oin = vin == vin, oout = vout == vout;
                  sum[q] += (oin ? (double)vin : 0.0) - (oout ? (double)vout : 0.0);
                  cnt[q] += (int)oin - (int)oout;
                }
              }
            }
          }
#pragma unroll
          for (int q = 0; q < V7_MAXF; ++q)
            if (q == jv) { rs0[q] = sum[0]; rc0[q] = cnt[0]; rs1[q] = sum[1]; rc1[q] = cnt[1]; }
          if (tid == V7_HT) sflag[slot * 4 + 2] = 1;
        }
        __syncthreads();
      } else if (((nanmask >> jv) & 1u) && sflag[slot * 4 + 1] == 0) {
        nanmask &= ~(1u << jv);   // every window complete again at the end of the batch: next batch runs dense
      }
      if (jv == nf - 1) issue_fill(bv + 2);
    }
    // flags of the previous step are no longer read by anyone: clear them for step s + 3
    if (tid == 0 && s >= 1) { sflag[((s - 1) & 3) * 4 + 0] = 0; sflag[((s - 1) & 3) * 4 + 1] = 0; }
  }
  // drain the copies that are still in flight (issued past the last batch they are empty)
  if (bulk) {
    mbar_wait(bar0 + 8u * (unsigned)(nbatches & 1), fill_par[nbatches & 1]);
    mbar_wait(bar0 + 8u * (unsigned)((nbatches + 1) & 1), fill_par[(nbatches + 1) & 1]);
  } else if (!is_h) {
    cp_async_wait_all();
  }
}

}  // namespace fsg
