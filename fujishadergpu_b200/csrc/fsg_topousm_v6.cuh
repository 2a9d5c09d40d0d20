// fused_kernel_v6: leaner full-resolution pass of topousm_fast (same arithmetic as fused_kernel /
// fused_kernel_fast -- reference: compute_topousm_fast_efficient_block, _impl_topousm_fast.py:49-100;
// box means: handle_nan_with_uniform, _nan_utils.py:34-47; zoom: _upsample_to_shape, :671-698).
//
// What changed against fused_kernel_fast (the ncu profile of round 1 showed ~240 thread-instructions per
// pixel, half of them address arithmetic / predication, at 11 warps per SM):
//   * compile-time strides (ring halo RH is a template parameter) -> immediate-offset LDS/STS;
//   * ring fill by 1-D bulk async copies (cp.async.bulk + mbarrier, one instruction per row) for
//     interior strips whose rows are 16-byte aligned; LDGSTS per column otherwise;
//   * no NaN probe: the dense vertical pass runs optimistically and a NaN shows up in the running
//     window sum (sticky); the batch is then redone for that radius in the NaN-aware form;
//   * the coarse (decimated) terms no longer carry per-pixel "advance" predication: each thread first
//     row-interpolates the few coarse cells its 24-pixel segment touches into a private shared-memory
//     slot (the plane region is idle then), and the pixel loop is branch-free;
//   * window-sum initialisation with four independent accumulators (the sums are exact, so the
//     association does not matter); 128-bit loads of the centre pixels.
// Included by fsg_topousm.cu (needs FusedParams, DevTerm, hphase_generic, coarse_taps, reflect1 ...).
#pragma once

namespace fsg {

constexpr int V6_NB = 32;                     // rows per batch (lanes of a warp = the 32 rows in the horizontal pass)
constexpr int V6_MAXF = 4;                    // fused radii with register-resident window sums
constexpr int V6_MAXLV = 3;                   // pyramid levels with a column table in shared memory
constexpr int V6_CG = 8;                      // coarse terms: pixels per rounding-boundary check (v7)
constexpr unsigned V6_GUARD = 1024u;          // ... and the guard band in f64 ulps (see the coarse term)

// Launch geometry: strip width TW, pixels per horizontal thread SEG (threads = 32 rows x TW/SEG segments),
// smallest decimation factor the coarse cell slots are sized for.
//   V6CfgA: 264 columns, 24-pixel segments, 352 threads (11 warps, 168 registers)
//   V6CfgB: 240 columns, 12-pixel segments, 640 threads (20 warps, <= 102 registers): more warps to hide
//           latency, at the price of twice the window-start work per pixel
struct V6CfgA { static constexpr int TW = 264, SEG = 24, FMIN = 2; };
struct V6CfgB { static constexpr int TW = 240, SEG = 12, FMIN = 4; };

struct V6ColEntry {   // per (level, strip column): zoom fraction and byte offset of the coarse cell slot
  double tc;
  int koff;           // (coarse column - first coarse column of the segment) * 32 lanes * 16 bytes
  int pad;
};

template <int RH, typename CFG>
struct V6Geom {
  static constexpr int TW = CFG::TW, SEG = CFG::SEG, NSEG = TW / SEG, THREADS = NSEG * V6_NB;
  static constexpr int KMAX = SEG / CFG::FMIN + 2;   // coarse cells a segment can touch
  static_assert(SEG * NSEG == TW && SEG % 4 == 0, "segments must tile the strip");
  static constexpr int SW = TW + 2 * RH;                       // strip width incl. halo
  static constexpr int RS = ((SW / 4) % 2 == 1) ? SW : SW + 4;    // ring row stride (floats): RS/4 odd -> LDS.128 by rows is conflict-free
  static constexpr int PS = SW | 1;                               // plane row stride (doubles): odd -> LDS.64 by rows is conflict-free
  static constexpr int NRING = V6_NB + 2 * RH + 1;
  static constexpr size_t OFF_PLANE = ((size_t)NRING * RS * 4 + 15) / 16 * 16;
  static constexpr size_t PLANE_BYTES = (size_t)V6_NB * PS * 8;
  static constexpr size_t OFF_SLOT = OFF_PLANE + PLANE_BYTES;
  static constexpr size_t OFF_TC = (OFF_SLOT + (size_t)(NRING + 1) * 4 + 15) / 16 * 16;
  static constexpr size_t OFF_BAR = OFF_TC + (size_t)V6_MAXLV * TW * sizeof(V6ColEntry);
  static constexpr size_t BYTES = OFF_BAR + 16;
  static_assert(THREADS >= SW, "one vertical thread per strip column");
  static_assert((size_t)NSEG * KMAX * 32 * 16 <= PLANE_BYTES, "coarse cell slots must fit in the plane region");
  static_assert(RS % 4 == 0 && SW % 4 == 0, "rows must be whole 16-byte groups");
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "V6_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra V6_DONE;\n"
      "bra V6_WAIT;\n"
      "V6_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// f64 -> nearest f32-representable value, kept as f64 (scipy's f32 store after axis 0)
// (magic-number form on the FP64 pipe: conversions run on the XU at 16 lanes/clk/SM and are the
// scarce resource of this kernel)
__device__ __forceinline__ double v6_round_f32(double q) { return round_to_f32_grid(q); }

template <int RH, typename CFG>
__global__ void __launch_bounds__(V6Geom<RH, CFG>::THREADS, 1) fused_kernel_v6(FusedParams p) {
  using G = V6Geom<RH, CFG>;
  constexpr int NB = V6_NB, SEG = G::SEG, SW = G::SW, RS = G::RS, PS = G::PS, NRING = G::NRING;
  constexpr int FK_TW = G::TW, FK_THREADS = G::THREADS, V6_KMAX = G::KMAX;   // (shadow the v5 constants)
  extern __shared__ __align__(16) unsigned char smraw[];
  const NormDev nd = resolve_norm(p);
  float* ring = reinterpret_cast<float*>(smraw);
  double* plane64 = reinterpret_cast<double*>(smraw + G::OFF_PLANE);
  float* plane32 = reinterpret_cast<float*>(smraw + G::OFF_PLANE);
  unsigned char* cplane = smraw + G::OFF_PLANE + (size_t)NB * PS * 4;
  int* slot_tab = reinterpret_cast<int*>(smraw + G::OFF_SLOT);
  V6ColEntry* coltab = reinterpret_cast<V6ColEntry*>(smraw + G::OFF_TC);
  const unsigned bar = smem_u32(smraw + G::OFF_BAR);

  const int tid = threadIdx.x;
  const int W = (int)p.W;
  const int64_t H = p.H;
  int x0 = p.col0 + ((int)blockIdx.x + p.strip0) * FK_TW;
  int Wc = p.col_end > 0 ? p.col_end : W;   // output columns stop here (rectangle launches)
  const int64_t out_end = p.out_row0 + p.out_rows;
  int64_t yb0 = p.out_row0 + (int64_t)blockIdx.y * p.band_rows;
  int64_t yb1 = (yb0 + p.band_rows < out_end) ? yb0 + p.band_rows : out_end;
  if (p.tile_flags) {   // flagged tiles only (blocks the interior fast path gave up on)
    if (p.tile_flags[(size_t)blockIdx.y * gridDim.x + blockIdx.x] == 0) return;
    x0 = p.col0 + (int)blockIdx.x * p.tile_w;
    if (x0 + p.tile_w < Wc) Wc = x0 + p.tile_w;
    yb0 = p.tile_row0 + (int64_t)blockIdx.y * p.tile_rows;
    yb1 = (yb0 + p.tile_rows < p.tile_row1) ? yb0 + p.tile_rows : p.tile_row1;
  }
  const int cs0 = x0 - RH;
  const bool edge_strip = (cs0 < 0) || (cs0 + SW > W);
  const bool bulk = p.bulk_ok && !edge_strip;

  const int vc = tid;
  const int vgx = cs0 + vc;
  const bool vcol_ok = vc < SW && vgx >= 0 && vgx < W;

  double rs[V6_MAXF];
  int rc[V6_MAXF];
#pragma unroll
  for (int k = 0; k < V6_MAXF; ++k) { rs[k] = 0.0; rc[k] = 0; }
  int last_fused = -1;
  for (int t = 0; t < p.n_terms; ++t)
    if (p.terms[t].kind == TERM_BOX_FUSED) last_fused = t;

  const int hi = tid % NB;     // row of the batch (lanes of a warp = the 32 rows)
  const int hg = tid / NB;     // column segment
  const int hj0 = hg * SEG;
  int hjn = SEG;
  if (x0 + hj0 + hjn > Wc) hjn = (Wc - x0 - hj0 > 0) ? (Wc - x0 - hj0) : 0;

  // Column tables of the coarse levels (same for every row of the strip) and the first coarse column of
  // this thread's segment.
  int lv_c0[V6_MAXLV];
#pragma unroll
  for (int l = 0; l < V6_MAXLV; ++l) {
    lv_c0[l] = 0;
    if (l < p.n_lvls) {
      const double cs = p.lvl_cscale[l];
      const int gwm1 = p.lvl_gw[l] - 1;
      for (int j = tid; j < FK_TW; j += FK_THREADS) {
        double ci = (double)(x0 + j) * cs;
        double fl = floor(ci);
        if (fl > (double)gwm1) fl = (double)gwm1;
        int cseg = (int)floor((double)(x0 + (j / SEG) * SEG) * cs);
        if (cseg > gwm1) cseg = gwm1;
        V6ColEntry e;
        e.tc = ci - fl;
        e.koff = ((int)fl - cseg) * (32 * 16);
        e.pad = 0;
        coltab[l * FK_TW + j] = e;
      }
      int c = (int)floor((double)(x0 + hj0) * cs);
      lv_c0[l] = c > gwm1 ? gwm1 : c;
    }
  }

  // rows live at ring slot (row - row_org) mod NRING
  int64_t row_org = yb0 - RH < 0 ? 0 : yb0 - RH;
  if (row_org < p.dem_row0) row_org = p.dem_row0;
  const int64_t dem_last = p.dem_row0 + p.dem_rows - 1;   // last row the caller's buffer holds
  auto need_hi_of = [&](int64_t yy) {
    int64_t v = yy + NB + RH >= H ? H - 1 : yy + NB + RH;
    return v > dem_last ? dem_last : v;
  };
  const unsigned ring_sa = smem_u32(ring);
  unsigned fill_parity = 0;
  if (bulk) {
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
  }
  auto issue_rows = [&](int64_t from, int64_t to) {
    if (bulk) {
      if (tid < 32) {
        const int n = to >= from ? (int)(to - from + 1) : 0;
        if (tid == 0) mbar_expect_tx(bar, (unsigned)n * (unsigned)(SW * 4));
        __syncwarp();
        for (int k = tid; k < n; k += 32) {
          int sl = (int)((from + k - row_org) % NRING);
          bulk_copy_g2s(ring_sa + (unsigned)sl * (unsigned)(RS * 4), p.dem + (from + k - p.dem_row0) * p.ld_in + cs0,
                        (unsigned)(SW * 4), bar);
        }
      }
    } else {
      if (vcol_ok && from <= to) {
        int sl = (int)((from - row_org) % NRING);
        int n = (int)(to - from + 1);
        const float* src = p.dem + (from - p.dem_row0) * p.ld_in + vgx;
        while (n > 0) {
          int run = NRING - sl < n ? NRING - sl : n;   // rows until the ring wraps
          unsigned sa = ring_sa + 4u * (unsigned)vc + (unsigned)sl * (unsigned)(RS * 4);
#pragma unroll 4
          for (int k = 0; k < run; ++k) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src) : "memory");
            sa += RS * 4;
            src += p.ld_in;
          }
          n -= run;
          sl += run;
          if (sl >= NRING) sl = 0;
        }
      }
      cp_async_commit();
    }
  };
  auto wait_rows = [&]() {
    if (bulk) {
      mbar_wait(bar, fill_parity);
      fill_parity ^= 1u;
    } else {
      cp_async_wait_all();
    }
  };

  unsigned nanmask = 0;   // bit fk: fused term fk is in the NaN-aware form (CTA-uniform)
  int base = (int)(yb0 - row_org);  // ring slot of row y
  issue_rows(row_org, need_hi_of(yb0));

  for (int64_t y = yb0; y < yb1; y += NB) {
    const int64_t need_hi = need_hi_of(y);
    wait_rows();
    // slot table: entry (d + RH + 1) = RS * ring slot of (mirrored) row y + d, d in [-RH-1, NB+RH]
    for (int e = tid; e <= NRING; e += FK_THREADS) {
      int d = e - RH - 1;
      int dd = reflect1(y + d, H) - (int)y;
      int sl = (base + dd) % NRING;
      if (sl < 0) sl += NRING;
      slot_tab[e] = sl * RS;
    }
    __syncthreads();   // this batch's rows and the slot table are visible
    const int nrows_b = (int)((yb1 - y) < NB ? (yb1 - y) : NB);
    const bool interior_rows = (y - RH - 1 >= 0) && (y + NB + RH + 1 < H);
    const bool more = y + NB < yb1;

    // (re)initialise the window sums of one fused term at the first row of this batch
    auto init_term = [&](int r, double& s, int& c) {
      s = 0.0;
      c = 0;
      if (vcol_ok) {
        double s0 = 0.0, s1 = 0.0;
        for (int d = -r; d <= r; ++d) {
          float v = ring[slot_tab[d + RH + 1] + vc];
          bool ok = v == v;
          if (d & 1) s1 += ok ? (double)v : 0.0;
          else s0 += ok ? (double)v : 0.0;
          c += ok;
        }
        s = s0 + s1;
      }
    };
    if (y == yb0) {
      int fk = 0;
      unsigned mynan = 0;
      for (int t = 0; t < p.n_terms; ++t) {
        if (p.terms[t].kind != TERM_BOX_FUSED) continue;
        double s;
        int c;
        init_term(p.terms[t].r, s, c);
#pragma unroll
        for (int k = 0; k < V6_MAXF; ++k) if (k == fk) { rs[k] = s; rc[k] = c; }
        if (vcol_ok && c != 2 * p.terms[t].r + 1) mynan |= 1u << fk;
        ++fk;
      }
      // a column that starts with an incomplete window puts its term into the NaN-aware form
      nanmask = 0;
      for (int k = 0; k < fk; ++k) nanmask |= __syncthreads_or((int)((mynan >> k) & 1u)) ? (1u << k) : 0u;
    }

    const int64_t orow = y + hi;
    const bool hrow_ok = orow < yb1;
    const bool hfull = hrow_ok && !edge_strip && hjn == SEG;
    float acc[SEG], xr[SEG];
    {
      const float* xrow = ring + slot_tab[hi + RH + 1] + RH + hj0;
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
    if (last_fused < 0) {   // no vertical pass will read the ring: the next rows can start right away
      __syncthreads();
      issue_rows(need_hi + 1, more ? need_hi_of(y + NB) : need_hi);
    }
    bool plane_dirty = false;   // coarse cell slots were written since the last barrier
    int fk = 0;
    for (int t = 0; t < p.n_terms; ++t) {
      const DevTerm& T = p.terms[t];
      if (T.kind == TERM_BOX_FUSED) {
        const int r = T.r;
        const double n = (double)(2 * r + 1), inv = 1.0 / n;
        const bool vactive = vcol_ok && vc >= RH - r && vc < RH + FK_TW + r;
        if (plane_dirty) { __syncthreads(); plane_dirty = false; }
        double s = 0.0;
        int c = 0;
#pragma unroll
        for (int k = 0; k < V6_MAXF; ++k) if (k == fk) { s = rs[k]; c = rc[k]; }
        bool nanmode = (nanmask >> fk) & 1u;
        if (!nanmode) {
          // ---- dense vertical pass (optimistic: a NaN poisons s and is detected below) ----
          if (vactive) {
            if (interior_rows) {
              // rows are consecutive ring slots: pointer increments with compile-time strides, the batch
              // is split where the incoming or the outgoing row wraps around the ring
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
                for (; k + 4 <= run; k += 4) {
                  float a0 = pin[0], a1 = pin[RS], a2 = pin[2 * RS], a3 = pin[3 * RS];
                  float b0 = pout[0], b1 = pout[RS], b2 = pout[2 * RS], b3 = pout[3 * RS];
                  double d0 = (double)a0 - (double)b0, d1 = (double)a1 - (double)b1;
                  double d2 = (double)a2 - (double)b2, d3 = (double)a3 - (double)b3;
                  double s1 = s + d0, s2 = s1 + d1, s3 = s2 + d2;
                  pv[0] = v6_round_f32(s * inv);
                  pv[PS] = v6_round_f32(s1 * inv);
                  pv[2 * PS] = v6_round_f32(s2 * inv);
                  pv[3 * PS] = v6_round_f32(s3 * inv);
                  s = s3 + d3;
                  pin += 4 * RS; pout += 4 * RS; pv += 4 * PS;
                }
                for (; k < run; ++k) {
                  pv[0] = v6_round_f32(s * inv);
                  s += (double)pin[0] - (double)pout[0];
                  pin += RS; pout += RS; pv += PS;
                }
                i += run;
                sin += run; if (sin >= NRING) sin -= NRING;
                sout += run; if (sout >= NRING) sout -= NRING;
              }
            } else {
              const int* tin = slot_tab + (r + 1 + RH + 1);
              const int* tout = slot_tab + (-r + RH + 1);
              for (int i = 0; i < nrows_b; ++i) {
                plane64[i * PS + vc] = v6_round_f32(s * inv);
                s += (double)ring[tin[i] + vc] - (double)ring[tout[i] + vc];
              }
            }
          }
          const int bad = __syncthreads_or((int)(vactive && (s != s)));
          if (bad) {   // some window of this radius met a NaN: redo the batch in the NaN-aware form
            nanmode = true;
            nanmask |= 1u << fk;
            init_term(r, s, c);
            __syncthreads();   // every thread has left the plane it wrote optimistically
          }
        }
        if (nanmode) {
          // ---- NaN-aware vertical pass (value sum over valid rows + count) ----
          bool full = true;
          if (vactive) {
            const int* tin = slot_tab + (r + 1 + RH + 1);
            const int* tout = slot_tab + (-r + RH + 1);
            for (int i = 0; i < nrows_b; ++i) {
              plane32[i * PS + vc] = (float)div_by_count(s, n, inv);
              cplane[i * PS + vc] = (unsigned char)c;
              float vin = ring[tin[i] + vc];
              float vout = ring[tout[i] + vc];
              bool oin = vin == vin, oout = vout == vout;
              s += (oin ? (double)vin : 0.0) - (oout ? (double)vout : 0.0);
              c += (int)oin - (int)oout;
            }
            full = (c == 2 * r + 1);
          }
          // all windows complete again at the end of the batch: the next batch may run dense
          if (__syncthreads_and((int)full)) nanmask &= ~(1u << fk);
        }
#pragma unroll
        for (int k = 0; k < V6_MAXF; ++k) if (k == fk) { rs[k] = s; rc[k] = c; }
        if (t == last_fused) {   // the ring rows this batch no longer needs are free: start the next copy
          issue_rows(need_hi + 1, more ? need_hi_of(y + NB) : need_hi);
        }
        // ---- horizontal pass ----
        if (hfull && !nanmode) {
          const double* pl = plane64 + hi * PS + RH + hj0 - r;   // leftmost tap of the first window
          const double* pr = pl + 2 * r + 1;                     // first tap entering
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
        __syncthreads();
        ++fk;
      } else if (T.kind == TERM_COARSE) {
        if (hrow_ok && hjn > 0) {
          // scipy zoom(order=1): coordinate = index*(n_in-1)/(n_out-1), weights [1-t, t], 'nearest' edge.
          // Rows are interpolated first (A = a0*(1-tr) + a1*tr at the coarse columns), then one FMA
          // along the row: A0 + tc*(A1 - A0).
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
          for (int l = 0; l < V6_MAXLV; ++l) if (l == T.lvl) c0 = lv_c0[l];
          const float wgt = T.weight;
          if (hjn == SEG) {
            // private cell slots: (A_k, A_{k+1} - A_k) for the coarse columns this segment touches
            double2* cell = reinterpret_cast<double2*>(smraw + G::OFF_PLANE) + (size_t)hg * (V6_KMAX * 32) + hi;
            const V6ColEntry* ce = coltab + T.lvl * FK_TW + hj0;
            const int npair = ce[SEG - 1].koff / (32 * 16) + 1;
            int ca = c0;
            double Ak = (double)__ldg(g0 + ca) * wr0 + (double)__ldg(g1 + ca) * tr;
            if (npair <= 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < npair) {
                  int cb = ca + 1 < gwm1 ? ca + 1 : gwm1;
                  double An = (double)__ldg(g0 + cb) * wr0 + (double)__ldg(g1 + cb) * tr;
                  cell[k * 32] = make_double2(Ak, An - Ak);
                  Ak = An;
                  ca = cb;
                }
              }
            } else {
#pragma unroll
              for (int k = 0; k < V6_KMAX; ++k) {
                if (k < npair) {
                  int cb = ca + 1 < gwm1 ? ca + 1 : gwm1;
                  double An = (double)__ldg(g0 + cb) * wr0 + (double)__ldg(g1 + cb) * tr;
                  cell[k * 32] = make_double2(Ak, An - Ak);
                  Ak = An;
                  ca = cb;
                }
              }
            }
            const unsigned char* cellb = reinterpret_cast<const unsigned char*>(cell);
            // A0 + tc*dA agrees with scipy's four-tap sum ((p00*wc0 + p01*tc) + p10*wc0) + p11*tc to a few
            // f64 ulps, so the two differ after the f32 rounding only when the value sits within V6_GUARD
            // ulps of a rounding boundary (bit pattern x..x1000...0 in the low 29 bits).  Those pixels
            // (about 4e-6 of them) are recomputed with the four-tap form, group by group.
#pragma unroll
            constexpr int CG = SEG % 8 == 0 ? 8 : 4;   // pixels per rounding-boundary check
#pragma unroll
            for (int g = 0; g < SEG; g += CG) {
              double m64[CG];
              unsigned risk = 0xffffffffu;
#pragma unroll
              for (int u = 0; u < CG; ++u) {
                const V6ColEntry e = ce[g + u];
                const double2 ad = *reinterpret_cast<const double2*>(cellb + e.koff);
                m64[u] = fma(e.tc, ad.y, ad.x);
                const unsigned rk = ((unsigned)__double2loint(m64[u]) + (V6_GUARD - 0x10000000u)) << 3;
                risk = rk < risk ? rk : risk;
              }
              if (risk < (2u * V6_GUARD) << 3) {
#pragma unroll
                for (int u = 0; u < CG; ++u) {
                  const V6ColEntry e = ce[g + u];
                  int ca = c0 + e.koff / (32 * 16);
                  int cb = ca + 1 < gwm1 ? ca + 1 : gwm1;
                  double p00 = (double)__ldg(g0 + ca) * wr0, p01 = (double)__ldg(g0 + cb) * wr0;
                  double p10 = (double)__ldg(g1 + ca) * tr, p11 = (double)__ldg(g1 + cb) * tr;
                  double wc0 = 1.0 - e.tc;
                  double v = p00 * wc0;
                  v += p01 * e.tc;
                  v += p10 * wc0;
                  v += p11 * e.tc;
                  m64[u] = v;
                }
              }
#pragma unroll
              for (int u = 0; u < CG; ++u) {
                float mean = (float)m64[u];
                acc[g + u] = acc[g + u] + wgt * (xr[g + u] - mean);
              }
            }
          } else {
            // partial segment (last strip): scipy's four-tap sum directly
            const V6ColEntry* ce = coltab + T.lvl * FK_TW + hj0;
#pragma unroll 1
            for (int jj = 0; jj < hjn; ++jj) {
              const V6ColEntry e = ce[jj];
              int ca = c0 + e.koff / (32 * 16);
              int cb = ca + 1 < gwm1 ? ca + 1 : gwm1;
              double p00 = (double)__ldg(g0 + ca) * wr0, p01 = (double)__ldg(g0 + cb) * wr0;
              double p10 = (double)__ldg(g1 + ca) * tr, p11 = (double)__ldg(g1 + cb) * tr;
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
        plane_dirty = true;   // CTA-uniform: some thread may have used its cell slots
      } else {
        if (hrow_ok && hjn > 0) {
          const float* prow = T.grid + (orow - T.grow0) * W;
#pragma unroll
          for (int jj = 0; jj < SEG; ++jj)
            if (jj < hjn) acc[jj] = acc[jj] + T.weight * (xr[jj] - __ldg(prow + x0 + hj0 + jj));
        }
      }
    }

    if (plane_dirty) __syncthreads();   // coarse cell slots alias the staging area
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
      const int ncols = (Wc - x0) < FK_TW ? (Wc - x0) : FK_TW;
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
  wait_rows();
}

}  // namespace fsg
