// Exact percentile of a pooled sample (np.percentile, method='linear') without sorting:
// 3-level radix select on the order-preserving integer image of the f32 values.
//
// Reference: topousm_fast_stat_func (algorithms/_normalization.py:22-32): percentile(|x[~isnan]|, 99)
//            robust_unsigned_stretch_stat_func (algorithms/_global_stats.py:181-203): p1 / p99 of finite x
// NumPy evaluates the virtual index (n-1)*q/100 in the ARRAY dtype (f32 here), so the host layer first
// asks for the sample count (rank < 0), derives k with NumPy's own scalar arithmetic, then asks for the
// order statistics a[k], a[k+1]; the two-point interpolation is likewise done host-side in f32.
#include "fsg_common.cuh"

namespace fsg {

constexpr int MAX_CHUNKS = 16;

struct Chunks {
  const float* ptr[MAX_CHUNKS];
  int64_t rows[MAX_CHUNKS], cols[MAX_CHUNKS], ld[MAX_CHUNKS];
  int64_t start[MAX_CHUNKS + 1];  // prefix of rows*cols
  int64_t rowstart[MAX_CHUNKS + 1];  // prefix of rows * ceil(cols / SCAN_TILE): (row, column tile) work items
  int n;
};

struct SelState {
  unsigned long long n;       // contributing samples
  unsigned long long rank;    // remaining rank inside the current prefix bucket
  unsigned int prefix;        // key bits fixed so far
  unsigned int mask;          // which bits are fixed
  unsigned int key_k;         // key of a[k]
  unsigned long long cnt_le;  // #keys <= key_k
  unsigned int key_next;      // min key > key_k
  double q;
  double out[4];              // a[k], a[k+1], gamma, n
  // compaction after level 0 (fsg_select_compact): the keys of the selected level-0 bucket, local to this rank
  unsigned long long cmp_n;       // keys in the compact buffer
  unsigned long long cmp_below;   // local keys below the bucket
  unsigned int cmp_above_min;     // smallest local key above the bucket (0xffffffff: none)
  unsigned int cmp_overflow;      // the buffer was too small (cannot happen with capacity = sample count)
};

__device__ __forceinline__ bool sample_key(float v, int take_abs, int finite_only, unsigned int* key) {
  if (finite_only ? !isfinite(v) : (v != v)) return false;
  unsigned int b = __float_as_uint(v);
  if (take_abs) b &= 0x7fffffffu;
  else b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  *key = b;
  return true;
}
__device__ __forceinline__ float key_value(unsigned int key, int take_abs) {
  if (take_abs) return __uint_as_float(key);
  unsigned int b = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
  return __uint_as_float(b);
}


// Visit every sample: CTAs stride over the rows of all chunks, the threads of a CTA over the columns of
// a row (coalesced, no per-element division).  f(value, in_range) is called by ALL threads of the CTA
// the same number of times, so warp-collective code is allowed inside.
constexpr int SCAN_TILE = 2048;   // columns per (row, tile) work item
template <typename F>
__device__ __forceinline__ void scan_chunks(const Chunks& c, F f) {
  const int64_t ntiles = c.rowstart[c.n];   // prefix of rows * tiles-per-row (see fill_chunks)
  for (int64_t T = blockIdx.x; T < ntiles; T += gridDim.x) {
    int k = 0;
    while (k + 1 < c.n && T >= c.rowstart[k + 1]) ++k;
    const int64_t cols = c.cols[k];
    const int64_t tpr = (cols + SCAN_TILE - 1) / SCAN_TILE;
    const int64_t t = T - c.rowstart[k];
    const int64_t r = t / tpr, cb = (t - r * tpr) * SCAN_TILE;
    const float* row = c.ptr[k] + r * c.ld[k];
    const int64_t cend = cb + SCAN_TILE < cols ? cb + SCAN_TILE : cols;
    constexpr int NL = SCAN_TILE / 256;      // all loads of the tile in flight before the first use
    float v[NL];
#pragma unroll
    for (int q = 0; q < NL; ++q) {
      const int64_t x = cb + (int64_t)q * 256 + threadIdx.x;
      v[q] = x < cend ? row[x] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < NL; ++q) {
      const int64_t x = cb + (int64_t)q * 256 + threadIdx.x;
      f(v[q], x < cend);
    }
  }
}

// shared-memory histogram increment aggregated per warp: the keys of |x| cluster in a handful of bins
// (same exponent), where plain atomics serialise 32-fold
__device__ __forceinline__ void hist_add(unsigned int* sh, unsigned int bin, bool contrib) {
  unsigned int todo = __ballot_sync(0xffffffffu, contrib);
  if (todo == 0u) return;   // levels 1 and 2: most warps hold no key of the selected bucket
  const int lane = (int)(threadIdx.x & 31);
  // two rounds of "the first pending lane collects everyone with its bin" (two ballots and a shuffle each -- much
  // cheaper than match.any), the few lanes left use their own atomic
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    const int leader = __ffs(todo) - 1;
    const unsigned int b0 = __shfl_sync(0xffffffffu, bin, leader);
    const bool mine = contrib && bin == b0;
    const unsigned int same = __ballot_sync(0xffffffffu, mine);
    if (lane == leader) atomicAdd(&sh[b0], (unsigned int)__popc(same));
    contrib = contrib && !mine;
    todo &= ~same;
    if (todo == 0u) return;
  }
  if (contrib) atomicAdd(&sh[bin], 1u);
}

// level: 0 -> bits 31..21 (2048 bins), 1 -> bits 20..10 (2048), 2 -> bits 9..0 (1024)
__global__ void __launch_bounds__(256) hist_kernel(Chunks c, int level, int take_abs, int finite_only,
                                                   const SelState* st, unsigned int* hist, unsigned long long* count) {
  __shared__ unsigned int sh[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
  __syncthreads();
  const unsigned int prefix = level ? st->prefix : 0u, mask = level ? st->mask : 0u;
  const int shift = level == 0 ? 21 : (level == 1 ? 10 : 0);
  const unsigned int bins_mask = level == 2 ? 1023u : 2047u;
  unsigned long long local = 0;
  scan_chunks(c, [&](float v, bool in) {
    unsigned int key = 0;
    const bool ok = in && sample_key(v, take_abs, finite_only, &key);
    local += ok;
    hist_add(sh, (key >> shift) & bins_mask, ok && (key & mask) == prefix);
  });
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += 256)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
  if (level == 0) {
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
  }
}

__global__ void pick_kernel(int level, SelState* st, unsigned int* hist, const unsigned long long* count) {
  if (threadIdx.x != 0) return;
  if (level == 0) {
    st->n = *count;
    if (st->n == 0) { st->rank = 0; st->prefix = 0; st->mask = 0; return; }
    if (st->rank >= st->n) st->rank = st->n - 1;
    st->out[2] = (double)st->rank;
    st->prefix = 0; st->mask = 0;
  }
  if (st->n == 0) return;
  const int shift = level == 0 ? 21 : (level == 1 ? 10 : 0);
  const int bins = level == 2 ? 1024 : 2048;
  unsigned long long r = st->rank;
  int b = 0;
  for (; b < bins; ++b) {
    unsigned int h = hist[b];
    if (r < h) break;
    r -= h;
  }
  if (b >= bins) b = bins - 1;
  st->rank = r;
  st->prefix |= ((unsigned int)b) << shift;
  st->mask |= ((unsigned int)(bins - 1)) << shift;
  for (int i = 0; i < 2048; ++i) hist[i] = 0;
  if (level == 2) st->key_k = st->prefix;
}

__global__ void __launch_bounds__(256) next_kernel(Chunks c, int take_abs, int finite_only, SelState* st) {
  if (st->n == 0) return;
  const unsigned int kk = st->key_k;
  unsigned long long le = 0;
  unsigned int mn = 0xffffffffu;
  scan_chunks(c, [&](float v, bool in) {
    unsigned int key = 0;
    if (in && sample_key(v, take_abs, finite_only, &key)) {
      if (key <= kk) ++le;
      else if (key < mn) mn = key;
    }
  });
  for (int o = 16; o; o >>= 1) {
    le += __shfl_down_sync(0xffffffffu, le, o);
    unsigned int other = __shfl_down_sync(0xffffffffu, mn, o);
    mn = other < mn ? other : mn;
  }
  if ((threadIdx.x & 31) == 0) {
    if (le) atomicAdd(&st->cnt_le, le);
    atomicMin(&st->key_next, mn);
  }
}

__global__ void finish_kernel(SelState* st, int take_abs, double* result, unsigned long long abs_rank_plus1_hint) {
  if (threadIdx.x != 0) return;
  (void)abs_rank_plus1_hint;
  if (st->n == 0) {
    result[0] = nan(""); result[1] = nan(""); result[2] = 0.0; result[3] = 0.0;
    return;
  }
  unsigned long long k = (unsigned long long)st->out[2];
  float ak = key_value(st->key_k, take_abs);
  float ak1 = ak;
  if (k + 1 < st->n && st->cnt_le < k + 2) ak1 = key_value(st->key_next, take_abs);
  result[0] = (double)ak;
  result[1] = (double)ak1;
  result[2] = st->out[2];
  result[3] = (double)st->n;
}

__global__ void count_only_kernel(const unsigned long long* count, double* result) {
  if (threadIdx.x == 0) { result[0] = nan(""); result[1] = nan(""); result[2] = 0.0; result[3] = (double)*count; }
}

__global__ void init_state_kernel(SelState* st, unsigned int* hist, unsigned long long* count, long long rank) {
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) {
    memset(st, 0, sizeof(SelState));
    st->rank = rank < 0 ? 0ull : (unsigned long long)rank;
    st->key_next = 0xffffffffu;
    *count = 0;
  }
}

// by-value variants for the staged (distributed) selection
__global__ void __launch_bounds__(256) hist_kernel_v(Chunks c, int level, int take_abs, int finite_only,
                                                     unsigned int prefix, unsigned int mask, unsigned int* hist,
                                                     unsigned long long* count) {
  __shared__ unsigned int sh[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
  __syncthreads();
  const int shift = level == 0 ? 21 : (level == 1 ? 10 : 0);
  const unsigned int bins_mask = level == 2 ? 1023u : 2047u;
  unsigned long long local = 0;
  scan_chunks(c, [&](float v, bool in) {
    unsigned int key = 0;
    const bool ok = in && sample_key(v, take_abs, finite_only, &key);
    local += ok;
    hist_add(sh, (key >> shift) & bins_mask, ok && (key & mask) == prefix);
  });
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += 256)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
  for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

__global__ void __launch_bounds__(256) rank_info_kernel(Chunks c, int take_abs, int finite_only, unsigned int kk,
                                                        unsigned long long* out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { /* out[] is pre-set by init below */ }
  unsigned long long le = 0;
  unsigned int mn = 0xffffffffu;
  scan_chunks(c, [&](float v, bool in) {
    unsigned int key = 0;
    if (in && sample_key(v, take_abs, finite_only, &key)) {
      if (key <= kk) ++le;
      else if (key < mn) mn = key;
    }
  });
  for (int o = 16; o; o >>= 1) {
    le += __shfl_down_sync(0xffffffffu, le, o);
    unsigned int other = __shfl_down_sync(0xffffffffu, mn, o);
    mn = other < mn ? other : mn;
  }
  if ((threadIdx.x & 31) == 0) {
    if (le) atomicAdd(&out[0], le);
    atomicMin(&out[1], (unsigned long long)mn);
  }
}

// per-chunk sample counts (the statistics pre-pass checks the valid fraction of every window, _norm_stats.py:268-270)
__global__ void __launch_bounds__(256) count_chunks_kernel(Chunks c, int finite_only, unsigned long long* counts) {
  const int64_t ntiles = c.rowstart[c.n];
  // per-thread running count of the current chunk; flushed (one atomic per warp) only when the CTA moves on to
  // another chunk -- one atomic per warp and tile made 2.4 M atomics on nine addresses the whole cost of the scan
  unsigned int local = 0;
  int cur = -1;
  auto flush = [&]() {
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&counts[cur], (unsigned long long)local);
    local = 0;
  };
  for (int64_t T = blockIdx.x; T < ntiles; T += gridDim.x) {
    int k = 0;
    while (k + 1 < c.n && T >= c.rowstart[k + 1]) ++k;
    if (k != cur) {          // (CTA-uniform)
      if (cur >= 0) flush();
      cur = k;
    }
    const int64_t cols = c.cols[k];
    const int64_t tpr = (cols + SCAN_TILE - 1) / SCAN_TILE;
    const int64_t t = T - c.rowstart[k];
    const int64_t r = t / tpr, cb = (t - r * tpr) * SCAN_TILE;
    const float* row = c.ptr[k] + r * c.ld[k];
    const int64_t cend = cb + SCAN_TILE < cols ? cb + SCAN_TILE : cols;
    constexpr int NL = SCAN_TILE / 256;
    float v[NL];
#pragma unroll
    for (int q = 0; q < NL; ++q) {
      const int64_t x = cb + (int64_t)q * 256 + threadIdx.x;
      v[q] = x < cend ? row[x] : nanf("");
    }
#pragma unroll
    for (int q = 0; q < NL; ++q) local += finite_only ? (isfinite(v[q]) ? 1u : 0u) : (v[q] == v[q] ? 1u : 0u);
  }
  if (cur >= 0) flush();
}

// ---- device-staged selection (multi-GPU): the selection state never leaves the device ----------------
// Exchange area `x` (int64, all-reduced by the host layer between the stages, stream-ordered):
//   x[0..2047] histogram of the current radix level (SUM), x[2048] sample count (SUM, level 0),
//   x[2049] #keys <= key_k (SUM), x[2050] smallest key > key_k (MIN)
constexpr int SEL_X_COUNT = 2048, SEL_X_LE = 2049, SEL_X_NEXT = 2050, SEL_X_WORDS = 2056;

__global__ void sel_begin_kernel(SelState* st, unsigned long long* x) {
  for (int i = threadIdx.x; i < SEL_X_WORDS; i += blockDim.x) x[i] = i == SEL_X_NEXT ? 0xffffffffull : 0ull;
  if (threadIdx.x == 0) {
    memset(st, 0, sizeof(SelState));
    st->key_next = 0xffffffffu;
    st->cmp_above_min = 0xffffffffu;
  }
}

__global__ void __launch_bounds__(256) sel_hist_kernel(Chunks c, int level, int take_abs, int finite_only,
                                                       const SelState* st, unsigned long long* x) {
  __shared__ unsigned int sh[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
  __syncthreads();
  if (level > 0 && st->n == 0) return;
  const unsigned int prefix = level ? st->prefix : 0u, mask = level ? st->mask : 0u;
  const int shift = level == 0 ? 21 : (level == 1 ? 10 : 0);
  const unsigned int bins_mask = level == 2 ? 1023u : 2047u;
  unsigned long long local = 0;
  scan_chunks(c, [&](float v, bool in) {
    unsigned int key = 0;
    const bool ok = in && sample_key(v, take_abs, finite_only, &key);
    local += ok;
    hist_add(sh, (key >> shift) & bins_mask, ok && (key & mask) == prefix);
  });
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += 256)
    if (sh[i]) atomicAdd(&x[i], (unsigned long long)sh[i]);
  if (level == 0) {
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&x[SEL_X_COUNT], local);
  }
}

// One warp: picks the bucket of the (all-reduced) histogram that holds the wanted rank and clears the
// histogram for the next level.  Level 0 also derives the rank from the percentile the way NumPy does for an
// f32 sample: virtual index (n-1)*f32(q/100) evaluated in f32, rank = floor (np.percentile, method 'linear').
__global__ void sel_pick_kernel(int level, float q32, SelState* st, unsigned long long* x) {
  const int lane = threadIdx.x & 31;
  if (level == 0) {
    if (lane == 0) {
      st->n = x[SEL_X_COUNT];
      st->prefix = 0; st->mask = 0;
      if (st->n) {
        const float vi = __fmul_rn(__ull2float_rn(st->n - 1ull), q32);
        long long k = (long long)floorf(vi);
        if (k < 0) k = 0;
        if ((unsigned long long)k > st->n - 1ull) k = (long long)(st->n - 1ull);
        st->rank = (unsigned long long)k;
        st->out[2] = (double)k;
      }
    }
    __syncwarp();
  }
  if (st->n == 0) {
    for (int i = lane; i < 2048; i += 32) x[i] = 0ull;
    return;
  }
  const int shift = level == 0 ? 21 : (level == 1 ? 10 : 0);
  const int bins = level == 2 ? 1024 : 2048;
  const int per = bins / 32;
  unsigned long long mine = 0;
  for (int i = 0; i < per; ++i) mine += x[lane * per + i];
  unsigned long long incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned long long excl = incl - mine;
  const unsigned long long r = st->rank;
  const bool holds = r >= excl && r < incl;
  unsigned int ballot = __ballot_sync(0xffffffffu, holds);
  int owner = ballot ? __ffs(ballot) - 1 : 31;   // rank beyond the total (cannot happen): last bucket
  __syncwarp();
  if (lane == owner) {
    unsigned long long rr = r - excl;
    int b = lane * per;
    const int bend = b + per;
    for (; b < bend; ++b) {
      const unsigned long long h = x[b];
      if (rr < h) break;
      rr -= h;
    }
    if (b >= bend) { b = bend - 1; rr = 0; }
    st->rank = rr;
    st->prefix |= ((unsigned int)b) << shift;
    st->mask |= ((unsigned int)(bins - 1)) << shift;
    if (level == 2) st->key_k = st->prefix;
  }
  __syncwarp();
  for (int i = lane; i < 2048; i += 32) x[i] = 0ull;
}

__global__ void __launch_bounds__(256) sel_next_kernel(Chunks c, int take_abs, int finite_only, const SelState* st,
                                                       unsigned long long* x) {
  if (st->n == 0) return;
  const unsigned int kk = st->key_k;
  unsigned long long le = 0;
  unsigned int mn = 0xffffffffu;
  scan_chunks(c, [&](float v, bool in) {
    unsigned int key = 0;
    if (in && sample_key(v, take_abs, finite_only, &key)) {
      if (key <= kk) ++le;
      else if (key < mn) mn = key;
    }
  });
  for (int o = 16; o; o >>= 1) {
    le += __shfl_down_sync(0xffffffffu, le, o);
    unsigned int other = __shfl_down_sync(0xffffffffu, mn, o);
    mn = other < mn ? other : mn;
  }
  if ((threadIdx.x & 31) == 0) {
    if (le) atomicAdd(&x[SEL_X_LE], le);
    atomicMin(&x[SEL_X_NEXT], (unsigned long long)mn);
  }
}

__global__ void sel_finish_kernel(const SelState* st, const unsigned long long* x, int take_abs, double* result) {
  if (threadIdx.x != 0) return;
  if (st->n == 0) {
    result[0] = nan(""); result[1] = nan(""); result[2] = 0.0; result[3] = 0.0;
    return;
  }
  const unsigned long long k = (unsigned long long)st->out[2];
  const float ak = key_value(st->key_k, take_abs);
  float ak1 = ak;
  if (k + 1 < st->n && x[SEL_X_LE] < k + 2) ak1 = key_value((unsigned int)x[SEL_X_NEXT], take_abs);
  result[0] = (double)ak;
  result[1] = (double)ak1;
  result[2] = st->out[2];
  result[3] = (double)st->n;
}

// ---- levels 1, 2 and the a[k+1] pass on the compacted level-0 bucket -------------------------------------------
// After the level-0 pick only the keys of one of 2048 buckets can still matter (a few per cent of the sample for
// |topousm|): one more scan copies them out (warp-aggregated append) and counts what lies below / finds the smallest
// key above, and the remaining three stages read the compact keys instead of the sample: two scans instead of four.
__global__ void __launch_bounds__(256) sel_compact_kernel(Chunks c, int take_abs, int finite_only, SelState* st,
                                                          unsigned int* __restrict__ keys, unsigned long long cap) {
  if (st->n == 0) return;
  const unsigned int prefix = st->prefix, mask = st->mask;
  const int lane = threadIdx.x & 31;
  unsigned long long below = 0;
  unsigned int above = 0xffffffffu;
  scan_chunks(c, [&](float v, bool in) {
    unsigned int key = 0;
    const bool ok = in && sample_key(v, take_abs, finite_only, &key);
    const bool hit = ok && (key & mask) == prefix;
    if (ok && !hit) {
      if (key < prefix) ++below;
      else if (key < above) above = key;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      unsigned long long pos = 0;
      if (lane == 0) pos = atomicAdd(&st->cmp_n, (unsigned long long)__popc(m));
      pos = __shfl_sync(0xffffffffu, pos, 0);
      if (hit) {
        const unsigned long long at = pos + __popc(m & ((1u << lane) - 1u));
        if (at < cap) keys[at] = key;
        else st->cmp_overflow = 1u;
      }
    }
  });
  for (int o = 16; o; o >>= 1) {
    below += __shfl_down_sync(0xffffffffu, below, o);
    const unsigned int other = __shfl_down_sync(0xffffffffu, above, o);
    above = other < above ? other : above;
  }
  if (lane == 0) {
    if (below) atomicAdd(&st->cmp_below, below);
    if (above != 0xffffffffu) atomicMin(&st->cmp_above_min, above);
  }
}

__global__ void __launch_bounds__(256) sel_hist_keys_kernel(const unsigned int* __restrict__ keys, int level,
                                                            const SelState* st, unsigned long long* x) {
  __shared__ unsigned int sh[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
  __syncthreads();
  if (st->n == 0) return;
  const unsigned int prefix = st->prefix, mask = st->mask;
  const int shift = level == 1 ? 10 : 0;
  const unsigned int bins_mask = level == 2 ? 1023u : 2047u;
  const unsigned long long n = st->cmp_n;
  const unsigned long long per = (unsigned long long)gridDim.x * 256ull;
  for (unsigned long long base = (unsigned long long)blockIdx.x * 256ull; base < n; base += per) {   // (whole warps)
    const unsigned long long i = base + threadIdx.x;
    const unsigned int key = i < n ? keys[i] : 0u;
    hist_add(sh, (key >> shift) & bins_mask, i < n && (key & mask) == prefix);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += 256)
    if (sh[i]) atomicAdd(&x[i], (unsigned long long)sh[i]);
}

__global__ void __launch_bounds__(256) sel_next_keys_kernel(const unsigned int* __restrict__ keys, const SelState* st,
                                                            unsigned long long* x) {
  if (st->n == 0) return;
  const unsigned int kk = st->key_k;
  const unsigned long long n = st->cmp_n;
  unsigned long long le = 0;
  unsigned int mn = 0xffffffffu;
  for (unsigned long long i = (unsigned long long)blockIdx.x * 256ull + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256ull) {
    const unsigned int key = keys[i];
    if (key <= kk) ++le;
    else if (key < mn) mn = key;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // what lies outside the bucket (counted once per rank)
    le += st->cmp_below;
    mn = st->cmp_above_min < mn ? st->cmp_above_min : mn;
  }
  for (int o = 16; o; o >>= 1) {
    le += __shfl_down_sync(0xffffffffu, le, o);
    const unsigned int other = __shfl_down_sync(0xffffffffu, mn, o);
    mn = other < mn ? other : mn;
  }
  if ((threadIdx.x & 31) == 0) {
    if (le) atomicAdd(&x[SEL_X_LE], le);
    atomicMin(&x[SEL_X_NEXT], (unsigned long long)mn);
  }
}

// ---- exchange over peer memory (NVLink / NVSwitch) instead of all-reduce calls --------------------------------
// Every rank owns SEL_PEER_SLOTS slots of SEL_X_WORDS words in symmetric memory (mapped by all ranks of the node).
// A stage publishes its part of the exchange area into the slot of the stage; after a stream-ordered barrier every
// rank sums (word SEL_X_NEXT: minimum) the slots of all ranks into its own exchange area -- the same integers on
// every rank, so all ranks take the same decisions.  Loads of peer memory bypass L1 (the slots are rewritten every
// call at the same addresses).
constexpr int SEL_PEER_SLOTS = 4;

__global__ void sel_publish_kernel(const unsigned long long* __restrict__ x, unsigned long long* slot, int first, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) slot[first + i] = x[first + i];
}

__global__ void sel_peer_reduce_kernel(unsigned long long* x, const unsigned long long* const* peers, int world,
                                       int slot_words0, int first, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int w = first + i;
    unsigned long long acc = w == SEL_X_NEXT ? 0xffffffffffffffffull : 0ull;
    for (int p = 0; p < world; ++p) {
      const unsigned long long v = __ldcv(peers[p] + slot_words0 + w);
      acc = w == SEL_X_NEXT ? (v < acc ? v : acc) : acc + v;
    }
    x[w] = acc;
  }
}

// np.percentile(..., method 'linear') of an f32 sample, finished on the device: virtual index, gamma and the lerp
// in f32 exactly as NumPy evaluates them (python-int / python-float operands are weak: (n - 1) * q32, diff * gamma
// and 1 - gamma are f32 operations), then the rule of topousm_fast_stat_func (_normalization.py:22-32): a NaN or
// a scale <= 1e-9 means "no global scale" -> NaN (the fused kernels read NaN as "do not normalise").
__global__ void sel_finish_scale_kernel(const SelState* st, const unsigned long long* x, int take_abs, float q32,
                                        float min_valid, float* scale) {
  if (threadIdx.x != 0) return;
  float out = nanf("");
  if (st->n != 0) {
    const unsigned long long k = (unsigned long long)st->out[2];
    const float ak = key_value(st->key_k, take_abs);
    float ak1 = ak;
    if (k + 1 < st->n && x[SEL_X_LE] < k + 2) ak1 = key_value((unsigned int)x[SEL_X_NEXT], take_abs);
    const float vi = __fmul_rn(__ull2float_rn(st->n - 1ull), q32);
    const float gamma = __fsub_rn(vi, floorf(vi));
    const float diff = __fsub_rn(ak1, ak);
    out = __fadd_rn(ak, __fmul_rn(diff, gamma));
    if (gamma >= 0.5f) out = __fsub_rn(ak1, __fmul_rn(diff, __fsub_rn(1.f, gamma)));
    if (!(out == out) || out <= min_valid) out = nanf("");
  }
  scale[0] = out;
}

// Bounding box of the finite samples of a strided overview of a row band (algorithms/_norm_stats.py:254-264 looks
// at a <= 512 px overview): box = (max of -row, max row, max of -col, max col) in overview indices, so that one
// MAX all-reduce merges the ranks; the caller presets the four words to INT_MIN.
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ band, int64_t ld, int64_t first_row,
                                                   int64_t cov, int64_t n_rows, int64_t n_cols, int64_t row_index0,
                                                   int* box) {
  const int64_t r = blockIdx.x;
  if (r >= n_rows) return;
  const float* row = band + (first_row + r * cov) * ld;
  int cmin = 0x7fffffff, cmax = -1;
  for (int64_t c = threadIdx.x; c < n_cols; c += 256) {
    const float v = row[c * cov];
    if (isfinite(v)) {
      cmin = cmin < (int)c ? cmin : (int)c;
      cmax = cmax > (int)c ? cmax : (int)c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int a = __shfl_xor_sync(0xffffffffu, cmin, o), b = __shfl_xor_sync(0xffffffffu, cmax, o);
    cmin = a < cmin ? a : cmin;
    cmax = b > cmax ? b : cmax;
  }
  if ((threadIdx.x & 31) == 0 && cmax >= 0) {
    const int ri = (int)(row_index0 + r);
    atomicMax(&box[0], -ri);
    atomicMax(&box[1], ri);
    atomicMax(&box[2], -cmin);
    atomicMax(&box[3], cmax);
  }
}

static int fill_chunks(Chunks& c, const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                       const int64_t* ld_host, int n_chunks, int64_t* total) {
  c.n = n_chunks;
  int64_t tot = 0;
  for (int i = 0; i < n_chunks; ++i) {
    if (!chunks_host[i] || rows_host[i] < 0 || cols_host[i] < 0 || ld_host[i] < cols_host[i]) return -1;
    c.ptr[i] = chunks_host[i]; c.rows[i] = rows_host[i]; c.cols[i] = cols_host[i]; c.ld[i] = ld_host[i];
    c.start[i] = tot;
    c.rowstart[i] = i == 0 ? 0 : c.rowstart[i - 1] + rows_host[i - 1] * ((cols_host[i - 1] + SCAN_TILE - 1) / SCAN_TILE);
    tot += rows_host[i] * cols_host[i];
  }
  c.start[n_chunks] = tot;
  c.rowstart[n_chunks] = n_chunks ? c.rowstart[n_chunks - 1] + rows_host[n_chunks - 1] * ((cols_host[n_chunks - 1] + SCAN_TILE - 1) / SCAN_TILE) : 0;
  *total = tot;
  return 0;
}

}  // namespace fsg

extern "C" {

size_t fsg_order_stats_workspace_bytes(void) { return 2048 * 4 + 256 + sizeof(fsg::SelState) + 256; }

int fsg_order_stats(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                    const int64_t* ld_host, int n_chunks, int64_t rank, int take_abs, int finite_only,
                    double* result_dev, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace fsg;
  if (n_chunks < 1 || n_chunks > MAX_CHUNKS) return fail(FSG_E_INVALID, "fsg_order_stats: 1..%d chunks supported", MAX_CHUNKS);
  if (!workspace || workspace_bytes < fsg_order_stats_workspace_bytes() || !result_dev)
    return fail(FSG_E_WORKSPACE, "fsg_order_stats: workspace too small");
  Chunks c{};
  c.n = n_chunks;
  int64_t tot = 0;
  for (int i = 0; i < n_chunks; ++i) {
    if (!chunks_host[i] || rows_host[i] < 0 || cols_host[i] < 0 || ld_host[i] < cols_host[i])
      return fail(FSG_E_INVALID, "fsg_order_stats: bad chunk %d", i);
    c.ptr[i] = chunks_host[i]; c.rows[i] = rows_host[i]; c.cols[i] = cols_host[i]; c.ld[i] = ld_host[i];
    c.start[i] = tot;
    c.rowstart[i] = i == 0 ? 0 : c.rowstart[i - 1] + rows_host[i - 1] * ((cols_host[i - 1] + 2047) / 2048);
    tot += rows_host[i] * cols_host[i];
  }
  c.start[n_chunks] = tot;
  c.rowstart[n_chunks] = n_chunks ? c.rowstart[n_chunks - 1] + rows_host[n_chunks - 1] * ((cols_host[n_chunks - 1] + 2047) / 2048) : 0;
  unsigned char* base = (unsigned char*)workspace;
  unsigned int* hist = (unsigned int*)base;
  unsigned long long* count = (unsigned long long*)(base + 2048 * 4);
  SelState* st = (SelState*)(base + 2048 * 4 + 256);
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = (int)((tot + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  init_state_kernel<<<1, 256, 0, s>>>(st, hist, count, (long long)rank);
  FSG_LAUNCH_OK();
  if (rank < 0) {  // count only
    hist_kernel<<<blocks, 256, 0, s>>>(c, 0, take_abs, finite_only, st, hist, count);
    FSG_LAUNCH_OK();
    count_only_kernel<<<1, 32, 0, s>>>(count, result_dev);
    FSG_LAUNCH_OK();
    return FSG_OK;
  }
  for (int level = 0; level < 3; ++level) {
    hist_kernel<<<blocks, 256, 0, s>>>(c, level, take_abs, finite_only, st, hist, count);
    FSG_LAUNCH_OK();
    pick_kernel<<<1, 32, 0, s>>>(level, st, hist, count);
    FSG_LAUNCH_OK();
  }
  next_kernel<<<blocks, 256, 0, s>>>(c, take_abs, finite_only, st);
  FSG_LAUNCH_OK();
  finish_kernel<<<1, 32, 0, s>>>(st, take_abs, result_dev, 0ull);
  FSG_LAUNCH_OK();
  return FSG_OK;
}


/* ---- staged selection for distributed (multi-GPU) percentiles ---------------------------------
 * Every rank histograms the keys of its own chunks that match (key & mask) == prefix; the host
 * all-reduces the 2048 bins, picks the bucket that holds the wanted rank and recurses (3 levels:
 * bits 31..21, 20..10, 9..0).  fsg_key_rank_info then returns #keys <= key and the smallest larger
 * key, which the host all-reduces (sum / min) to obtain a[k+1].  Keys: take_abs -> bits of |x|;
 * otherwise the order-preserving signed mapping.  fsg_key_to_float converts a key back. */
int fsg_key_histogram(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                      const int64_t* ld_host, int n_chunks, int level, uint32_t prefix, uint32_t mask, int take_abs,
                      int finite_only, uint32_t* hist_dev /* 2048 */, uint64_t* count_dev, void* stream) {
  using namespace fsg;
  if (n_chunks < 0 || n_chunks > MAX_CHUNKS || level < 0 || level > 2 || !hist_dev || !count_dev)
    return fail(FSG_E_INVALID, "fsg_key_histogram: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  FSG_CUDA_OK(cudaMemsetAsync(hist_dev, 0, 2048 * sizeof(uint32_t), s));
  FSG_CUDA_OK(cudaMemsetAsync(count_dev, 0, sizeof(uint64_t), s));
  if (n_chunks == 0) return FSG_OK;
  Chunks c{};
  c.n = n_chunks;
  int64_t tot = 0;
  for (int i = 0; i < n_chunks; ++i) {
    c.ptr[i] = chunks_host[i]; c.rows[i] = rows_host[i]; c.cols[i] = cols_host[i]; c.ld[i] = ld_host[i];
    c.start[i] = tot;
    c.rowstart[i] = i == 0 ? 0 : c.rowstart[i - 1] + rows_host[i - 1] * ((cols_host[i - 1] + 2047) / 2048);
    tot += rows_host[i] * cols_host[i];
  }
  c.start[n_chunks] = tot;
  c.rowstart[n_chunks] = n_chunks ? c.rowstart[n_chunks - 1] + rows_host[n_chunks - 1] * ((cols_host[n_chunks - 1] + 2047) / 2048) : 0;
  if (tot == 0) return FSG_OK;
  int blocks = (int)((tot + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  hist_kernel_v<<<blocks, 256, 0, s>>>(c, level, take_abs, finite_only, prefix, mask, hist_dev,
                                       (unsigned long long*)count_dev);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_key_rank_info(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                      const int64_t* ld_host, int n_chunks, uint32_t key, int take_abs, int finite_only,
                      uint64_t* out_dev /* [0] = #keys <= key, [1] = min key > key (0xffffffff if none) */, void* stream) {
  using namespace fsg;
  if (n_chunks < 0 || n_chunks > MAX_CHUNKS || !out_dev) return fail(FSG_E_INVALID, "fsg_key_rank_info: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  Chunks c{};
  c.n = n_chunks;
  int64_t tot = 0;
  for (int i = 0; i < n_chunks; ++i) {
    c.ptr[i] = chunks_host[i]; c.rows[i] = rows_host[i]; c.cols[i] = cols_host[i]; c.ld[i] = ld_host[i];
    c.start[i] = tot;
    c.rowstart[i] = i == 0 ? 0 : c.rowstart[i - 1] + rows_host[i - 1] * ((cols_host[i - 1] + 2047) / 2048);
    tot += rows_host[i] * cols_host[i];
  }
  c.start[n_chunks] = tot;
  c.rowstart[n_chunks] = n_chunks ? c.rowstart[n_chunks - 1] + rows_host[n_chunks - 1] * ((cols_host[n_chunks - 1] + 2047) / 2048) : 0;
  int blocks = (int)((tot + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const unsigned long long init[2] = {0ull, 0xffffffffull};
  FSG_CUDA_OK(cudaMemcpyAsync(out_dev, init, sizeof(init), cudaMemcpyHostToDevice, s));
  rank_info_kernel<<<blocks, 256, 0, s>>>(c, take_abs, finite_only, key, (unsigned long long*)out_dev);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_count_samples(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                      const int64_t* ld_host, int n_chunks, int finite_only, uint64_t* counts_dev, void* stream) {
  using namespace fsg;
  if (n_chunks < 1 || n_chunks > MAX_CHUNKS || !counts_dev) return fail(FSG_E_INVALID, "fsg_count_samples: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  Chunks c{};
  c.n = n_chunks;
  int64_t tot = 0;
  for (int i = 0; i < n_chunks; ++i) {
    if (!chunks_host[i] || rows_host[i] < 0 || cols_host[i] < 0 || ld_host[i] < cols_host[i])
      return fail(FSG_E_INVALID, "fsg_count_samples: bad chunk %d", i);
    c.ptr[i] = chunks_host[i]; c.rows[i] = rows_host[i]; c.cols[i] = cols_host[i]; c.ld[i] = ld_host[i];
    c.start[i] = tot;
    c.rowstart[i] = i == 0 ? 0 : c.rowstart[i - 1] + rows_host[i - 1] * ((cols_host[i - 1] + 2047) / 2048);
    tot += rows_host[i] * cols_host[i];
  }
  c.start[n_chunks] = tot;
  c.rowstart[n_chunks] = c.rowstart[n_chunks - 1] + rows_host[n_chunks - 1] * ((cols_host[n_chunks - 1] + 2047) / 2048);
  FSG_CUDA_OK(cudaMemsetAsync(counts_dev, 0, (size_t)n_chunks * sizeof(uint64_t), s));
  if (tot == 0) return FSG_OK;
  int64_t blocks = c.rowstart[n_chunks];
  if (blocks > 148 * 16) blocks = 148 * 16;
  count_chunks_kernel<<<(int)blocks, 256, 0, s>>>(c, finite_only, (unsigned long long*)counts_dev);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

/* ---- device-staged selection: same radix select, but the state stays on the device and the host layer only
 * enqueues kernels and all-reduces of the exchange area (no host round trip between the levels).
 * workspace: fsg_select_workspace_bytes() bytes, 8-byte aligned; the exchange area is its first
 * fsg_select_exchange_words() int64 words (see SEL_X_* above). */
size_t fsg_select_exchange_words(void) { return fsg::SEL_X_WORDS; }
size_t fsg_select_workspace_bytes(void) { return fsg::SEL_X_WORDS * 8 + sizeof(fsg::SelState) + 64; }

static fsg::SelState* sel_state(void* ws) { return (fsg::SelState*)((unsigned char*)ws + fsg::SEL_X_WORDS * 8); }

int fsg_select_begin(void* workspace, size_t workspace_bytes, void* stream) {
  using namespace fsg;
  if (!workspace || workspace_bytes < fsg_select_workspace_bytes() || ((uintptr_t)workspace & 7))
    return fail(FSG_E_WORKSPACE, "fsg_select_begin: workspace too small or misaligned");
  sel_begin_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sel_state(workspace), (unsigned long long*)workspace);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_hist(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                    const int64_t* ld_host, int n_chunks, int level, int take_abs, int finite_only, void* workspace,
                    void* stream) {
  using namespace fsg;
  if (n_chunks < 0 || n_chunks > MAX_CHUNKS || level < 0 || level > 2 || !workspace)
    return fail(FSG_E_INVALID, "fsg_select_hist: bad argument");
  if (n_chunks == 0) return FSG_OK;
  Chunks c{};
  int64_t tot = 0;
  if (fill_chunks(c, chunks_host, rows_host, cols_host, ld_host, n_chunks, &tot)) return fail(FSG_E_INVALID, "fsg_select_hist: bad chunk");
  if (tot == 0) return FSG_OK;
  int blocks = (int)((tot + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sel_hist_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(c, level, take_abs, finite_only, sel_state(workspace),
                                                            (unsigned long long*)workspace);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_pick(int level, float q32, void* workspace, void* stream) {
  using namespace fsg;
  if (level < 0 || level > 2 || !workspace) return fail(FSG_E_INVALID, "fsg_select_pick: bad argument");
  sel_pick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(level, q32, sel_state(workspace), (unsigned long long*)workspace);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_next(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                    const int64_t* ld_host, int n_chunks, int take_abs, int finite_only, void* workspace, void* stream) {
  using namespace fsg;
  if (n_chunks < 0 || n_chunks > MAX_CHUNKS || !workspace) return fail(FSG_E_INVALID, "fsg_select_next: bad argument");
  if (n_chunks == 0) return FSG_OK;
  Chunks c{};
  int64_t tot = 0;
  if (fill_chunks(c, chunks_host, rows_host, cols_host, ld_host, n_chunks, &tot)) return fail(FSG_E_INVALID, "fsg_select_next: bad chunk");
  if (tot == 0) return FSG_OK;
  int blocks = (int)((tot + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sel_next_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(c, take_abs, finite_only, sel_state(workspace),
                                                            (unsigned long long*)workspace);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_compact(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                       const int64_t* ld_host, int n_chunks, int take_abs, int finite_only, void* workspace,
                       uint32_t* keys_dev, int64_t capacity, void* stream) {
  using namespace fsg;
  if (n_chunks < 0 || n_chunks > MAX_CHUNKS || !workspace || capacity < 0) return fail(FSG_E_INVALID, "fsg_select_compact: bad argument");
  if (n_chunks == 0) return FSG_OK;
  Chunks c{};
  int64_t tot = 0;
  if (fill_chunks(c, chunks_host, rows_host, cols_host, ld_host, n_chunks, &tot)) return fail(FSG_E_INVALID, "fsg_select_compact: bad chunk");
  if (tot == 0) return FSG_OK;
  if (!keys_dev || capacity < tot) return fail(FSG_E_WORKSPACE, "fsg_select_compact: the key buffer must hold one key per sample");
  int blocks = (int)((tot + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sel_compact_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(c, take_abs, finite_only, sel_state(workspace), keys_dev,
                                                               (unsigned long long)capacity);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_hist_keys(const uint32_t* keys_dev, int level, void* workspace, void* stream) {
  using namespace fsg;
  if (!workspace || level < 1 || level > 2) return fail(FSG_E_INVALID, "fsg_select_hist_keys: bad argument");
  if (!keys_dev) return FSG_OK;
  sel_hist_keys_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(keys_dev, level, sel_state(workspace), (unsigned long long*)workspace);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_next_keys(const uint32_t* keys_dev, void* workspace, void* stream) {
  using namespace fsg;
  if (!workspace) return fail(FSG_E_INVALID, "fsg_select_next_keys: bad argument");
  if (!keys_dev) return FSG_OK;
  sel_next_keys_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(keys_dev, sel_state(workspace), (unsigned long long*)workspace);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

size_t fsg_select_peer_slot_words(void) { return (size_t)fsg::SEL_PEER_SLOTS * fsg::SEL_X_WORDS; }

int fsg_select_peer_publish(const void* workspace, void* my_slots, int stage, void* stream) {
  using namespace fsg;
  if (!workspace || !my_slots || stage < 0 || stage >= SEL_PEER_SLOTS) return fail(FSG_E_INVALID, "fsg_select_peer_publish: bad argument");
  const int first = stage < 3 ? 0 : SEL_X_LE, n = stage == 0 ? SEL_X_COUNT + 1 : (stage < 3 ? SEL_X_COUNT : 2);
  sel_publish_kernel<<<stage < 3 ? 4 : 1, 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)workspace,
      (unsigned long long*)my_slots + (size_t)stage * SEL_X_WORDS, first, n);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_peer_reduce(void* workspace, const void* const* peer_slots_dev, int world, int stage, void* stream) {
  using namespace fsg;
  if (!workspace || !peer_slots_dev || world < 1 || stage < 0 || stage >= SEL_PEER_SLOTS)
    return fail(FSG_E_INVALID, "fsg_select_peer_reduce: bad argument");
  const int first = stage < 3 ? 0 : SEL_X_LE, n = stage == 0 ? SEL_X_COUNT + 1 : (stage < 3 ? SEL_X_COUNT : 2);
  sel_peer_reduce_kernel<<<stage < 3 ? 9 : 1, 256, 0, (cudaStream_t)stream>>>((unsigned long long*)workspace,
      (const unsigned long long* const*)peer_slots_dev, world, stage * SEL_X_WORDS, first, n);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_finish(void* workspace, int take_abs, double* result_dev, void* stream) {
  using namespace fsg;
  if (!workspace || !result_dev) return fail(FSG_E_INVALID, "fsg_select_finish: bad argument");
  sel_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sel_state(workspace), (const unsigned long long*)workspace, take_abs, result_dev);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_select_finish_scale(void* workspace, int take_abs, float q32, float min_valid, float* scale_dev, void* stream) {
  using namespace fsg;
  if (!workspace || !scale_dev) return fail(FSG_E_INVALID, "fsg_select_finish_scale: bad argument");
  sel_finish_scale_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sel_state(workspace), (const unsigned long long*)workspace,
                                                              take_abs, q32, min_valid, scale_dev);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

int fsg_valid_bbox(const float* band, int64_t ld, int64_t first_row, int64_t cov, int64_t n_rows, int64_t n_cols,
                   int64_t row_index0, int32_t* box_dev, void* stream) {
  using namespace fsg;
  if (!band || !box_dev || cov < 1 || n_cols < 0 || n_rows < 0) return fail(FSG_E_INVALID, "fsg_valid_bbox: bad argument");
  FSG_CUDA_OK(cudaMemsetAsync(box_dev, 0x80, 4 * sizeof(int32_t), (cudaStream_t)stream));   // 0x80808080 < any index
  if (n_rows == 0 || n_cols == 0) return FSG_OK;
  if (n_rows > 0x7fffffff) return fail(FSG_E_UNSUPPORTED, "fsg_valid_bbox: overview too tall");
  bbox_kernel<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(band, ld, first_row, cov, n_rows, n_cols, row_index0, box_dev);
  FSG_LAUNCH_OK();
  return FSG_OK;
}

float fsg_key_to_float(uint32_t key, int take_abs) {
  uint32_t b = take_abs ? key : ((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
  float f;
  memcpy(&f, &b, 4);
  return f;
}

}  // extern "C"
