/* fsg_b200.h -- C ABI of libfsg_b200.so: B200 (sm_100a) kernels for the per-block
 * terrain-shading hot path of geoign/FujiShaderGPU v1.0.1.
 *
 * Every entry point replaces one CuPy "block function" of the reference (paths below are
 * relative to FujiShaderGPU/ in the reference tree).  Conventions:
 *   - plain pointers and sizes only; rasters are row-major float32, NaN = NoData;
 *     `ld_*` are row strides in ELEMENTS of the respective buffer;
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - the caller owns every buffer (input, output, workspace); the library never calls
 *     cudaMalloc/cudaFree/cudaDeviceSynchronize, and enqueues all work on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return value 0 = ok, <0 = error (FSG_E_*); fsg_last_error() gives a thread-local text;
 *   - re-entrant: no mutable global state (the tile backend calls from up to 6 threads).
 *   - "None" for an optional double parameter is passed as NaN.
 *
 * Row windows: functions taking (H_global, row0, rows) operate on a horizontal band of a
 * larger raster: `dem` points at global row `buf_row0` and holds `buf_rows` rows; outputs are
 * produced for global rows [out_row0, out_row0+out_rows) into `out` (whose row 0 is global
 * row out_row0).  Raster-edge rules (one-sided differences, reflect / edge replicate) are
 * applied at the GLOBAL edges only, so row-band shards reproduce the single-block result.
 */
#ifndef FSG_B200_H
#define FSG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSG_OK 0
#define FSG_E_INVALID (-1)   /* bad argument (maps to ValueError on the Python side) */
#define FSG_E_CUDA (-2)      /* CUDA runtime error, see fsg_last_error() */
#define FSG_E_WORKSPACE (-3) /* workspace too small */
#define FSG_E_UNSUPPORTED (-4)

/* output encodings (io/output_encoding.py:130-190, core/dask_processor.py:997-1008) */
#define FSG_OUT_F32 0
#define FSG_OUT_I16 1
#define FSG_OUT_U8 2

/* slope units (algorithms/_impl_slope.py:28-33) */
#define FSG_SLOPE_DEGREE 0
#define FSG_SLOPE_PERCENT 1
#define FSG_SLOPE_RADIAN 2

/* curvature types (algorithms/_impl_curvature.py:34-52) */
#define FSG_CURV_MEAN 0
#define FSG_CURV_GAUSSIAN 1
#define FSG_CURV_PLANFORM 2
#define FSG_CURV_PROFILE 3

/* Integer encoding applied in the kernel epilogue: DN = clip(rint(f32(a)*v + f32(b)), dn_min,
 * dn_max), non-finite -> 0.  kind = FSG_OUT_F32 disables it. */
typedef struct fsg_encode {
  int32_t kind;
  int32_t dn_min;
  int32_t dn_max;
  int32_t _pad;
  double a_coef;
  double b_coef;
} fsg_encode;

/* Raster window descriptor shared by the stencil entry points. */
typedef struct fsg_window {
  int64_t H_global;  /* rows of the whole raster (edge rules)            */
  int64_t W;         /* columns (always the full width)                  */
  int64_t buf_row0;  /* global row of dem[0]                             */
  int64_t buf_rows;  /* rows available in dem                            */
  int64_t out_row0;  /* first global row to produce                      */
  int64_t out_rows;  /* number of rows to produce                        */
  int64_t ld_in;     /* dem row stride (elements)                        */
  int64_t ld_out;    /* out row stride (elements of the output dtype)    */
} fsg_window;

const char* fsg_last_error(void);
int fsg_version(void);
/* number of kernels this thread has launched through the library since fsg_reset_launch_count */
int64_t fsg_launch_count(void);
void fsg_reset_launch_count(void);
/* Optional per-thread timing of the dominant kernels (CUDA events on the launch stream); used by
 * bench.py for the roofline figure.  tags: 1 topousm fused, 2 pyramid, 4 gradient, 5 openness. */
void fsg_profile_enable(int on);
int fsg_profile_read(int* tags, float* ms, int max);

/* ---- gradient family -------------------------------------------------------------------
 * replaces compute_hillshade_block   (algorithms/_impl_hillshade.py:20-54)
 *          compute_slope_block       (algorithms/_impl_slope.py:19-35)
 *          compute_curvature_block   (algorithms/_impl_curvature.py:19-57)
 * incl. handle_nan_for_gradient     (algorithms/_nan_utils.py:50-74): NaN gap fill with the
 * NaN-aware sigma=1 Gaussian, np.gradient(edge_order=2), NaN restore.  Band halo: 2 rows
 * (hillshade/slope) or 3 rows (curvature) each side, +4 rows when the band may contain NaN. */
int fsg_hillshade(const float* dem, void* out, const fsg_window* win,
                  double azimuth, double altitude, double z_factor,
                  double pixel_size, double pixel_scale_x, double pixel_scale_y,
                  const fsg_encode* enc, void* stream);
int fsg_slope(const float* dem, void* out, const fsg_window* win, int unit,
              double pixel_size, double pixel_scale_x, double pixel_scale_y,
              const fsg_encode* enc, void* stream);
int fsg_curvature(const float* dem, void* out, const fsg_window* win, int curvature_type,
                  double pixel_size, double pixel_scale_x, double pixel_scale_y,
                  const fsg_encode* enc, void* stream);

/* ---- topousm_fast -----------------------------------------------------------------------
 * replaces compute_topousm_fast_efficient_block (algorithms/_impl_topousm_fast.py:49-100)
 * + apply_global_normalization/topousm_fast_norm_func (algorithms/_global_stats.py:123-153,
 * _normalization.py:35-41) + the integer encoding, in one pipeline:
 *   pyramid  : _downsample_nan_aware   (algorithms/_nan_utils.py:604-668)
 *   coarse   : handle_nan_with_uniform / _gaussian on the decimated grids (:18-47)
 *   fused    : full-res box means (reflect), align-corners bilinear taps of the coarse means
 *              (_upsample_to_shape, :671-698), ordered f32 weighted sum, /scale, NaN restore.
 * radii/weights are the already-resolved lists (host side keeps the reference's scale
 * construction).  norm_scale <= 0 or NaN: no normalisation (raw block output).
 * Whole-raster semantics: the block IS the raster [0,H) x [0,W). */
size_t fsg_topousm_fast_workspace_bytes(int64_t H, int64_t W, const int32_t* radii_host, int n_radii,
                                        double pixel_size);
int fsg_topousm_fast(const float* dem, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out,
                     const int32_t* radii_host, const float* weights_host, int n_radii,
                     double pixel_size, double norm_scale, const fsg_encode* enc,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Same pipeline, output restricted to a region of interest (the statistics pre-pass only needs the
 * centre of each stratified window, algorithms/_norm_stats.py:275-277).  `out` is still the full
 * H x W buffer; rows / column strips outside the region are left untouched (whole 264-column strips
 * and whole rows are written, so a little more than the region may be filled). */
int fsg_topousm_fast_roi(const float* dem, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out,
                         const int32_t* radii_host, const float* weights_host, int n_radii, double pixel_size,
                         double norm_scale, const fsg_encode* enc, void* workspace, size_t workspace_bytes,
                         int64_t roi_row0, int64_t roi_rows, int64_t roi_col0, int64_t roi_cols, void* stream);


/* ---- row-band shards of topousm_fast (multi-GPU; one band of rows per GPU) ----------------------
 * The whole-raster pipeline above, cut into stages so that the host can exchange halo rows between
 * neighbouring bands (fujishadergpu_b200/core/sharding.py does it with torch.distributed / NCCL):
 *   fsg_topousm_plan       : per radius -> kind (0 fused full-res box, 1 decimated level, 2 full-res
 *                            plane), decimation factor, box taps on its grid (0 = sigma-1 Gaussian),
 *                            and the DEM halo rows the fused pass needs (max fused radius).
 *   fsg_pyramid_band       : f x f valid means of the band's OWN rows (the band must start at a
 *                            global row that is a multiple of 16); flags_dev[k] = level k holds a void
 *                            cell (the enclosed-void fill needs the whole grid; see sharding.py).
 *   fsg_grid_mean_band     : NaN-aware box ('reflect') / sigma-1 Gaussian ('nearest') mean of a window
 *                            of grid rows; the edge rules act at the GLOBAL grid edges (gh rows).
 *   fsg_topousm_fused_band : the fused pass for global rows [out_row0, +out_rows) given DEM rows with a
 *                            fused-halo on each side and, per non-fused term, a window of its mean grid.
 * A band pipeline reproduces the whole-raster result bit for bit. */
int fsg_topousm_plan(const int32_t* radii_host, int n_radii, double pixel_size, int32_t* kind_host,
                     int32_t* factor_host, int32_t* size_host, int32_t* fused_halo_host);
int fsg_pyramid_band(const float* dem, int64_t rows, int64_t W, int64_t ld, const int32_t* factors_host,
                     int n_levels, float* const* grids_host, int32_t* flags_dev, void* stream);
size_t fsg_grid_mean_band_workspace_bytes(int64_t out_rows, int64_t gw);
int fsg_grid_mean_band(const float* src, int64_t src_row0, int64_t src_rows, int64_t gh, int64_t gw, int64_t ld,
                       int size, int64_t out_row0, int64_t out_rows, float* out,
                       void* workspace, size_t workspace_bytes, void* stream);
int fsg_topousm_fused_band(const float* dem, int64_t dem_row0, int64_t dem_rows, int64_t H, int64_t W, int64_t ld_in,
                           void* out, int64_t out_row0, int64_t out_rows, int64_t ld_out,
                           const int32_t* radii_host, const float* weights_host, int n_radii, double pixel_size,
                           const float* const* term_grids_host, const int64_t* term_grow0_host,
                           const int64_t* term_grows_host, double norm_scale, const fsg_encode* enc, void* stream);

/* fsg_topousm_fused_band with (1) a scratch area (fsg_topousm_fused_band_workspace_bytes) that lets the interior
 * fast path run (without it the general kernel does the whole band) and (2) optionally the normalisation scale
 * read from device memory at kernel time (norm_scale_dev != NULL: *norm_scale_dev replaces norm_scale; NaN: no
 * normalisation, <= 0: zeros), so that the launch can be enqueued before the p99 of the statistics pre-pass
 * (algorithms/_norm_stats.py:176-298) has reached the host. */
size_t fsg_topousm_fused_band_workspace_bytes(int64_t out_rows, int64_t W);
int fsg_topousm_fused_band_ws(const float* dem, int64_t dem_row0, int64_t dem_rows, int64_t H, int64_t W, int64_t ld_in,
                              void* out, int64_t out_row0, int64_t out_rows, int64_t ld_out,
                              const int32_t* radii_host, const float* weights_host, int n_radii, double pixel_size,
                              const float* const* term_grids_host, const int64_t* term_grow0_host,
                              const int64_t* term_grows_host, double norm_scale, const float* norm_scale_dev,
                              const fsg_encode* enc, void* workspace, size_t workspace_bytes, void* stream);

int fsg_grid_void_fill(float* grid, int64_t gh, int64_t gw, void* workspace, size_t workspace_bytes, void* stream);
/* Re-reads the FSG_* debug environment switches (they are otherwise read once per process). */
void fsg_debug_reload_switches(void);
/* host-only: rows per CTA the interior fast path chooses for `rows` x `strips` (its makespan model; for tests) */
int64_t fsg_debug_v8_band_rows(int64_t rows, int64_t strips);

/* ---- overview large-radius part (algorithms/_impl_topousm_fast.py:158-186,
 *      algorithms/_nan_utils.py:255-281): out = f32(w_large)*block - bilinear(field) */
int fsg_topousm_large_part(const float* block, float* out, int64_t h, int64_t w, int64_t ld_in, int64_t ld_out,
                           const float* field, int64_t ch, int64_t cw, int64_t ld_field,
                           int64_t off_r, int64_t off_c, int64_t full_h, int64_t full_w,
                           double w_large, void* stream);

/* ---- openness (algorithms/_impl_openness.py:31-132) -------------------------------------- */
int fsg_openness(const float* dem, void* out, const fsg_window* win, int negative,
                 int num_directions, int max_distance,
                 double pixel_size, double pixel_scale_x, double pixel_scale_y,
                 double stretch_lo, double stretch_scale, /* NaN/<=1e-12 scale: no stretch */
                 const fsg_encode* enc, void* stream);

/* Same kernel with an explicit ray-sample table: samples of azimuth d are
 * [dir_start[d], dir_start[d+1]); offsets in pixels, dist = f32(max(hypot(ox*|sx|, oy*|sy|), 1e-9)).
 * The Python host layer builds the table with NumPy exactly as the reference does (:58-68, :96-109). */
int fsg_openness_samples(const float* dem, void* out, const fsg_window* win, int negative, int num_directions,
                         const int32_t* dir_start_host, const int32_t* ox_host, const int32_t* oy_host,
                         const float* dist_host, double stretch_lo, double stretch_scale,
                         const fsg_encode* enc, void* stream);

/* ---- ambient occlusion (algorithms/_impl_ambient_occlusion.py:33-118) --------------
 * compute_ambient_occlusion_block on the whole H x W block: 4 rings x num_samples edge-replicated gathers,
 * clip(1 - mean occlusion * intensity), sigma=1 Gaussian ('nearest'), power 1/2.2, NaN restore; optional
 * display stretch (tile/dask_bridge.py:173-187) and integer encoding.  The sample table (offsets with the
 * (0,0) entries dropped, f32 physical distance, f32 ring weight 1 - 0.3 k/4) is built by the host layer with
 * NumPy exactly as the reference does (:52-88). */
size_t fsg_ambient_occlusion_workspace_bytes(int64_t H, int64_t W);
int fsg_ambient_occlusion(const float* dem, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out,
                          int n_samples, const int32_t* ox_host, const int32_t* oy_host, const float* dist_host,
                          const float* factor_host, double intensity, double stretch_lo, double stretch_scale,
                          const fsg_encode* enc, void* workspace, size_t workspace_bytes, void* stream);

/* ---- output pyramid (core/dask_processor.py:201-228, io/cog_builder.py:295-310) ---
 * One 2 x 2 AVERAGE overview level of an encoded raster (element kind FSG_OUT_*): NoData members (DN == nodata
 * when has_nodata, NaN for f32) are skipped, an empty cell is NoData, integers round half up.  out is
 * ceil(H/2) x ceil(W/2).  The host layer (io/cog_writer.py) cascades it for the COG's overview levels. */
int fsg_overview_average(const void* in, void* out, int64_t H, int64_t W, int64_t ld_in, int64_t ld_out, int kind,
                         double nodata, int has_nodata, void* stream);

/* ---- helpers ------------------------------------------------------------------------------
 * fsg_decimate      : _downsample_nan_aware incl. enclosed-void fill (needs workspace)
 * fsg_upsample      : _upsample_to_shape (plain and NaN-aware branch)
 * fsg_encode_f32    : _quantize_block_cp / quantize_array
 * fsg_scale_f32     : out = in / f32(scale) with NaN kept (topousm_fast_norm_func)
 * fsg_stretch_f32   : max((x - lo)/scale, 0)   (tile/dask_bridge.py:173-187)
 * fsg_order_stats   : exact order statistics for percentile(|x|, 99) (topousm_fast_stat_func,
 *                     _normalization.py:22-32) and p1/p99 (robust_unsigned_stretch_stat_func) */
size_t fsg_decimate_workspace_bytes(int64_t H, int64_t W, int factor);
int fsg_decimate(const float* in, float* out, int64_t H, int64_t W, int64_t ld_in, int factor,
                 void* workspace, size_t workspace_bytes, void* stream);
int fsg_upsample(const float* in, float* out, int64_t h, int64_t w, int64_t H, int64_t W,
                 void* workspace /* >= 256 bytes */, size_t workspace_bytes, void* stream);
int fsg_encode_f32(const float* in, void* out, int64_t n, const fsg_encode* enc, void* stream);
int fsg_scale_f32(const float* in, float* out, int64_t n, double scale, void* stream);
int fsg_stretch_f32(const float* in, float* out, int64_t n, double lo, double scale, void* stream);

/* ---- spatial mode of the gradient family --------------------------------
 * fsg_gaussian_nan : handle_nan_with_gaussian(block, sigma, mode='nearest')[0]
 *                    (algorithms/_nan_utils.py:18-31; the smoothing of _smooth_for_radius, :527-552)
 * fsg_combine_f32  : one step of _combine_direct (algorithms/tile/dask_bridge.py:28-69) /
 *                    _combine_multiscale_dask (_nan_utils.py:182-213); mode 0 acc=a*w, 1 acc+=a*w,
 *                    2 acc=0+a*w, 3 acc=maximum(acc,a), 4 acc=minimum(acc,a), 5 acc=a */
size_t fsg_gaussian_nan_workspace_bytes(int64_t H, int64_t W, double sigma);
int fsg_gaussian_nan(const float* in, float* out, int64_t H, int64_t W, int64_t ld_in, double sigma,
                     void* workspace, size_t workspace_bytes, void* stream);
int fsg_combine_f32(const float* a, float* acc, int64_t n, double w, int mode, void* stream);

size_t fsg_order_stats_workspace_bytes(void);
/* Pooled order statistics over up to 16 2-D chunks (|x| first when take_abs != 0; samples are the
 * non-NaN, or with finite_only the finite, values).  rank < 0: count only.  Otherwise exact radix
 * selection of the rank-th smallest sample.  result_dev (4 doubles): a[rank], a[min(rank+1,n-1)],
 * rank used, n.  NumPy computes the percentile's virtual index in f32 for f32 input, so the host
 * layer derives `rank` and interpolates with NumPy scalars (kernels.percentile). */
int fsg_order_stats(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                    const int64_t* ld_host, int n_chunks, int64_t rank, int take_abs, int finite_only,
                    double* result_dev, void* workspace, size_t workspace_bytes, void* stream);
/* Per-chunk counts of non-NaN (finite_only: finite) samples, counts_dev[n_chunks] (the statistics pre-pass
 * skips windows with less than 2 % valid pixels, algorithms/_norm_stats.py:268-270). */
int fsg_count_samples(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                      const int64_t* ld_host, int n_chunks, int finite_only, uint64_t* counts_dev, void* stream);


/* Staged selection for percentiles of a sample that is spread over several GPUs: per-rank key
 * histograms (3 radix levels) and rank information, combined by the host with all-reduces
 * (fujishadergpu_b200/core/sharding.py::distributed_percentile). */
int fsg_key_histogram(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                      const int64_t* ld_host, int n_chunks, int level, uint32_t prefix, uint32_t mask, int take_abs,
                      int finite_only, uint32_t* hist_dev, uint64_t* count_dev, void* stream);
int fsg_key_rank_info(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                      const int64_t* ld_host, int n_chunks, uint32_t key, int take_abs, int finite_only,
                      uint64_t* out_dev, void* stream);
float fsg_key_to_float(uint32_t key, int take_abs);

/* Device-staged variant of the same selection: the state stays on the device; between the stages the host
 * layer all-reduces the exchange area (the first fsg_select_exchange_words() int64 words of the workspace:
 * [0..2047] level histogram and [2048] sample count with SUM, [2049] #keys <= key with SUM, [2050] next key
 * with MIN) in stream order, so a percentile costs no host round trip until the 4-double result is read.
 * Stages: begin; for level 0..2 { hist; all-reduce x[0..2048]; pick }; next; all-reduce x[2049] (SUM),
 * x[2050] (MIN); finish.  q32 = f32(q)/f32(100): the rank is derived on the device with NumPy's f32
 * virtual-index arithmetic (np.percentile, method 'linear'; reference _normalization.py:22-32). */
size_t fsg_select_exchange_words(void);
size_t fsg_select_workspace_bytes(void);
int fsg_select_begin(void* workspace, size_t workspace_bytes, void* stream);
int fsg_select_hist(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                    const int64_t* ld_host, int n_chunks, int level, int take_abs, int finite_only, void* workspace,
                    void* stream);
int fsg_select_pick(int level, float q32, void* workspace, void* stream);
int fsg_select_next(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                    const int64_t* ld_host, int n_chunks, int take_abs, int finite_only, void* workspace, void* stream);
int fsg_select_finish(void* workspace, int take_abs, double* result_dev, void* stream);
/* Two scans of the sample instead of four: after fsg_select_pick(level 0) fsg_select_compact copies the keys of the
 * selected level-0 bucket into `keys_dev` (capacity >= the sample count, uint32) and counts what lies outside it;
 * fsg_select_hist_keys (levels 1, 2) and fsg_select_next_keys then replace fsg_select_hist / fsg_select_next.  The
 * exchange area and its reductions between the stages are unchanged. */
int fsg_select_compact(const float* const* chunks_host, const int64_t* rows_host, const int64_t* cols_host,
                       const int64_t* ld_host, int n_chunks, int take_abs, int finite_only, void* workspace,
                       uint32_t* keys_dev, int64_t capacity, void* stream);
int fsg_select_hist_keys(const uint32_t* keys_dev, int level, void* workspace, void* stream);
int fsg_select_next_keys(const uint32_t* keys_dev, void* workspace, void* stream);
/* The same exchange without collective calls, over peer memory (NVLink / NVSwitch): every rank owns
 * fsg_select_peer_slot_words() int64 words of symmetric memory mapped by all ranks of the node.  After a stage
 * (0..2: fsg_select_hist of that level, 3: fsg_select_next) a rank publishes its part of the exchange area into its
 * own slots; after a stream-ordered barrier of the ranks, fsg_select_peer_reduce sums (smallest key: minimum) the
 * slots of all `world` ranks (`peer_slots_dev`: device array of their base pointers, as mapped on this device) into
 * the workspace, which then is what the all-reduce would have left there. */
size_t fsg_select_peer_slot_words(void);
int fsg_select_peer_publish(const void* workspace, void* my_slots, int stage, void* stream);
int fsg_select_peer_reduce(void* workspace, const void* const* peer_slots_dev, int world, int stage, void* stream);
/* Same selection, finished on the device: scale_dev[0] = np.percentile(sample, 100 * q32) (method 'linear', NumPy's
 * f32 arithmetic), or NaN when the sample is empty or the value is NaN / <= min_valid (topousm_fast_stat_func,
 * algorithms/_normalization.py:22-32) -- the value fsg_topousm_fused_band_ws takes as norm_scale_dev, so that the
 * statistics pre-pass needs no host round trip. */
int fsg_select_finish_scale(void* workspace, int take_abs, float q32, float min_valid, float* scale_dev, void* stream);
/* Bounding box of the finite samples of the overview band[first_row + i * cov, j * cov], i < n_rows, j < n_cols
 * (algorithms/_norm_stats.py:254-264): box_dev = (max of -(row_index0 + i), max of row_index0 + i, max of -j, max of j),
 * INT_MIN-like words when no sample is finite; one MAX all-reduce merges the row bands of several ranks. */
int fsg_valid_bbox(const float* band, int64_t ld, int64_t first_row, int64_t cov, int64_t n_rows, int64_t n_cols,
                   int64_t row_index0, int32_t* box_dev, void* stream);


/* synthetic DEM generator used by bench/tests: eight sinusoid octaves +
 * hash noise, optional NoData wedge/ellipses; rows [row0,row0+rows) of an H x W raster. */
int fsg_synth_dem(float* out, int64_t H, int64_t W, int64_t row0, int64_t rows, int64_t ld,
                  uint64_t seed, int nodata, void* stream);

/* rows x cols f32 rectangle, device to device on the copy engines (cudaMemcpy2DAsync): `src` may be another GPU's
 * memory mapped into this process (symmetric memory; the window gather of the sharded statistics pre-pass,
 * reference: algorithms/_norm_stats.py:176-298 reads its windows from the one Dask array).  Row strides in elements. */
int fsg_copy_rect_f32(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int64_t cols,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSG_B200_H */
