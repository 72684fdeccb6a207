/*
 * machisplin_b200.h - C ABI of the B200-native TPS + ensemble raster-interpolation engine.
 *
 * This is the drop-in boundary for the hot path of jasonleebrown/machisplin
 * (reference @ e3a31fa).  The reference has no FFI of its own: its hot path is a set of R
 * generic calls into CRAN packages.  Each entry point below replaces one such call site
 * (V73:n = R/ensemble.machine.learning.thin.plate.splines.V73.R line n); INTEGRATION.md
 * shows the .Call shim an R maintainer adds.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success or a negative MB_E_*
 *     code and never throws or aborts; mb_last_error() gives the message (thread-local).
 *   - host entry points take caller-owned HOST arrays and perform H2D/D2H themselves;
 *     the *_dev twins take DEVICE pointers (of the context's device) and a cudaStream_t
 *     passed as void* (NULL = the context's stream) and do not synchronise.
 *   - rasters: terra cell order = row-major from the NW corner, x fastest.  NA = any NaN in,
 *     quiet NaN out.  Point / coefficient matrices are column-major ("R order").
 *   - all handles are opaque; the library owns device memory, the caller owns host memory.
 */
#ifndef MACHISPLIN_B200_H
#define MACHISPLIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_VERSION 100

enum {
  MB_OK = 0,
  MB_E_ARG = -1,        /* bad argument */
  MB_E_CUDA = -2,       /* CUDA runtime / driver failure */
  MB_E_NUMERIC = -3,    /* collinear null space, non-SPD system, eigen failure ... */
  MB_E_NOMEM = -4,
  MB_E_UNSUPPORTED = -5
};

/* Grid geometry of a terra SpatRaster: extent + dimensions. */
typedef struct {
  double xmin, xmax, ymin, ymax;
  int32_t nrow, ncol;
} mb_grid;

/* Half-open cell window rows [r0,r1) x cols [c0,c1). */
typedef struct {
  int32_t r0, r1, c0, c1;
} mb_window;

typedef struct mb_ctx mb_ctx;
typedef struct mb_spline mb_spline;
typedef struct mb_ensemble mb_ensemble;

/* ---- lifetime ---------------------------------------------------------------------- */
int mb_version(void);
int mb_device_count(void);
const char* mb_last_error(void);
int mb_init(int device, mb_ctx** ctx);
void mb_shutdown(mb_ctx* ctx);
int mb_sync(mb_ctx* ctx);
/* number of kernels this context has launched since creation (bench bookkeeping) */
int64_t mb_launch_count(const mb_ctx* ctx);

/* ---- a1: fields::Tps  (V73:722, V73:751) --------------------------------------------- */
/* Fit a 2-D thin-plate smoothing spline (m=2, scale.type="range") to n observations.
 * xy: n x 2 column-major (LONG, LAT); y: n x L column-major (L responses sharing the knots,
 * L = 1 at the reference's call sites; L = 20 for BASELINE config 5).
 * lambda < 0 selects fields' GCV search per response; lambda >= 0 is used as given.
 * Duplicate locations are pooled exactly like Krig.replicates.  splines[L] receives handles. */
int mb_tps_fit(mb_ctx* ctx, const double* xy, const double* y, int n, int L, double lambda,
               mb_spline** splines);
/* Build a spline handle from coefficients computed elsewhere (e.g. an R Krig object):
 * knots_xy np x 2 column-major UNSCALED, c[np], d[3], center[2], scale[2]. */
int mb_spline_create(mb_ctx* ctx, const double* knots_xy, int np, const double* c, const double* d,
                     const double* center, const double* scale, mb_spline** out);
int mb_spline_np(const mb_spline* s);
/* Any output pointer may be NULL.  knots_xy np x 2 column-major unscaled. */
int mb_spline_get(const mb_spline* s, double* c, double* d, double* center, double* scale,
                  double* knots_xy, double* lambda, double* eff_df, double* gcv_at_lambda);
/* GCV fits only (all outputs NULL-able): eigenvalues eta[np-3] of Q2'KQ2 (decreasing, = 1/D of Krig's "WBW"
 * decomposition), and the tridiagonal form the lambda search runs on: T = Q'(Q2'KQ2)Q as tri_diag[np-3],
 * tri_off[np-4], and the rotated data zhat[np-3] = Q'Q2' sqrt(w) yM. */
int mb_spline_get_decomp(const mb_spline* s, double* eta, double* tri_diag, double* tri_off, double* zhat);
void mb_spline_free(mb_spline* s);

/* ---- a2: terra::interpolate(rast(template), Tps)  (V73:726, V73:753) ----------------- */
enum {
  MB_EVAL_DIRECT = 0,  /* O(cells x knots) float64 pair sum, libm log: the parity kernel     */
  MB_EVAL_FAST = 1     /* hierarchical Chebyshev far field + float64 near field: the roofline kernel */
};
/* Predict every cell of window w of grid g; out is (w.r1-w.r0) x (w.c1-w.c0) row-major. */
int mb_tps_eval(mb_ctx* ctx, const mb_spline* s, const mb_grid* g, const mb_window* w, int method,
                double* out_host);
int mb_tps_eval_dev(mb_ctx* ctx, const mb_spline* s, const mb_grid* g, const mb_window* w, int method,
                    double* out_dev, int64_t out_row_stride, void* stream);
/* predict.Krig at arbitrary points (xy n x 2 column-major), float64 direct. */
int mb_tps_predict_points(mb_ctx* ctx, const mb_spline* s, const double* xy, int n, double* out_host);

/* ---- a5: terra::predict(rast_stack, model_k) x6 + weighted sum  (V73:468-606, 619) ---- */
/* Flat descriptors of the fitted models.  P = C + 2 features in the
 * order cov_1..cov_C, LONG, LAT (xnam, V73:194).  A NULL array / zero count = model absent. */
typedef struct {
  int32_t P;
  /* g: mgcv::gam, parametric formula            coef[P+1], intercept first */
  const double* gam_coef;
  /* n: nnet(size=H, linout=TRUE)                 wts[(P+1)*H + H + 1]; y = net*max2 + min (V73:469-470) */
  const double* nn_wts; int32_t nn_H; double nn_max2, nn_min;
  /* m: earth                                     dirs[T*P] row-major (0,+-1,2), cuts[T*P], coef[T] */
  int32_t mars_T; const int8_t* mars_dirs; const double* mars_cuts; const double* mars_coef;
  /* v: kernlab::ksvm eps-svr rbfdot scaled       sv[S*P] row-major (scaled), alpha[S] */
  int32_t svm_S; const double* svm_sv; const double* svm_alpha; double svm_b, svm_sigma;
  const double* svm_x_center; const double* svm_x_scale; double svm_y_center, svm_y_scale;
  /* r: randomForest regression                   arrays [ntree*nrnodes] tree-major; daughters 1-based, 0 = none */
  int32_t rf_ntree, rf_nrnodes;
  const int32_t* rf_left; const int32_t* rf_right; const int8_t* rf_status; const int32_t* rf_bestvar;
  const double* rf_split; const double* rf_nodepred;
  /* b: gbm gaussian                              tree t = nodes [tree_off[t], tree_off[t+1]) ; 0-based child offsets */
  int32_t gbm_ntrees; double gbm_initF; const int32_t* gbm_tree_off;
  const int32_t* gbm_splitvar; const double* gbm_splitcode;
  const int32_t* gbm_left; const int32_t* gbm_right; const int32_t* gbm_missing;
} mb_models;

/* kept: letters of the retained models in evaluation order, e.g. "bgnmrv" (V73:340-362);
 * w[strlen(kept)]: the round(p,2) weights; w_total: the UNROUNDED sum over all candidates (V73:337). */
int mb_ensemble_create(mb_ctx* ctx, const mb_grid* g, const mb_models* m, const char* kept,
                       const double* w, double w_total, mb_ensemble** out);
void mb_ensemble_free(mb_ensemble* e);

/* pred.elev (+ TPS surface, + NA propagation of part 5, V73:906-907) for window w.
 * cov: C planes of the FULL grid, plane stride nrow*ncol, float32, NaN = NA.
 * spline may be NULL (tps=FALSE); tps_surface may carry a precomputed TPS raster for the window
 * (row-major, e.g. from mb_tiles_tps) and is added instead of evaluating `spline`. */
int mb_ensemble_eval(mb_ctx* ctx, const mb_ensemble* e, const float* cov_host, int C,
                     const mb_spline* spline, const double* tps_surface_host,
                     const mb_window* w, double* out_host);
int mb_ensemble_eval_dev(mb_ctx* ctx, const mb_ensemble* e, const float* cov_dev, int C,
                         const mb_spline* spline, const double* tps_surface_dev,
                         const mb_window* w, double* out_dev, void* stream);
/* the same predictors at n points; X is n x P column-major float64 (V73:477-611 residual side) */
int mb_ensemble_predict_points(mb_ctx* ctx, const mb_ensemble* e, const double* X, int n, double* out_host);

/* ---- a3 + a4: mltps part 3 / 4  (V73:649-895) ---------------------------------------- */
/* The TPS-of-residuals raster of the full grid with the reference's internal tiling:
 * ceil(n/tile_px) tiles, fit box = tile +- fit_halo, keep box = tile +- keep_halo, < min_pts knots
 * -> zero tile, mean mosaic + linear seam feather.  knots_xy n x 2 column-major (cell-centre
 * LONG/LAT), resid[n].  Defaults of the reference: 1500, 0.2, 0.025, 10, lambda < 0 (GCV). */
int mb_tiles_tps(mb_ctx* ctx, const mb_grid* g, const double* knots_xy, const double* resid, int n,
                 int tile_px, double fit_halo, double keep_halo, int min_pts, double lambda,
                 int method, double* out_host);
int mb_tiles_tps_dev(mb_ctx* ctx, const mb_grid* g, const double* knots_xy, const double* resid, int n,
                     int tile_px, double fit_halo, double keep_halo, int min_pts, double lambda,
                     int method, double* out_dev, void* stream);

/* ---- machisplin.tiles.merge  (V73:1392-1548) ------------------------------------------ */
/* ntiles = nC*nR tile rasters in the reference's order (row-major from the SW tile); tile t covers
 * window wins[t] of grid g and is (r1-r0) x (c1-c0) row-major. */
int mb_tiles_merge(mb_ctx* ctx, const mb_grid* g, int nC, int nR, const mb_window* wins,
                   const double* const* tiles_host, double* out_host);
int mb_tiles_merge_dev(mb_ctx* ctx, const mb_grid* g, int nC, int nR, const mb_window* wins,
                       const double* const* tiles_dev, double* out_dev, void* stream);

/* machisplin.tiles.merge with the tiles spread over the GPUs of the context's communicator (the tile-border blend of the
 * multi-GPU path): tile t lives on rank t % mb_comm_size(ctx).  The raster is partitioned into one OWNED window per tile (its
 * share of the raster, cut through the middle of every overlap zone; mb_tiles_owned_window, host-only arithmetic); the ranks
 * exchange the seam strips they need point to point (ncclSend / ncclRecv: rectangles a feather.d / 2 wide), agree on the seam
 * boxes with one ncclAllReduce(min) of 4 integers per seam, and each blends the cells its tiles own.  No rank gathers a
 * neighbour's tile; the merged raster stays distributed: out_dev[t] (owned-window-shaped, row-major) for the tiles of this rank.
 * my_tiles_dev[t] / out_dev[t] are NULL for tiles of other ranks.  Without a communicator every tile is local (one process).
 * Values are bit-identical to mb_tiles_merge_dev on the same tiles. */
int mb_tiles_owned_window(const mb_grid* g, int nC, int nR, const mb_window* wins, int t, mb_window* own);
int mb_tiles_merge_shard_dev(mb_ctx* ctx, const mb_grid* g, int nC, int nR, const mb_window* wins,
                             const double* const* my_tiles_dev, double* const* out_dev, void* stream);

/* ---- a6: RSS objective of the ensemble-weight search  (V73:329-333, 369-373) ---------- */
/* G = R'R for R n x K column-major (K <= 8); fit(w) = w'Gw / (sum w)^2. */
int mb_gram(mb_ctx* ctx, const double* R_host, int n, int K, double* G_host);
int mb_gram_dev(mb_ctx* ctx, const double* R_dev, int n, int K, double* G_dev, void* stream);

/* ---- a7: part 5  (V73:906-930) --------------------------------------------------------- */
/* gather raster values at the cells (row[i], col[i]) - f.actual <- extract(final, points), V73:910.  The raster has nrow
 * rows of ncol cells, row_stride doubles apart; a cell outside it (e.g. row = col = -1 for a point outside the extent) gives
 * NaN, which is what terra::extract returns there. */
int mb_gather_cells_dev(mb_ctx* ctx, const double* raster_dev, int64_t row_stride, int nrow, int ncol, const int32_t* row,
                        const int32_t* col, int n, double* out_host, void* stream);

/* ---- mltps parts 2-5 for one response in ONE call  (V73:442-932) --------------------------- */
/* final = (sum_k round(w_k,2) f_k(cell)) / sum_all w  +  TPS(residuals)(cell), NA-propagating.
 * The per-cell tree / svm kernels (part 2) run on a second stream while fields::Tps is fitted (part 3);
 * the TPS surface, the smooth models and the combine (parts 3-5) are one fused pass over the grid.
 * e may be NULL (tps only); knots_xy / resid may be NULL (tps = FALSE).  tile_px <= 0: one global spline
 * (the nRx*nCx == 1 branch, V73:748-753); tile_px > 0: the reference's rule with that tile size (1500).
 * spline_out (NULL-able) receives the global spline (NULL in tiled mode).  g is the full raster grid. */
int mb_mltps_predict_dev(mb_ctx* ctx, const mb_grid* g, const mb_ensemble* e, const float* cov_dev, int C,
                         const double* knots_xy, const double* resid, int n, double lambda, int tile_px,
                         double* out_dev, mb_spline** spline_out, void* stream);
int mb_mltps_predict(mb_ctx* ctx, const mb_grid* g, const mb_ensemble* e, const float* cov_host, int C,
                     const double* knots_xy, const double* resid, int n, double lambda, int tile_px,
                     double* out_host, mb_spline** spline_out);

/* ---- multi-GPU: one process per GPU, one NCCL communicator per context -------------------------------------------------
 * The raster shards by tile (machisplin.tiles.create boundaries, V73:1165-1256) or by row block; cells are independent given the
 * spline and model descriptors, so no raster strip ever moves.  What crosses NVLink: the K x K Gram of the k-fold residuals
 * (V73:329-333), and the descriptor of a fitted spline (24 bytes per knot) from the rank that ran fields::Tps to the ranks
 * that evaluate it on their own cells.  NCCL is bound at run time (dlopen of libnccl.so.2; MB_NCCL_LIB overrides): a single-GPU
 * host needs none.  Rank 0 creates an id (mb_comm_unique_id), the HOST ships its 128 bytes to the other processes (R: a file,
 * a socket, parallel::clusterExport; the Python tests: torch.distributed), every process calls mb_comm_init. */
#define MB_COMM_ID_BYTES 128
enum { MB_SUM = 0, MB_MAX = 1 };
int mb_comm_unique_id(void* id_out /* MB_COMM_ID_BYTES */);
int mb_comm_init(mb_ctx* ctx, int nranks, int rank, const void* id);
int mb_comm_destroy(mb_ctx* ctx);            /* also done by mb_shutdown */
int mb_comm_rank(const mb_ctx* ctx);         /* 0 without a communicator */
int mb_comm_size(const mb_ctx* ctx);         /* 1 without a communicator */
const char* mb_comm_backend(void);           /* which libnccl was bound, or why none could be */
/* in-place all-reduce of n host doubles (bench bookkeeping: max over ranks of a time; sums of R^2 terms) */
int mb_comm_allreduce_f64(mb_ctx* ctx, double* values_host, int n, int op);
/* a6 with the residual rows sharded over the ranks: G = sum_r R_r' R_r, every rank gets G (n_local may be 0) */
int mb_gram_allreduce(mb_ctx* ctx, const double* R_host, int n_local, int K, double* G_host);
/* ship a fitted spline from `root` to every rank (*s is input on the root, output elsewhere); max_knots >= its knot count */
int mb_spline_bcast(mb_ctx* ctx, mb_spline** s, int max_knots, int root);
/* mb_mltps_predict* for ONE ROW BLOCK of a raster sharded over the communicator (global spline, V73:748-753): g_block is the
 * extent + shape of this rank's rows (same xmin / xmax as the raster), e / cov / out belong to the block.  fields::Tps runs
 * once, on `root`, beside the per-cell ensemble kernels that every rank - the root too - runs on its own block; the spline
 * descriptor is broadcast and each rank finishes its own cells.  knots_xy / resid are read on the root only; n (the number of
 * observations) must be the same on every rank.  Without a communicator this is mb_mltps_predict* with tile_px = 0. */
int mb_mltps_predict_shard_dev(mb_ctx* ctx, const mb_grid* g_block, const mb_ensemble* e, const float* cov_dev, int C,
                               const double* knots_xy, const double* resid, int n, double lambda, int root,
                               double* out_dev, mb_spline** spline_out, void* stream);
int mb_mltps_predict_shard(mb_ctx* ctx, const mb_grid* g_block, const mb_ensemble* e, const float* cov_host, int C,
                           const double* knots_xy, const double* resid, int n, double lambda, int root,
                           double* out_host, mb_spline** spline_out);

/* ---- device memory helpers for hosts without a CUDA allocator (R) ---------------------- */
int mb_dev_alloc(mb_ctx* ctx, size_t bytes, void** out);
int mb_dev_free(mb_ctx* ctx, void* p);
int mb_h2d(mb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int mb_d2h(mb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- .C()-loadable twins: all-pointer arguments, void return, status[0] = MB_OK or a negative MB_E_* ---------
 * dyn.load("libmachisplin_b200.so"); .C("mbC_tps_surface", as.double(xy), as.double(y), as.integer(n), as.double(grid6), ...).
 * One process-wide context (device $MB_DEVICE, default 0) is created on first use; grid6 = c(xmin, xmax, ymin, ymax, nrow, ncol).
 * Covered: the call sites whose inputs / outputs are plain numeric vectors.  The per-cell ensemble needs the model descriptors
 * and device handles and goes through the .Call shim (r/mb_shim.c). */
/* V73:329-333: G = R'R, R n x K column-major */
void mbC_gram(const double* R, const int* n, const int* K, double* G, int* status);
/* V73:722-753 (+ 649-895 when tile_px > 0): fields::Tps(xy, y) + terra::interpolate over the grid, values in terra cell order;
 * lambda < 0 = GCV; lambda_out (length 1) gets the selected lambda of the global spline (NaN in tiled mode) */
void mbC_tps_surface(const double* xy, const double* y, const int* n, const double* grid6, const double* lambda,
                     const int* tile_px, double* out, double* lambda_out, int* status);
/* V73:1392-1548: wins = 4 integers (r0, r1, c0, c1) per tile, tiles_flat = the tile rasters one after another, row-major */
void mbC_tiles_merge(const double* grid6, const int* nC, const int* nR, const int* wins, const double* tiles_flat, double* out,
                     int* status);
void mbC_last_error(char** buf);     /* message of the last failure, copied into buf[0] up to its allocated length */
void mbC_shutdown(void);

/* ---- per-kernel device timing (bench bookkeeping) ---------------------------------------- */
/* When enabled, every kernel launch of this context is bracketed by CUDA events on the launching
 * stream.  mb_timing_collect synchronises, sums the elapsed time per kernel name and resets the
 * log; it returns the number of distinct names written (<= cap).  names[i] points to static strings. */
int mb_timing_enable(mb_ctx* ctx, int on);
int mb_timing_collect(mb_ctx* ctx, int cap, const char** names, double* total_ms, int64_t* launches);

/* Named diagnostic vectors; returns the number of values written or a negative MB_E_* code.
 * "sytrd_phase_ms": time per phase of the last tridiagonalisation seen by CTA 0 while timing was enabled
 * (P1, barrier, P2, barrier, P3, barrier, panel update, barrier). */
int mb_debug_values(mb_ctx* ctx, const char* name, double* out, int cap);

/* Tunables of the fast evaluator (0 = automatic). */
int mb_set_fast_eval_params(mb_ctx* ctx, int cheb_p, int leaf_cols, int leaf_rows);
/* Named integer tunables (0 = automatic):
 * "ens_overlap" = 2 (default) the forest kernel, then the tensor-pipe ksvm kernel; 1 side by side on two streams when both are
 * kept (measured slower); "svm_ctas_per_sm" = persistent grid of the ksvm kernel in that mode (default 2); "ens_tma" = 1 (default)
 * covariate tiles of the ksvm kernel by TMA tensor copies (cp.async.bulk.tensor), 2 plain loads;
 * "gc_split" = SM count of the fit partition when mb_mltps_predict* splits the device with CUDA green contexts (stage 1 of the GCV fit
 * on one partition, the forest kernel on the other from the start; 0 = default 64, -1 = no partitions: the ensemble waits for stage 1;
 * the partitions are created once per context, with the size in force at the first call that uses them); "gc_share" = percent of
 * the raster's rows whose forest kernel runs on the ensemble partition beside stage 1, the rest follows on all SMs (0 = default = all rows);
 * "leaf_impl" = TPS-only grid evaluation: 1 (default) one warp per leaf box, 2 one CTA per leaf box;
 * "leaf_tma" = 1 (default) the grid-evaluation kernel fetches the accumulator tile of a box with one 2-D tensor copy, 2 row by row;
 * "ens_order" = 1 (default) the forest kernel runs before the ksvm kernel, 2 after it;
 * "tree_levels" = 1 forest kernel with the CTA-level interval prune only, 2 (default) + the warp-level prune;
 * "eval_precision" = leaf kernel code path of the fast evaluator (1 = float64 only, 2 = force mixed);
 * "sytrd_mode" = tridiagonalisation of the GCV fit: 0 / 3 = two-stage (band reduction + bulge chasing, the default),
 * 1 = one-stage persistent kernel, 2 = one-stage with one kernel per phase; "sytrd_ctas_per_sm" = grid size of the
 * one-stage persistent kernel; "sbr_qr_grid" = 1 makes the panel QR of the two-stage path use the software grid
 * barrier instead of a thread-block cluster; "sbr_qr_impl" = 1 / 2 keeps the panel rows in shared memory / registers
 * (0 = chosen by cluster size); "sbr_debug" = 1 keeps the band matrix for mb_debug_values("sbr_band");
 * "svm_impl" (before mb_ensemble_create) = ksvm dot products: 3 on the tensor pipe with FP16 split operands (two HMMA.16816 per 16 cells x 8
 * support vectors; the default for P <= 8, falls back to 1 if a scaled support vector leaves the half range), 1 on the tensor pipe as
 * 3 x TF32 (three HMMA.1688), 2 packed FP32 (the only one for P > 8); 0 = default;
 * "coef_impl" = 0 coefficients from the band form of the two-stage reduction when cond(M + lambda I) <= 1e8 (default), 1 whenever
 * that form exists, 2 always the dense Cholesky of M + lambda I;
 * "sbr_chase_impl" = bulge chase: 1 flags with watcher / publisher warps, 2 flags with three warps per sweep, 3 tagged band elements
 * (every double travels as two 8-byte words carrying the number of the sweep that wrote it: no flags, no fences), 0 (default) = 3
 * for a fit that has the GPU to itself and 2 inside mb_mltps_predict* with an ensemble (smallest register footprint beside the
 * per-cell kernels); all three are bit-identical;
 * "sbr_chase_ctas" = cap of the bulge chase's grid (0 = one CTA per sweep that can be in flight);
 * "sbr_chase_sleep" = nanoseconds of back-off in the spin loops of the chase (0 = none; measured: no effect);
 * "defer_ensemble" = 0 starts the per-cell ensemble kernels of mb_mltps_predict* before the fit instead of behind stage 1
 * of its tridiagonalisation; "eigen_impl" = 1 routes the GCV fit through cuSOLVER Dsyevd (validation of the in-house path only). */
int mb_set_param(mb_ctx* ctx, const char* name, int value);

/* ---- GeoTIFF in / out: terra::rast(path) (README Example 1, V73:68-70) and terra::writeRaster() (V73:1011, 1020) -----------------
 * Host-only (no context, no GPU) except mb_tiff_read_f32_dev.  Readers accept classic TIFF, striped or tiled, chunky or planar bands, 8 / 16 / 32-bit integers
 * and 32 / 64-bit floats, compression none / LZW / PackBits, predictor 1 / 2 - which covers the rasters the reference bundles
 * (INT16, 128 x 128 tiles, none or LZW, GDAL_NODATA) and what terra writes by default.  Everything else is refused with a message.
 * Cell order is terra's: row-major from the NW corner.  NoData cells come back as NaN (terra: NA). */
typedef struct {
  mb_grid grid;            /* extent from ModelPixelScale + ModelTiepoint (has_georef), else [0, ncol] x [0, nrow] */
  int32_t nbands, bits, sample_format /* 1 unsigned, 2 signed, 3 float */, compression, predictor;
  int32_t tiled, chunk_w, chunk_h;     /* tile size, or (image width, rows per strip) */
  int32_t has_georef, has_nodata, epsg /* 0 = not given */;
  double nodata;
} mb_tiff_meta;
int mb_tiff_info(const char* path, mb_tiff_meta* out);
/* one band (0-based) as float32, nrow * ncol values; tiles / strips are decoded by nthreads host threads (0 = all) */
int mb_tiff_read_f32(const char* path, int band, float* out, int nthreads);
/* The same band decoded ON THE DEVICE into a float32 plane in HBM (out_dev: nrow * ncol floats), ready for mb_mltps_predict_dev:
 * the file's compressed bytes cross PCIe through pinned bounce buffers, the GPU undoes LZW (one warp per tile / strip), the
 * horizontal predictor, the sample type and NoData.  Little-endian files with compression none / LZW take this path; big-endian
 * files, PackBits and predictor 2 on pixel-interleaved bands fall back to the host decoder + one upload (stats says which).
 * Synchronous: the plane is complete when the call returns.  Values are bit-identical to mb_tiff_read_f32. */
typedef struct {
  int32_t decoded_on_gpu;   /* 1 = LZW / predictor / conversion ran on the device, 0 = host decoder + upload */
  int32_t chunks;           /* tiles or strips of the band */
  int64_t h2d_bytes;        /* bytes that crossed PCIe */
} mb_tiff_dev_stats;
int mb_tiff_read_f32_dev(mb_ctx* ctx, const char* path, int band, float* out_dev, int nthreads, void* stream, mb_tiff_dev_stats* stats /* may be NULL */);
/* FLT4S GeoTIFF (terra::writeRaster's default datatype), 256 x 256 tiles, compression 1 (none) or 5 (LZW), NaN = NoData;
 * epsg > 0 adds the CRS key (4326 for the LONG / LAT rasters of the reference) */
int mb_tiff_write_f32(const char* path, const mb_grid* g, const float* data, int compression, int epsg, int nthreads);
int mb_tiff_write_f64(const char* path, const mb_grid* g, const double* data, int compression, int epsg, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* MACHISPLIN_B200_H */
