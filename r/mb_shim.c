/*
 * .Call shim between R and libmachisplin_b200.so (include/machisplin_b200.h).
 * Build where R exists with
 *     R CMD SHLIB mb_shim.c -I../include -L../machisplin_b200 -lmachisplin_b200
 * The build image has no R: there the shim is compiled against the stub R API of tests/r_stub/ (fake SEXPs) and every wrapper is
 * called by tests/test_rshim_*.py - on the GPU box through to the kernels.
 * Every wrapper converts SEXP <-> plain pointers, calls ONE C-ABI entry point and turns a non-zero status
 * into an R error after the C frames have unwound.  Device-resident handles (context, spline, ensemble) are
 * R external pointers with finalizers.  R owns every host array for the duration of the call.
 */
#include <R.h>
#include <Rinternals.h>
#include <string.h>
#include "machisplin_b200.h"

#define MB_CHECK(call) do { int rc_ = (call); if (rc_ != MB_OK) Rf_error("machisplin_b200: %s", mb_last_error()); } while (0)

static void ctx_fin(SEXP p) { mb_shutdown((mb_ctx*)R_ExternalPtrAddr(p)); R_ClearExternalPtr(p); }
static void spl_fin(SEXP p) { mb_spline_free((mb_spline*)R_ExternalPtrAddr(p)); R_ClearExternalPtr(p); }
static void ens_fin(SEXP p) { mb_ensemble_free((mb_ensemble*)R_ExternalPtrAddr(p)); R_ClearExternalPtr(p); }

/* `keep` (the context's external pointer, or R_NilValue for the context itself) becomes the `prot` field of the handle: R then
 * keeps the context alive at least as long as any spline / ensemble created from it, whatever order the GC finalizes in. */
static SEXP wrap(void* h, R_CFinalizer_t fin, SEXP keep) {
  SEXP p = PROTECT(R_MakeExternalPtr(h, R_NilValue, keep));
  R_RegisterCFinalizerEx(p, fin, TRUE);
  UNPROTECT(1);
  return p;
}
static mb_grid grid_of(SEXP g) {           /* c(xmin, xmax, ymin, ymax, nrow, ncol) = c(ext(r)[1:4], dim(r)[1:2]) */
  const double* v = REAL(g);
  mb_grid out = {v[0], v[1], v[2], v[3], (int32_t)v[4], (int32_t)v[5]};
  return out;
}

SEXP mbR_init(SEXP device) {
  mb_ctx* ctx = NULL;
  MB_CHECK(mb_init(Rf_asInteger(device), &ctx));
  return wrap(ctx, ctx_fin, R_NilValue);
}

/* fields::Tps(x, Y) -> external pointer of the fitted spline (V73:722, V73:751).  lambda < 0 = GCV. */
SEXP mbR_tps_fit(SEXP ctx, SEXP xy, SEXP y, SEXP lambda) {
  const int n = Rf_nrows(xy);
  mb_spline* s = NULL;
  MB_CHECK(mb_tps_fit((mb_ctx*)R_ExternalPtrAddr(ctx), REAL(xy), REAL(y), n, 1, Rf_asReal(lambda), &s));
  return wrap(s, spl_fin, ctx);
}

/* terra::interpolate(rast(template), Tps): values in terra cell order (V73:726, V73:753). */
SEXP mbR_tps_eval(SEXP ctx, SEXP spline, SEXP grid) {
  mb_grid g = grid_of(grid);
  mb_window w = {0, g.nrow, 0, g.ncol};
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)g.nrow * g.ncol));
  MB_CHECK(mb_tps_eval((mb_ctx*)R_ExternalPtrAddr(ctx), (mb_spline*)R_ExternalPtrAddr(spline), &g, &w, MB_EVAL_FAST,
                       REAL(out)));
  UNPROTECT(1);
  return out;
}

/* Flat model descriptors (a named list built by mb_export_models() in machisplin_b200.R) -> ensemble handle. */
static const double* dbl(SEXP l, const char* name);
static const int* int_(SEXP l, const char* name);
static SEXP elt(SEXP l, const char* name) {
  SEXP names = Rf_getAttrib(l, R_NamesSymbol);
  for (R_xlen_t i = 0; i < Rf_xlength(l); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(l, i);
  return R_NilValue;
}
static const double* dbl(SEXP l, const char* name) { SEXP e = elt(l, name); return e == R_NilValue ? NULL : REAL(e); }
static const int* int_(SEXP l, const char* name) { SEXP e = elt(l, name); return e == R_NilValue ? NULL : INTEGER(e); }
static double num(SEXP l, const char* name, double dflt) { SEXP e = elt(l, name); return e == R_NilValue ? dflt : Rf_asReal(e); }

SEXP mbR_ensemble_create(SEXP ctx, SEXP grid, SEXP d, SEXP kept, SEXP w, SEXP w_total) {
  mb_grid g = grid_of(grid);
  mb_models m;
  memset(&m, 0, sizeof m);
  m.P = (int32_t)num(d, "P", 0);
  m.gam_coef = dbl(d, "gam_coef");
  m.nn_wts = dbl(d, "nn_wts"); m.nn_H = (int32_t)num(d, "nn_H", 0); m.nn_max2 = num(d, "nn_max2", 1); m.nn_min = num(d, "nn_min", 0);
  m.mars_T = (int32_t)num(d, "mars_T", 0); m.mars_dirs = (const int8_t*)(elt(d, "mars_dirs") == R_NilValue ? NULL : RAW(elt(d, "mars_dirs")));
  m.mars_cuts = dbl(d, "mars_cuts"); m.mars_coef = dbl(d, "mars_coef");
  m.svm_S = (int32_t)num(d, "svm_S", 0); m.svm_sv = dbl(d, "svm_sv"); m.svm_alpha = dbl(d, "svm_alpha");
  m.svm_b = num(d, "svm_b", 0); m.svm_sigma = num(d, "svm_sigma", 0);
  m.svm_x_center = dbl(d, "svm_x_center"); m.svm_x_scale = dbl(d, "svm_x_scale");
  m.svm_y_center = num(d, "svm_y_center", 0); m.svm_y_scale = num(d, "svm_y_scale", 1);
  m.rf_ntree = (int32_t)num(d, "rf_ntree", 0); m.rf_nrnodes = (int32_t)num(d, "rf_nrnodes", 0);
  m.rf_left = int_(d, "rf_left"); m.rf_right = int_(d, "rf_right");
  m.rf_status = (const int8_t*)(elt(d, "rf_status") == R_NilValue ? NULL : RAW(elt(d, "rf_status")));
  m.rf_bestvar = int_(d, "rf_bestvar"); m.rf_split = dbl(d, "rf_split"); m.rf_nodepred = dbl(d, "rf_nodepred");
  m.gbm_ntrees = (int32_t)num(d, "gbm_ntrees", 0); m.gbm_initF = num(d, "gbm_initF", 0); m.gbm_tree_off = int_(d, "gbm_tree_off");
  m.gbm_splitvar = int_(d, "gbm_splitvar"); m.gbm_splitcode = dbl(d, "gbm_splitcode");
  m.gbm_left = int_(d, "gbm_left"); m.gbm_right = int_(d, "gbm_right"); m.gbm_missing = int_(d, "gbm_missing");
  mb_ensemble* e = NULL;
  MB_CHECK(mb_ensemble_create((mb_ctx*)R_ExternalPtrAddr(ctx), &g, &m, CHAR(STRING_ELT(kept, 0)), REAL(w),
                              Rf_asReal(w_total), &e));
  return wrap(e, ens_fin, ctx);
}

/* mltps parts 2-5 for one response (V73:442-932): cov = float32 planes packed in a raw vector (C x nrow x ncol,
 * terra cell order), knots_xy n x 2, resid n.  Returns the final raster values. */
SEXP mbR_mltps_predict(SEXP ctx, SEXP grid, SEXP ens, SEXP cov_raw, SEXP C, SEXP knots_xy, SEXP resid, SEXP lambda,
                       SEXP tile_px) {
  mb_grid g = grid_of(grid);
  const int tps = knots_xy != R_NilValue;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)g.nrow * g.ncol));
  MB_CHECK(mb_mltps_predict((mb_ctx*)R_ExternalPtrAddr(ctx), &g,
                            ens == R_NilValue ? NULL : (mb_ensemble*)R_ExternalPtrAddr(ens),
                            cov_raw == R_NilValue ? NULL : (const float*)RAW(cov_raw), Rf_asInteger(C),
                            tps ? REAL(knots_xy) : NULL, tps ? REAL(resid) : NULL, tps ? Rf_nrows(knots_xy) : 0,
                            Rf_asReal(lambda), Rf_asInteger(tile_px), REAL(out), NULL));
  UNPROTECT(1);
  return out;
}

/* G = R'R for the ensemble-weight objective (V73:329-333, 369-373). */
SEXP mbR_gram(SEXP ctx, SEXP R) {
  const int n = Rf_nrows(R), K = Rf_ncols(R);
  SEXP G = PROTECT(Rf_allocMatrix(REALSXP, K, K));
  MB_CHECK(mb_gram((mb_ctx*)R_ExternalPtrAddr(ctx), REAL(R), n, K, REAL(G)));
  UNPROTECT(1);
  return G;
}

/* terra::rast(path) for the covariate stack (README Example 1, V73:68-70): the rasters are decoded by the library's host
 * threads straight into float32 planes (C x nrow x ncol, terra cell order, NoData = NaN) packed in a raw vector - the layout
 * mbR_mltps_predict takes - without the round trip through terra::values() and writeBin().  Returns
 * list(grid = c(xmin, xmax, ymin, ymax, nrow, ncol), cov = raw). */
SEXP mbR_read_stack(SEXP paths) {
  const int C = Rf_length(paths);
  mb_tiff_meta m0;
  MB_CHECK(mb_tiff_info(CHAR(STRING_ELT(paths, 0)), &m0));
  const R_xlen_t ncell = (R_xlen_t)m0.grid.nrow * m0.grid.ncol;
  SEXP cov = PROTECT(Rf_allocVector(RAWSXP, ncell * C * (R_xlen_t)sizeof(float)));
  for (int k = 0; k < C; ++k) {
    mb_tiff_meta m;
    MB_CHECK(mb_tiff_info(CHAR(STRING_ELT(paths, k)), &m));
    if (m.grid.nrow != m0.grid.nrow || m.grid.ncol != m0.grid.ncol) Rf_error("machisplin_b200: extents do not match");
    MB_CHECK(mb_tiff_read_f32(CHAR(STRING_ELT(paths, k)), 0, (float*)RAW(cov) + ncell * k, 0));
  }
  SEXP grid = PROTECT(Rf_allocVector(REALSXP, 6));
  REAL(grid)[0] = m0.grid.xmin; REAL(grid)[1] = m0.grid.xmax; REAL(grid)[2] = m0.grid.ymin; REAL(grid)[3] = m0.grid.ymax;
  REAL(grid)[4] = m0.grid.nrow; REAL(grid)[5] = m0.grid.ncol;
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, grid);
  SET_VECTOR_ELT(out, 1, cov);
  UNPROTECT(3);
  return out;
}

/* terra::writeRaster(x, filename, overwrite = TRUE) (V73:1011, 1020): FLT4S GeoTIFF, LZW, NA = NoData. */
SEXP mbR_write_raster(SEXP path, SEXP grid, SEXP values, SEXP epsg) {
  mb_grid g = grid_of(grid);
  MB_CHECK(mb_tiff_write_f64(CHAR(STRING_ELT(path, 0)), &g, REAL(values), 5, Rf_asInteger(epsg), 0));
  return R_NilValue;
}

/* ---- machisplin.tiles.* and the multi-GPU path (one R process per GPU) ------------------------------------------------------ */
static void wins_of(SEXP wins, int nt, mb_window* out) {      /* integer matrix 4 x ntiles: (r0, r1, c0, c1) per tile, 0-based, half-open */
  const int* v = INTEGER(wins);
  for (int t = 0; t < nt; ++t) { out[t].r0 = v[4 * t]; out[t].r1 = v[4 * t + 1]; out[t].c0 = v[4 * t + 2]; out[t].c1 = v[4 * t + 3]; }
}

/* machisplin.tiles.merge(tiles, in.ncol, in.nrow) (V73:1392-1548): tiles = list of numeric vectors (tile rasters in terra cell
 * order, row-major from the SW tile), wins as above.  Returns the merged raster values. */
SEXP mbR_tiles_merge(SEXP ctx, SEXP grid, SEXP nC, SEXP nR, SEXP wins, SEXP tiles) {
  mb_grid g = grid_of(grid);
  const int nt = Rf_asInteger(nC) * Rf_asInteger(nR);
  if (nt < 1 || nt > 4096 || Rf_length(tiles) != nt || Rf_length(wins) != 4 * nt) Rf_error("machisplin_b200: tiles / wins do not match in.ncol x in.nrow");
  mb_window w[4096];
  const double* ptr[4096];
  wins_of(wins, nt, w);
  for (int t = 0; t < nt; ++t) ptr[t] = REAL(VECTOR_ELT(tiles, t));
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)g.nrow * g.ncol));
  MB_CHECK(mb_tiles_merge((mb_ctx*)R_ExternalPtrAddr(ctx), &g, Rf_asInteger(nC), Rf_asInteger(nR), w, ptr, REAL(out)));
  UNPROTECT(1);
  return out;
}

/* the window of cells tile t owns when the merge is spread over the GPUs (mb_tiles_owned_window): c(r0, r1, c0, c1) */
SEXP mbR_tiles_owned_window(SEXP grid, SEXP nC, SEXP nR, SEXP wins, SEXP t) {
  mb_grid g = grid_of(grid);
  const int nt = Rf_asInteger(nC) * Rf_asInteger(nR);
  if (nt < 1 || nt > 4096 || Rf_length(wins) != 4 * nt) Rf_error("machisplin_b200: wins does not match in.ncol x in.nrow");
  mb_window w[4096], own;
  wins_of(wins, nt, w);
  MB_CHECK(mb_tiles_owned_window(&g, Rf_asInteger(nC), Rf_asInteger(nR), w, Rf_asInteger(t), &own));
  SEXP out = PROTECT(Rf_allocVector(INTSXP, 4));
  INTEGER(out)[0] = own.r0; INTEGER(out)[1] = own.r1; INTEGER(out)[2] = own.c0; INTEGER(out)[3] = own.c1;
  UNPROTECT(1);
  return out;
}

/* communicator of the context: rank 0 draws the id (raw vector of MB_COMM_ID_BYTES), the R hosts ship it (file, socket), every
 * rank calls mbR_comm_init.  What the reference's deleted snowfall path did with sockets (old/...V69.R:937-968) */
SEXP mbR_comm_unique_id(void) {
  SEXP id = PROTECT(Rf_allocVector(RAWSXP, MB_COMM_ID_BYTES));
  MB_CHECK(mb_comm_unique_id(RAW(id)));
  UNPROTECT(1);
  return id;
}
SEXP mbR_comm_init(SEXP ctx, SEXP nranks, SEXP rank, SEXP id) {
  if (Rf_length(id) != MB_COMM_ID_BYTES) Rf_error("machisplin_b200: the communicator id must have %d bytes", MB_COMM_ID_BYTES);
  MB_CHECK(mb_comm_init((mb_ctx*)R_ExternalPtrAddr(ctx), Rf_asInteger(nranks), Rf_asInteger(rank), RAW(id)));
  return R_NilValue;
}
/* G = sum over the ranks of R_r' R_r: the objective of V73:329-333 with the cross-validation rows sharded */
SEXP mbR_gram_allreduce(SEXP ctx, SEXP R) {
  const int n = Rf_nrows(R), K = Rf_ncols(R);
  SEXP G = PROTECT(Rf_allocMatrix(REALSXP, K, K));
  MB_CHECK(mb_gram_allreduce((mb_ctx*)R_ExternalPtrAddr(ctx), REAL(R), n, K, REAL(G)));
  UNPROTECT(1);
  return G;
}
/* mbR_mltps_predict for THIS rank's row block of a raster sharded over the communicator: grid = the block's extent and shape,
 * fields::Tps runs on `root` only (knots_xy / resid may be NULL elsewhere; n = number of observations everywhere) */
SEXP mbR_mltps_predict_shard(SEXP ctx, SEXP grid, SEXP ens, SEXP cov_raw, SEXP C, SEXP knots_xy, SEXP resid, SEXP n, SEXP lambda,
                             SEXP root) {
  mb_grid g = grid_of(grid);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)g.nrow * g.ncol));
  MB_CHECK(mb_mltps_predict_shard((mb_ctx*)R_ExternalPtrAddr(ctx), &g,
                                  ens == R_NilValue ? NULL : (mb_ensemble*)R_ExternalPtrAddr(ens),
                                  cov_raw == R_NilValue ? NULL : (const float*)RAW(cov_raw), Rf_asInteger(C),
                                  knots_xy == R_NilValue ? NULL : REAL(knots_xy), resid == R_NilValue ? NULL : REAL(resid),
                                  Rf_asInteger(n), Rf_asReal(lambda), Rf_asInteger(root), REAL(out), NULL));
  UNPROTECT(1);
  return out;
}

