# R side of the drop-in: replacements for the call sites of machisplin.mltps() that the engine takes over.
# NOT exercised in the build image (no R there).  V73:n = R/ensemble.machine.learning.thin.plate.splines.V73.R line n.
#
#   dyn.load("mb_shim.so"); mb <- mb_open(0)
#
mb_open <- function(device = 0L) .Call("mbR_init", as.integer(device))
mb_grid <- function(r) as.numeric(c(terra::ext(r)[1:4], terra::nrow(r), terra::ncol(r)))

# V73:722 / V73:751 + V73:726 / V73:753 :
#   mod.tps.elev <- fields::Tps(dat[, c(LONG, LAT)], res.FINAL); tps <- terra::interpolate(terra::rast(rast_stack), mod.tps.elev)
mb_tps_surface <- function(mb, xy, resid, template) {
  sp <- .Call("mbR_tps_fit", mb, as.matrix(xy), as.numeric(resid), -1)        # lambda < 0: fields' GCV search
  out <- terra::rast(template[[1]])
  terra::values(out) <- .Call("mbR_tps_eval", mb, sp, mb_grid(template))       # terra cell order
  out
}

# Flat descriptors of the fitted models (SURVEY.md Appendix B): what mb_ensemble_create reads.
mb_export_models <- function(P, gam = NULL, nn = NULL, nn.max2 = 1, nn.min = 0, mars = NULL, svm = NULL, rf = NULL,
                             brt = NULL, brt.ntrees = NULL) {
  d <- list(P = P)
  if (!is.null(gam)) d$gam_coef <- as.numeric(coef(gam))                      # V73:600: parametric formula only
  if (!is.null(nn)) { d$nn_wts <- as.numeric(nn$wts); d$nn_H <- nn$n[2]; d$nn_max2 <- nn.max2; d$nn_min <- nn.min }   # V73:463-470
  if (!is.null(mars)) {                                                       # V73:539
    st <- mars$selected.terms
    d$mars_T <- length(st); d$mars_dirs <- as.raw(bitwAnd(as.integer(t(mars$dirs[st, , drop = FALSE])), 255L))
    d$mars_cuts <- as.numeric(t(mars$cuts[st, , drop = FALSE])); d$mars_coef <- as.numeric(mars$coefficients)
  }
  if (!is.null(svm)) {                                                        # V73:560 (kernlab::ksvm, eps-svr, rbfdot)
    sc <- kernlab::scaling(svm)
    d$svm_S <- kernlab::nSV(svm); d$svm_sv <- as.numeric(t(kernlab::xmatrix(svm))); d$svm_alpha <- as.numeric(kernlab::coef(svm))
    d$svm_b <- kernlab::b(svm); d$svm_sigma <- kernlab::kpar(kernlab::kernelf(svm))$sigma
    d$svm_x_center <- sc$x.scale$`scaled:center`; d$svm_x_scale <- sc$x.scale$`scaled:scale`
    d$svm_y_center <- sc$y.scale$`scaled:center`; d$svm_y_scale <- sc$y.scale$`scaled:scale`
  }
  if (!is.null(rf)) {                                                         # V73:517 (randomForest regression)
    f <- rf$forest
    d$rf_ntree <- f$ntree; d$rf_nrnodes <- f$nrnodes
    d$rf_left <- as.integer(f$leftDaughter); d$rf_right <- as.integer(f$rightDaughter)      # nrnodes x ntree, column-major = tree-major
    d$rf_status <- as.raw(bitwAnd(as.integer(f$nodestatus), 255L)); d$rf_bestvar <- as.integer(f$bestvar)
    d$rf_split <- as.numeric(f$xbestsplit); d$rf_nodepred <- as.numeric(f$nodepred)
  }
  if (!is.null(brt)) {                                                        # V73:493-501 (gbm, gaussian)
    nt <- if (is.null(brt.ntrees)) brt$gbm.call$best.trees else brt.ntrees
    tr <- brt$trees[seq_len(nt)]
    d$gbm_ntrees <- nt; d$gbm_initF <- brt$initF
    d$gbm_tree_off <- as.integer(c(0, cumsum(vapply(tr, function(t) length(t[[1]]), 0L))))
    d$gbm_splitvar <- as.integer(unlist(lapply(tr, `[[`, 1))); d$gbm_splitcode <- as.numeric(unlist(lapply(tr, `[[`, 2)))
    d$gbm_left <- as.integer(unlist(lapply(tr, `[[`, 3))); d$gbm_right <- as.integer(unlist(lapply(tr, `[[`, 4)))
    d$gbm_missing <- as.integer(unlist(lapply(tr, `[[`, 5)))
  }
  d
}

# V73:442-932 in one call: replaces the eleven terra::predict() calls, the weighted accumulation (V73:471-619),
# fields::Tps + terra::interpolate (V73:722-753 incl. the internal 1500-px tiling when the raster needs it) and the
# NA-propagating final sum (V73:906-907).  `kept`, `w` = round(p, 2) of the kept models, `w.total` = sum of ALL p (V73:337).
mb_mltps_predict <- function(mb, rast_stack, n.covars, models, kept, w, w.total, xy = NULL, res.FINAL = NULL, tile.px = 1500L) {
  g <- mb_grid(rast_stack)
  # terra::values() is ncell x nlyr (cell order within a column = terra cell order), so its column-major storage already is
  # the plane layout [C][nrow][ncol] the engine expects - no transpose (a t() here would interleave the layers per cell)
  cov <- writeBin(as.numeric(terra::values(rast_stack[[seq_len(n.covars)]])), raw(), size = 4)      # float32 planes
  ens <- .Call("mbR_ensemble_create", mb, g, models, kept, as.numeric(w), as.numeric(w.total))
  v <- .Call("mbR_mltps_predict", mb, g, ens, cov, as.integer(n.covars),
             if (is.null(xy)) NULL else as.matrix(xy), res.FINAL, -1, as.integer(tile.px))
  out <- terra::rast(rast_stack[[1]]); terra::values(out) <- v
  out
}

# V73:329-333 / 369-373: the objective handed to optimx becomes a 6 x 6 quadratic form.
mb_rss_objective <- function(mb, R) { G <- .Call("mbR_gram", mb, as.matrix(R)); function(k) drop(k %*% G %*% k) / sum(k)^2 }

# README Example 1 / V73:68-70 without terra in the data path: the covariate GeoTIFFs are decoded by the library into the float32
# planes mb_mltps_predict() uploads (list(grid, cov)); V73:1011 / 1020: the final raster as a FLT4S GeoTIFF.
mb_read_stack <- function(paths) .Call("mbR_read_stack", as.character(paths))
mb_write_raster <- function(path, grid, values, epsg = 4326L) invisible(.Call("mbR_write_raster", as.character(path), as.numeric(grid), as.numeric(values), as.integer(epsg)))

# machisplin.tiles.merge (V73:1392-1548) on the $final rasters of the tiles; wins = 4 x ntiles integer matrix of 0-based half-open
# cell windows (r0, r1, c0, c1), tiles in the reference's order (row-major from the SW tile)
mb_tiles_merge <- function(mb, template, in.ncol, in.nrow, wins, tiles)
  .Call("mbR_tiles_merge", mb, mb_grid(template), as.integer(in.ncol), as.integer(in.nrow), as.integer(wins), lapply(tiles, as.numeric))
mb_tiles_owned_window <- function(template, in.ncol, in.nrow, wins, t)
  .Call("mbR_tiles_owned_window", mb_grid(template), as.integer(in.ncol), as.integer(in.nrow), as.integer(wins), as.integer(t))

# Multi-GPU, one R process per GPU (what the reference's deleted snowfall path did, old/...V69.R:937-968): rank 0 draws the id,
# the hosts ship its bytes (saveRDS / socket), every rank joins; then the Gram of V73:329-333 is summed over the ranks and every
# rank predicts its own row block of the raster while fields::Tps runs on `root` only.
mb_comm_unique_id <- function() .Call("mbR_comm_unique_id")
mb_comm_init <- function(mb, nranks, rank, id) invisible(.Call("mbR_comm_init", mb, as.integer(nranks), as.integer(rank), id))
mb_rss_objective_sharded <- function(mb, R.local) { G <- .Call("mbR_gram_allreduce", mb, as.matrix(R.local)); function(k) drop(k %*% G %*% k) / sum(k)^2 }
mb_mltps_predict_block <- function(mb, block_stack, n.covars, models, kept, w, w.total, xy, res.FINAL, n, root = 0L) {
  g <- mb_grid(block_stack)
  cov <- writeBin(as.numeric(terra::values(block_stack[[seq_len(n.covars)]])), raw(), size = 4)
  ens <- .Call("mbR_ensemble_create", mb, g, models, kept, as.numeric(w), as.numeric(w.total))
  v <- .Call("mbR_mltps_predict_shard", mb, g, ens, cov, as.integer(n.covars), if (is.null(xy)) NULL else as.matrix(xy), res.FINAL,
             as.integer(n), -1, as.integer(root))
  out <- terra::rast(block_stack[[1]]); terra::values(out) <- v
  out
}

