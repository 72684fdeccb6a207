// Cross-file declarations of the device implementations behind the C ABI.
#pragma once
#include "common.cuh"

namespace mb {

// tps_fit.cu - fields::Tps (V73:722, 751)
void tps_fit(mb_ctx* ctx, const double* xy, const double* y, int n, int L, double lambda, mb_spline** out);
void fit_release(mb_ctx* ctx);   // library handles owned on behalf of the context

// sytrd.cu - in-house tridiagonalisation + Sturm bisection for the GCV fit
void sym_tridiag_eig(mb_ctx* ctx, double* A, int ld, int m, double* z_dev, int L, std::vector<double>& diag,
                     std::vector<double>& off, std::vector<double>& eta, cudaStream_t st);

// sbr.cu - two-stage alternative (dense -> band -> tridiagonal), mb_set_param("sytrd_mode", 3); d, e on the device
void sym_band_tridiag(mb_ctx* ctx, double* A, int ld, int m, double* z_dev, int L, double* d_dev, double* e_dev,
                      cudaStream_t st);

// (M + lambda I)^-1 z_r from the band form kept by the last sym_band_tridiag of this call (ctx->band_form.valid): block band
// Cholesky + back-transformation by the stored panel reflectors.  out_dev: m doubles.  False if the form cannot be used.
bool band_coefficients(mb_ctx* ctx, double lambda, int rhs, double* out_dev, cudaStream_t st);

// ensemble.cu - terra::predict x6 + weighted sum (V73:468-619) + part-5 combine (V73:906-907)
mb_ensemble* ensemble_create(mb_ctx* ctx, const mb_grid& g, const mb_models& m, const char* kept, const double* w,
                             double w_total);
void ensemble_free(mb_ensemble* e);
mb_grid ensemble_grid(const mb_ensemble* e);
void ensemble_eval(mb_ctx* ctx, const mb_ensemble* e, const float* cov_dev, int C, const mb_spline* spline,
                   const double* tps_surface_dev, const mb_window* w, double* out_dev, cudaStream_t st);
// the two halves of ensemble_eval, exposed so that mltps_predict can run the TPS fit between them
int ensemble_ncov(const mb_ensemble* e);
// acc: padded accumulator layout (AccFuse, common.cuh): acc_stride(w) * acc_rows(w) doubles
void ensemble_accumulate(mb_ctx* ctx, const mb_ensemble* e, const float* cov_dev, int C, const mb_window& w,
                         double* acc, cudaStream_t st, int part = 0);
bool ensemble_has_forest_kernel(const mb_ensemble* e);   // a forest kernel opens the chain and something follows it
void ensemble_finish(mb_ctx* ctx, const mb_ensemble* e, const mb_spline* spline, const double* tps_surface_dev,
                     const mb_window& w, const double* acc, double* out_dev, cudaStream_t st);
void ensemble_predict_points(mb_ctx* ctx, const mb_ensemble* e, const double* X, int n, double* out_host);

// greenctx.cu - SM partitions for the fit / ensemble overlap of mb_mltps_predict*
bool greenctx_setup(mb_ctx* ctx, int fit_sms);
void greenctx_release(mb_ctx* ctx);

// comm.cu - NCCL communicator of the context; every collective of the path
void comm_release(mb_ctx* ctx);
void comm_allreduce_f64(mb_ctx* ctx, double* dev, int n, int op, cudaStream_t st);
void comm_allgather_f64(mb_ctx* ctx, const double* send_dev, double* recv_dev, size_t count_per_rank, cudaStream_t st);
void comm_allreduce_i32_min(mb_ctx* ctx, int* dev, int n, cudaStream_t st);
struct CommMsg {
  int peer;          // the other rank
  double* ptr;       // device buffer (source of a send, destination of a receive)
  size_t count;      // doubles
  bool send;
};
void comm_exchange_f64(mb_ctx* ctx, const std::vector<CommMsg>& msgs, cudaStream_t st);
size_t spline_wire_doubles(int cap);
void spline_pack(const mb_spline* s, int cap, double* w);
mb_spline* spline_unpack(mb_ctx* ctx, const double* w, int cap);
// root sends its spline, the other ranks get a new handle (NULL on the root); on ctx->stream, synchronises it
mb_spline* spline_bcast(mb_ctx* ctx, const mb_spline* s_root, int cap, int root);

// tiles.cu - mltps part 3/4 (V73:649-895), machisplin.tiles.merge (V73:1392-1548), gram, gather
void tiles_tps(mb_ctx* ctx, const mb_grid& g, const double* knots_xy, const double* resid, int n, int tile_px,
               double fit_halo, double keep_halo, int min_pts, double lambda, int method, double* out_dev,
               cudaStream_t st);
void tiles_merge(mb_ctx* ctx, const mb_grid& g, int nC, int nR, const mb_window* wins,
                 const double* const* tiles_dev, double* out_dev, cudaStream_t st);
// machisplin.tiles.merge with the tiles spread over the ranks of the context's communicator (tile t on rank t % size): seam
// strips travel point to point, every rank blends the cells it owns (tiles_owned_window) - no gather
mb_window tiles_owned_window(const mb_grid& g, int nC, int nR, const mb_window* wins, int t);
void tiles_merge_shard(mb_ctx* ctx, const mb_grid& g, int nC, int nR, const mb_window* wins, const double* const* my_tiles_dev,
                       double* const* out_dev, cudaStream_t st);
void gram(mb_ctx* ctx, const double* R_dev, int n, int K, double* G_dev, cudaStream_t st);
void gather_cells(mb_ctx* ctx, const double* raster_dev, int64_t row_stride, int nrow, int ncol, const int32_t* row,
                  const int32_t* col, int n, double* out_host, cudaStream_t st);

}  // namespace mb
