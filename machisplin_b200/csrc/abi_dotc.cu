// .C()-loadable twins of the stateless entry points: every argument is a pointer to an R vector, nothing is
// returned, status[0] receives 0 or a negative MB_E_* code.  An R user needs only dyn.load("libmachisplin_b200.so") - no
// compiled glue, no R headers.  Handles cannot cross .C(), so these calls work on one process-wide context (device
// MB_DEVICE, default 0, created on first use) and cover the call sites whose inputs and outputs are plain numeric vectors:
// fields::Tps + terra::interpolate (incl. mltps' internal tiling), machisplin.tiles.merge, and the Gram of the RSS objective.
// The per-cell ensemble takes the fitted-model descriptors (a struct of ~30 arrays) and device-resident handles: that call site
// goes through the .Call shim (r/mb_shim.c).
#include "common.cuh"
#include "internal.h"

#include <cstdlib>
#include <mutex>

namespace {
std::mutex g_mu;
mb_ctx* g_ctx = nullptr;

mb_ctx* global_ctx(int* status) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_ctx) {
    const char* d = std::getenv("MB_DEVICE");
    const int rc = mb_init(d ? std::atoi(d) : 0, &g_ctx);
    if (rc != MB_OK) { *status = rc; g_ctx = nullptr; }
  }
  return g_ctx;
}
mb_grid grid_of(const double* g6) {
  return mb_grid{g6[0], g6[1], g6[2], g6[3], (int32_t)g6[4], (int32_t)g6[5]};
}
}  // namespace

extern "C" {

void mbC_gram(const double* R, const int* n, const int* K, double* G, int* status) {
  *status = MB_OK;
  mb_ctx* ctx = global_ctx(status);
  if (ctx) *status = mb_gram(ctx, R, *n, *K, G);
}

void mbC_tps_surface(const double* xy, const double* y, const int* n, const double* grid6, const double* lambda,
                     const int* tile_px, double* out, double* lambda_out, int* status) {
  *status = MB_OK;
  mb_ctx* ctx = global_ctx(status);
  if (!ctx) return;
  const mb_grid g = grid_of(grid6);
  mb_spline* sp = nullptr;
  *status = mb_mltps_predict(ctx, &g, nullptr, nullptr, 0, xy, y, *n, *lambda, *tile_px, out, &sp);
  if (lambda_out) {
    *lambda_out = std::nan("");                 // tiled mode: one lambda per sub-tile, none to report
    if (sp) mb_spline_get(sp, nullptr, nullptr, nullptr, nullptr, nullptr, lambda_out, nullptr, nullptr);
  }
  mb_spline_free(sp);
}

void mbC_tiles_merge(const double* grid6, const int* nC, const int* nR, const int* wins, const double* tiles_flat, double* out,
                     int* status) {
  *status = MB_OK;
  mb_ctx* ctx = global_ctx(status);
  if (!ctx) return;
  const mb_grid g = grid_of(grid6);
  const int nt = *nC * *nR;
  if (nt <= 0 || nt > 4096) { mb::set_last_error("bad tile count"); *status = MB_E_ARG; return; }
  std::vector<mb_window> w(nt);
  std::vector<const double*> ptrs(nt);
  size_t off = 0;
  for (int t = 0; t < nt; ++t) {
    w[t] = mb_window{wins[4 * t], wins[4 * t + 1], wins[4 * t + 2], wins[4 * t + 3]};
    if (w[t].r1 <= w[t].r0 || w[t].c1 <= w[t].c0) { mb::set_last_error("empty tile window"); *status = MB_E_ARG; return; }
    ptrs[t] = tiles_flat + off;
    off += (size_t)(w[t].r1 - w[t].r0) * (w[t].c1 - w[t].c0);
  }
  *status = mb_tiles_merge(ctx, &g, *nC, *nR, w.data(), ptrs.data(), out);
}

void mbC_last_error(char** buf) {
  // .C() hands a character vector over as char**; the message is copied into buf[0] up to the length R allocated for it
  if (!buf || !buf[0]) return;
  const size_t cap = std::strlen(buf[0]);
  const char* msg = mb_last_error();
  const size_t n = std::min(cap, std::strlen(msg));
  std::memcpy(buf[0], msg, n);
  for (size_t i = n; i < cap; ++i) buf[0][i] = ' ';
}

void mbC_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ctx) { mb_shutdown(g_ctx); g_ctx = nullptr; }
}

}  // extern "C"
