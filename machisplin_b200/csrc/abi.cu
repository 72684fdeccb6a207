// extern "C" entry points of libmachisplin_b200 (include/machisplin_b200.h).
// Each one is a thin guarded wrapper: argument checks, H2D/D2H for the host variants, then the
// device implementation in tps_eval.cu / tps_fit.cu / ensemble.cu / tiles.cu.
#include "common.cuh"
#include "internal.h"

#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>

namespace mb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
}  // namespace mb

using namespace mb;

// ---- cache of freed device blocks (DevBuf, common.cuh) ------------------------------------------------------------------
namespace mb {
namespace {
struct DevCache {
  std::mutex mu;
  std::map<void*, std::pair<int, size_t>> live;                       // block -> (device, capacity)
  std::map<std::pair<int, size_t>, std::vector<void*>> spare;         // (device, capacity) -> blocks
  size_t spare_bytes = 0;
  const bool on = [] { const char* v = std::getenv("MB_DEV_CACHE"); return !(v && v[0] == '0'); }();
};
DevCache& dev_cache() { static DevCache* c = new DevCache; return *c; }   // never destroyed: handles may outlive main()
constexpr size_t kSpareLimit = size_t(2) << 30;
size_t size_class(size_t bytes) {
  size_t c = 512;
  while (c < bytes && c < (size_t(1) << 20)) c <<= 1;                  // powers of two up to 1 MiB
  if (c >= bytes) return c;
  return (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);   // then whole MiB
}
}  // namespace

void* dev_cache_take(size_t bytes) {
  DevCache& c = dev_cache();
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t cap = c.on ? size_class(bytes) : bytes;
  if (c.on) {
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.spare.find({dev, cap});
    if (it != c.spare.end() && !it->second.empty()) {
      void* p = it->second.back();
      it->second.pop_back();
      c.spare_bytes -= cap;
      c.live[p] = {dev, cap};
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, cap);
  if (e != cudaSuccess && c.on) {                                      // give the cached blocks back and retry once
    (void)cudaGetLastError();
    dev_cache_flush();
    e = cudaMalloc(&p, cap);
  }
  if (e != cudaSuccess) throw Error(MB_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  if (c.on) {
    std::lock_guard<std::mutex> lk(c.mu);
    c.live[p] = {dev, cap};
  }
  return p;
}

void dev_cache_give(void* p) noexcept {
  if (!p) return;
  DevCache& c = dev_cache();
  if (!c.on) { cudaFree(p); return; }
  std::pair<int, size_t> key{-1, 0};
  {
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    if (it != c.live.end()) { key = it->second; c.live.erase(it); }
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (key.first != dev || key.second == 0) { cudaFree(p); return; }    // unknown block or another device: plain free
  cudaDeviceSynchronize();                                             // nothing in flight may still touch the block
  std::lock_guard<std::mutex> lk(c.mu);
  if (c.spare_bytes + key.second > kSpareLimit) { cudaFree(p); return; }
  c.spare[key].push_back(p);
  c.spare_bytes += key.second;
}

void dev_cache_flush() noexcept {
  DevCache& c = dev_cache();
  int dev = 0;
  cudaGetDevice(&dev);
  std::vector<void*> blocks;
  {
    std::lock_guard<std::mutex> lk(c.mu);
    for (auto& kv : c.spare)
      if (kv.first.first == dev) {
        for (void* p : kv.second) { blocks.push_back(p); c.spare_bytes -= kv.first.second; }
        kv.second.clear();
      }
  }
  for (void* p : blocks) cudaFree(p);
}
}  // namespace mb

extern "C" {

int mb_version(void) { return MB_VERSION; }

const char* mb_last_error(void) { return g_last_error.c_str(); }

int mb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int mb_init(int device, mb_ctx** out) {
  return guarded([&] {
    MB_REQUIRE(out != nullptr, "ctx out pointer is NULL");
    *out = nullptr;
    int n = 0;
    MB_CUDA(cudaGetDeviceCount(&n));
    MB_REQUIRE(device >= 0 && device < n, "no such CUDA device");
    MB_CUDA(cudaSetDevice(device));
    auto ctx = std::make_unique<mb_ctx>();
    ctx->device = device;
    cudaDeviceProp prop;
    MB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    // the fit is the critical path of mb_mltps_predict*: its stream outranks the one the bulk per-cell
    // kernels run on, so its many small kernels are scheduled as soon as an SM frees up
    int prio_lo = 0, prio_hi = 0;
    MB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    MB_CUDA(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi));
    MB_CUDA(cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_lo));
    MB_CUDA(cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking));
    MB_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    MB_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    MB_CUDA(cudaEventCreateWithFlags(&ctx->ev_stage1, cudaEventDisableTiming));
    init_logtab(ctx.get());
    *out = ctx.release();
  });
}

void mb_shutdown(mb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  fit_release(ctx);
  for (const mb_timed_launch& t : ctx->timed) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  cudaDeviceSynchronize();
  ctx->arena.release();
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_stage1) cudaEventDestroy(ctx->ev_stage1);
  for (cudaEvent_t ev : ctx->sbr_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->sbr_aux) cudaStreamDestroy(ctx->sbr_aux);
  greenctx_release(ctx);
  for (cudaEvent_t ev : ctx->ev_ens) if (ev) cudaEventDestroy(ev);
  if (ctx->ens_aux) cudaStreamDestroy(ctx->ens_aux);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->copy) cudaStreamDestroy(ctx->copy);
  for (auto& w : ctx->lanes) {
    w->logtab.p = nullptr; w->logtab.n = 0;
    w->arena.release();
    for (cudaEvent_t e : w->event_pool) cudaEventDestroy(e);
    for (cudaEvent_t ev : w->sbr_ev) if (ev) cudaEventDestroy(ev);
    if (w->sbr_aux) cudaStreamDestroy(w->sbr_aux);
    if (w->stream) cudaStreamDestroy(w->stream);
  }
  ctx->lanes.clear();
  for (cudaEvent_t ev : ctx->ev_blocks) cudaEventDestroy(ev);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  mb::dev_cache_flush();
}

int mb_sync(mb_ctx* ctx) {
  return guarded([&] {
    MB_REQUIRE(ctx, "ctx is NULL");
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int64_t mb_launch_count(const mb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mb_timing_enable(mb_ctx* ctx, int on) {
  return guarded([&] {
    MB_REQUIRE(ctx, "ctx is NULL");
    ctx->timing = on != 0;
  });
}

int mb_timing_collect(mb_ctx* ctx, int cap, const char** names, double* total_ms, int64_t* launches) {
  int n = 0;
  const int rc = guarded([&] {
    MB_REQUIRE(ctx && names && total_ms && launches && cap > 0, "bad argument");
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaDeviceSynchronize());
    for (const mb_timed_launch& t : ctx->timed) {
      float ms = 0.f;
      MB_CUDA(cudaEventElapsedTime(&ms, t.start, t.stop));
      int k = 0;
      for (; k < n; ++k)
        if (std::strcmp(names[k], t.name) == 0) break;
      if (k == n) {
        if (n == cap) continue;
        names[n] = t.name; total_ms[n] = 0; launches[n] = 0; ++n;
      }
      total_ms[k] += ms;
      launches[k] += 1;
      ctx->event_pool.push_back(t.start);
      ctx->event_pool.push_back(t.stop);
    }
    ctx->timed.clear();
  });
  return rc == MB_OK ? n : rc;
}

int mb_set_fast_eval_params(mb_ctx* ctx, int cheb_p, int leaf_cols, int leaf_rows) {
  return guarded([&] {
    MB_REQUIRE(ctx, "ctx is NULL");
    MB_REQUIRE(cheb_p == 0 || cheb_p == 8 || cheb_p == 10 || cheb_p == 12 || cheb_p == 14 || cheb_p == 16,
               "cheb_p must be 0 (auto), 8, 10, 12, 14 or 16");
    MB_REQUIRE(leaf_cols == 0 || leaf_cols == 32, "leaf_cols must be 0 (auto) or 32");
    MB_REQUIRE(leaf_rows >= 0 && leaf_rows <= 128, "leaf_rows must be in [0, 128]");
    ctx->cheb_p = cheb_p;
    ctx->leaf_cols = leaf_cols;
    ctx->leaf_rows = leaf_rows;
  });
}

int mb_debug_values(mb_ctx* ctx, const char* name, double* out, int cap) {
  int n = 0;
  const int rc = guarded([&] {
    MB_REQUIRE(ctx && name && out, "NULL argument");
    if (std::string(name) == "sytrd_phase_ms") {
      for (; n < 8 && n < cap; ++n) out[n] = ctx->sytrd_prof_ms[n];
    } else if (std::string(name) == "sbr_band") {
      // 64 x m band storage after stage 1 of the two-stage tridiagonalisation (B[off + 64 j] = A[j + off, j])
      for (; n < (int)ctx->dbg_band.size() && n < cap; ++n) out[n] = ctx->dbg_band[n];
    } else {
      throw Error(MB_E_ARG, std::string("unknown debug vector '") + name + "'");
    }
  });
  return rc < 0 ? rc : n;
}

int mb_set_param(mb_ctx* ctx, const char* name, int value) {
  return guarded([&] {
    MB_REQUIRE(ctx && name, "NULL argument");
    const std::string n(name);
    if (n == "ens_overlap") {
      MB_REQUIRE(value >= 0 && value <= 2, "ens_overlap must be 0 (default = 2), 1 (forest and ksvm kernels side by side) or 2 (one after the other)");
      ctx->ens_overlap = value;
    } else if (n == "gc_split") {
      MB_REQUIRE(value >= -1 && value <= 1024, "gc_split must be -1 (no SM partitions), 0 (default) or the SM count of the fit partition");
      ctx->gc_split = value;
    } else if (n == "gc_share") {
      MB_REQUIRE(value >= 0 && value <= 100, "gc_share must be 0 (default) or a percentage of the raster's rows");
      ctx->gc_share = value;
    } else if (n == "leaf_impl") {
      MB_REQUIRE(value >= 0 && value <= 2, "leaf_impl must be 0 (default = 1), 1 (one warp per box) or 2 (one CTA per box)");
      ctx->leaf_impl = value;
    } else if (n == "leaf_tma") {
      MB_REQUIRE(value >= 0 && value <= 2, "leaf_tma must be 0 (default = 1), 1 (2-D tensor copy of the accumulator tile) or 2 (row copies)");
      ctx->leaf_tma = value;
    } else if (n == "ens_order") {
      MB_REQUIRE(value >= 0 && value <= 2, "ens_order must be 0 (default = 1), 1 (forest kernel first) or 2 (ksvm kernel first)");
      ctx->ens_order = value;
    } else if (n == "svm_ctas_per_sm") {
      MB_REQUIRE(value >= 0 && value <= 4, "svm_ctas_per_sm must be in [0, 4]");
      ctx->svm_ctas_per_sm = value;
    } else if (n == "ens_tma") {
      MB_REQUIRE(value >= 0 && value <= 2, "ens_tma must be 0 (default = 1), 1 (TMA tensor copies) or 2 (plain loads)");
      ctx->ens_tma = value;
    } else if (n == "tree_levels") {
      MB_REQUIRE(value >= 0 && value <= 2, "tree_levels must be 0 (default = 2), 1 or 2");
      ctx->tree_levels = value;
    } else if (n == "eigen_impl") {
      MB_REQUIRE(value == 0 || value == 1, "eigen_impl must be 0 (in-house) or 1 (cuSOLVER, validation)");
      ctx->eigen_impl = value;
    } else if (n == "sytrd_mode") {
      MB_REQUIRE(value >= 0 && value <= 3,
                 "sytrd_mode must be 0 (default = two-stage), 1 (one-stage persistent kernel), 2 (one-stage, kernel per phase) or 3 (two-stage)");
      ctx->sytrd_mode = value;
    } else if (n == "sbr_chase_ctas") {
      MB_REQUIRE(value >= 0 && value <= 4096, "sbr_chase_ctas must be in [0, 4096]");
      ctx->sbr_chase_ctas = value;
    } else if (n == "sbr_chase_sleep") {
      MB_REQUIRE(value >= -1 && value <= 100000, "sbr_chase_sleep must be in [-1, 100000] nanoseconds");
      ctx->sbr_chase_sleep = value;
    } else if (n == "sbr_chase_impl") {
      MB_REQUIRE(value >= 0 && value <= 4, "sbr_chase_impl must be 0 (automatic), 1 (watcher / publisher warps), 2 (three warps per sweep), 3 (tagged elements, no flags) or 4 (3 with capped registers)");
      ctx->sbr_chase_impl = value;
    } else if (n == "coef_impl") {
      MB_REQUIRE(value >= 0 && value <= 2, "coef_impl must be 0 (band form when well conditioned), 1 (band form whenever it exists) or 2 (dense Cholesky)");
      ctx->coef_impl = value;
    } else if (n == "svm_impl") {
      MB_REQUIRE(value >= 0 && value <= 3, "svm_impl must be 0 (default: 3 when P <= 8, else 2), 1 (tensor-core dot products, 3 x TF32), 2 (packed FP32) or 3 (tensor-core dot products, FP16 split operands)");
      ctx->svm_impl = value;
    } else if (n == "defer_ensemble") {
      MB_REQUIRE(value == 0 || value == 1, "defer_ensemble must be 0 or 1");
      ctx->defer_ensemble = value;
    } else if (n == "sbr_qr_impl") {
      MB_REQUIRE(value >= 0 && value <= 2, "sbr_qr_impl must be 0 (automatic), 1 (panel rows in shared memory) or 2 (in registers)");
      ctx->sbr_qr_impl = value;
    } else if (n == "sbr_qr_grid") {
      MB_REQUIRE(value == 0 || value == 1, "sbr_qr_grid must be 0 or 1");
      ctx->sbr_qr_grid = value;
    } else if (n == "sbr_debug") {
      MB_REQUIRE(value == 0 || value == 1, "sbr_debug must be 0 or 1");
      ctx->sbr_debug = value;
    } else if (n == "sytrd_ctas_per_sm") {
      MB_REQUIRE(value >= 0 && value <= 4, "sytrd_ctas_per_sm must be in [0, 4]");
      ctx->sytrd_ctas_per_sm = value;
    } else if (n == "eval_precision") {
      MB_REQUIRE(value >= 0 && value <= 2, "eval_precision must be 0 (auto), 1 (float64) or 2 (mixed)");
      ctx->eval_precision = value;
    } else {
      throw Error(MB_E_ARG, "unknown parameter '" + n + "'");
    }
  });
}

// ---- device memory helpers -----------------------------------------------------------------
int mb_dev_alloc(mb_ctx* ctx, size_t bytes, void** out) {
  return guarded([&] {
    MB_REQUIRE(ctx && out, "NULL argument");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) throw Error(MB_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  });
}
int mb_dev_free(mb_ctx* ctx, void* p) {
  return guarded([&] {
    MB_REQUIRE(ctx, "ctx is NULL");
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaFree(p));
  });
}
int mb_h2d(mb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  return guarded([&] {
    MB_REQUIRE(ctx && dst && src, "NULL argument");
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}
int mb_d2h(mb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  return guarded([&] {
    MB_REQUIRE(ctx && dst && src, "NULL argument");
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- splines ---------------------------------------------------------------------------------
int mb_spline_create(mb_ctx* ctx, const double* knots_xy, int np, const double* c, const double* d,
                     const double* center, const double* scale, mb_spline** out) {
  return guarded([&] {
    MB_REQUIRE(ctx && knots_xy && c && d && center && scale && out, "NULL argument");
    MB_REQUIRE(np >= 1, "spline needs at least one knot");
    MB_REQUIRE(scale[0] > 0 && scale[1] > 0, "x.scale must be positive");
    MB_CUDA(cudaSetDevice(ctx->device));
    auto s = std::make_unique<mb_spline>();
    s->np = np;
    s->kx.assign(knots_xy, knots_xy + np);
    s->ky.assign(knots_xy + np, knots_xy + 2 * np);
    s->sx.resize(np);
    s->sy.resize(np);
    for (int i = 0; i < np; ++i) {
      s->sx[i] = (s->kx[i] - center[0]) / scale[0];
      s->sy[i] = (s->ky[i] - center[1]) / scale[1];
    }
    s->c.assign(c, c + np);
    for (int i = 0; i < 3; ++i) s->d[i] = d[i];
    for (int i = 0; i < 2; ++i) { s->center[i] = center[i]; s->scale[i] = scale[i]; }
    ctx->arena.begin(ctx->stream);
    spline_finalize(ctx, s.get());
    *out = s.release();
  });
}

int mb_spline_np(const mb_spline* s) { return s ? s->np : MB_E_ARG; }

int mb_spline_get(const mb_spline* s, double* c, double* d, double* center, double* scale, double* knots_xy,
                  double* lambda, double* eff_df, double* gcv) {
  return guarded([&] {
    MB_REQUIRE(s, "spline is NULL");
    if (c) std::copy(s->c.begin(), s->c.end(), c);
    if (d) std::copy(s->d, s->d + 3, d);
    if (center) std::copy(s->center, s->center + 2, center);
    if (scale) std::copy(s->scale, s->scale + 2, scale);
    if (knots_xy) {
      std::copy(s->kx.begin(), s->kx.end(), knots_xy);
      std::copy(s->ky.begin(), s->ky.end(), knots_xy + s->np);
    }
    if (lambda) *lambda = s->lambda;
    if (eff_df) *eff_df = s->eff_df;
    if (gcv) *gcv = s->gcv;
  });
}

int mb_spline_get_decomp(const mb_spline* s, double* eta, double* tri_diag, double* tri_off, double* zhat) {
  return guarded([&] {
    MB_REQUIRE(s, "spline is NULL");
    MB_REQUIRE(!s->eta.empty(), "spline carries no decomposition (fixed-lambda Cholesky fit or created from coefficients)");
    if (eta) std::copy(s->eta.begin(), s->eta.end(), eta);
    const int m = s->np - 3;
    const double nan = std::nan("");
    if (s->tri_diag.empty()) {   // fitted through the eigenvector (validation) path: no tridiagonal form
      if (tri_diag) std::fill(tri_diag, tri_diag + m, nan);
      if (tri_off) std::fill(tri_off, tri_off + m - 1, nan);
      if (zhat) std::fill(zhat, zhat + m, nan);
      return;
    }
    if (tri_diag) std::copy(s->tri_diag.begin(), s->tri_diag.end(), tri_diag);
    if (tri_off) std::copy(s->tri_off.begin(), s->tri_off.end(), tri_off);
    if (zhat) std::copy(s->zhat.begin(), s->zhat.end(), zhat);
  });
}

void mb_spline_free(mb_spline* s) {
  if (!s) return;
  cudaSetDevice(s->device);   // the context may already be gone (R finalizer order, Python GC): never touch s->ctx here
  delete s;
}

int mb_tps_fit(mb_ctx* ctx, const double* xy, const double* y, int n, int L, double lambda, mb_spline** splines) {
  return guarded([&] {
    MB_REQUIRE(ctx && xy && y && splines, "NULL argument");
    MB_REQUIRE(n > 3 && L >= 1, "need n > 3 observations and L >= 1 responses");
    MB_CUDA(cudaSetDevice(ctx->device));
    ctx->arena.begin(ctx->stream);
    tps_fit(ctx, xy, y, n, L, lambda, splines);
  });
}

// ---- evaluation --------------------------------------------------------------------------------
static void eval_dispatch(mb_ctx* ctx, const mb_spline* s, const mb_grid* g, const mb_window* w, int method,
                          double* out_dev, int64_t stride, cudaStream_t st) {
  if (method == MB_EVAL_DIRECT) tps_eval_direct(ctx, s, *g, *w, out_dev, stride, st);
  else if (method == MB_EVAL_FAST) tps_eval_fast(ctx, s, *g, *w, out_dev, stride, st);
  else throw Error(MB_E_ARG, "unknown evaluation method");
}

int mb_tps_eval_dev(mb_ctx* ctx, const mb_spline* s, const mb_grid* g, const mb_window* w, int method,
                    double* out_dev, int64_t out_row_stride, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && s && out_dev, "NULL argument");
    check_grid(g);
    check_window(g, w);
    MB_REQUIRE(out_row_stride >= w->c1 - w->c0, "row stride shorter than the window");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    eval_dispatch(ctx, s, g, w, method, out_dev, out_row_stride, st);
  });
}

int mb_tps_eval(mb_ctx* ctx, const mb_spline* s, const mb_grid* g, const mb_window* w, int method,
                double* out_host) {
  return guarded([&] {
    MB_REQUIRE(ctx && s && out_host, "NULL argument");
    check_grid(g);
    check_window(g, w);
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)(w->r1 - w->r0) * (w->c1 - w->c0);
    ctx->arena.begin(ctx->stream);
    ABuf<double> d(ctx->arena, n);
    eval_dispatch(ctx, s, g, w, method, d.p, w->c1 - w->c0, ctx->stream);
    MB_CUDA(cudaMemcpyAsync(out_host, d.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int mb_tps_predict_points(mb_ctx* ctx, const mb_spline* s, const double* xy, int n, double* out_host) {
  return guarded([&] {
    MB_REQUIRE(ctx && s && xy && out_host, "NULL argument");
    MB_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return;
    MB_CUDA(cudaSetDevice(ctx->device));
    ctx->arena.begin(ctx->stream);
    ABuf<double> dx(ctx->arena, n), dy(ctx->arena, n), df(ctx->arena, n);
    dx.upload(xy, n, ctx->stream);
    dy.upload(xy + n, n, ctx->stream);
    tps_predict_points_dev(ctx, s, dx.p, dy.p, n, df.p, ctx->stream);
    MB_CUDA(cudaMemcpyAsync(out_host, df.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- ensemble ------------------------------------------------------------------------------------
int mb_ensemble_create(mb_ctx* ctx, const mb_grid* g, const mb_models* m, const char* kept, const double* w,
                       double w_total, mb_ensemble** out) {
  return guarded([&] {
    MB_REQUIRE(ctx && m && kept && w && out, "NULL argument");
    check_grid(g);
    MB_CUDA(cudaSetDevice(ctx->device));
    *out = ensemble_create(ctx, *g, *m, kept, w, w_total);
  });
}

void mb_ensemble_free(mb_ensemble* e) { ensemble_free(e); }

int mb_ensemble_eval_dev(mb_ctx* ctx, const mb_ensemble* e, const float* cov_dev, int C, const mb_spline* spline,
                         const double* tps_surface_dev, const mb_window* w, double* out_dev, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && e && out_dev, "NULL argument");
    MB_REQUIRE(C == 0 || cov_dev, "covariate planes are NULL");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    ensemble_eval(ctx, e, cov_dev, C, spline, tps_surface_dev, w, out_dev, st);
  });
}

int mb_ensemble_eval(mb_ctx* ctx, const mb_ensemble* e, const float* cov_host, int C, const mb_spline* spline,
                     const double* tps_surface_host, const mb_window* w, double* out_host) {
  return guarded([&] {
    MB_REQUIRE(ctx && e && out_host && w, "NULL argument");
    MB_REQUIRE(C == 0 || cov_host, "covariate planes are NULL");
    MB_CUDA(cudaSetDevice(ctx->device));
    const mb_grid g = ensemble_grid(e);
    check_window(&g, w);
    const size_t plane = (size_t)g.nrow * g.ncol;
    const size_t nwin = (size_t)(w->r1 - w->r0) * (w->c1 - w->c0);
    // only the rows of the window travel: plane p rows [r0, r1) -> device plane of the same geometry
    ctx->arena.begin(ctx->stream);
    ABuf<float> d_cov(ctx->arena, (size_t)C * plane);
    for (int p = 0; p < C; ++p) {
      const size_t off = p * plane + (size_t)w->r0 * g.ncol;
      MB_CUDA(cudaMemcpyAsync(d_cov.p + off, cov_host + off, sizeof(float) * (size_t)(w->r1 - w->r0) * g.ncol,
                              cudaMemcpyHostToDevice, ctx->stream));
    }
    ABuf<double> d_tps(ctx->arena), d_out(ctx->arena, nwin);
    if (tps_surface_host) d_tps.upload(tps_surface_host, nwin, ctx->stream);
    ensemble_eval(ctx, e, d_cov.p, C, spline, tps_surface_host ? d_tps.p : nullptr, w, d_out.p, ctx->stream);
    MB_CUDA(cudaMemcpyAsync(out_host, d_out.p, nwin * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int mb_ensemble_predict_points(mb_ctx* ctx, const mb_ensemble* e, const double* X, int n, double* out_host) {
  return guarded([&] {
    MB_REQUIRE(ctx && e && X && out_host, "NULL argument");
    MB_CUDA(cudaSetDevice(ctx->device));
    ctx->arena.begin(ctx->stream);
    ensemble_predict_points(ctx, e, X, n, out_host);
  });
}

// ---- tiles ---------------------------------------------------------------------------------------
int mb_tiles_tps_dev(mb_ctx* ctx, const mb_grid* g, const double* knots_xy, const double* resid, int n,
                     int tile_px, double fit_halo, double keep_halo, int min_pts, double lambda, int method,
                     double* out_dev, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && knots_xy && resid && out_dev, "NULL argument");
    check_grid(g);
    MB_REQUIRE(tile_px > 0 && fit_halo >= 0 && keep_halo >= 0 && keep_halo <= fit_halo, "bad tiling parameters");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    tiles_tps(ctx, *g, knots_xy, resid, n, tile_px, fit_halo, keep_halo, min_pts, lambda, method, out_dev, st);
  });
}

int mb_tiles_tps(mb_ctx* ctx, const mb_grid* g, const double* knots_xy, const double* resid, int n, int tile_px,
                 double fit_halo, double keep_halo, int min_pts, double lambda, int method, double* out_host) {
  return guarded([&] {
    MB_REQUIRE(ctx && knots_xy && resid && out_host, "NULL argument");
    check_grid(g);
    MB_REQUIRE(tile_px > 0 && fit_halo >= 0 && keep_halo >= 0 && keep_halo <= fit_halo, "bad tiling parameters");
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t ncell = (size_t)g->nrow * g->ncol;
    ctx->arena.begin(ctx->stream);
    ABuf<double> d_out(ctx->arena, ncell);
    tiles_tps(ctx, *g, knots_xy, resid, n, tile_px, fit_halo, keep_halo, min_pts, lambda, method, d_out.p,
              ctx->stream);
    MB_CUDA(cudaMemcpyAsync(out_host, d_out.p, ncell * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int mb_tiles_merge_dev(mb_ctx* ctx, const mb_grid* g, int nC, int nR, const mb_window* wins,
                       const double* const* tiles_dev, double* out_dev, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && wins && tiles_dev && out_dev, "NULL argument");
    check_grid(g);
    MB_REQUIRE(nC >= 1 && nR >= 1, "need at least one tile");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    tiles_merge(ctx, *g, nC, nR, wins, tiles_dev, out_dev, st);
  });
}

int mb_tiles_owned_window(const mb_grid* g, int nC, int nR, const mb_window* wins, int t, mb_window* own) {
  return guarded([&] {
    MB_REQUIRE(wins && own, "NULL argument");
    check_grid(g);
    MB_REQUIRE(nC >= 1 && nR >= 1, "need at least one tile");
    for (int u = 0; u < nC * nR; ++u) check_window(g, &wins[u]);
    *own = tiles_owned_window(*g, nC, nR, wins, t);
  });
}

int mb_tiles_merge_shard_dev(mb_ctx* ctx, const mb_grid* g, int nC, int nR, const mb_window* wins,
                             const double* const* my_tiles_dev, double* const* out_dev, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && wins && my_tiles_dev && out_dev, "NULL argument");
    check_grid(g);
    MB_REQUIRE(nC >= 1 && nR >= 1, "need at least one tile");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    tiles_merge_shard(ctx, *g, nC, nR, wins, my_tiles_dev, out_dev, st);
  });
}

int mb_tiles_merge(mb_ctx* ctx, const mb_grid* g, int nC, int nR, const mb_window* wins,
                   const double* const* tiles_host, double* out_host) {
  return guarded([&] {
    MB_REQUIRE(ctx && wins && tiles_host && out_host, "NULL argument");
    check_grid(g);
    MB_REQUIRE(nC >= 1 && nR >= 1, "need at least one tile");
    MB_CUDA(cudaSetDevice(ctx->device));
    const int nt = nC * nR;
    ctx->arena.begin(ctx->stream);
    std::vector<const double*> ptrs(nt);
    for (int t = 0; t < nt; ++t) {
      check_window(g, &wins[t]);
      const size_t n = (size_t)(wins[t].r1 - wins[t].r0) * (wins[t].c1 - wins[t].c0);
      ptrs[t] = ctx->arena.upload(tiles_host[t], n, ctx->stream);
    }
    const size_t ncell = (size_t)g->nrow * g->ncol;
    ABuf<double> d_out(ctx->arena, ncell);
    tiles_merge(ctx, *g, nC, nR, wins, ptrs.data(), d_out.p, ctx->stream);
    MB_CUDA(cudaMemcpyAsync(out_host, d_out.p, ncell * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- gram / gather -----------------------------------------------------------------------------
int mb_gram_dev(mb_ctx* ctx, const double* R_dev, int n, int K, double* G_dev, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && R_dev && G_dev, "NULL argument");
    MB_REQUIRE(n >= 1 && K >= 1 && K <= 8, "need n >= 1 and 1 <= K <= 8");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    gram(ctx, R_dev, n, K, G_dev, st);
  });
}

int mb_gram(mb_ctx* ctx, const double* R_host, int n, int K, double* G_host) {
  return guarded([&] {
    MB_REQUIRE(ctx && R_host && G_host, "NULL argument");
    MB_REQUIRE(n >= 1 && K >= 1 && K <= 8, "need n >= 1 and 1 <= K <= 8");
    MB_CUDA(cudaSetDevice(ctx->device));
    ctx->arena.begin(ctx->stream);
    ABuf<double> dR(ctx->arena, (size_t)n * K), dG(ctx->arena, (size_t)K * K);
    dR.upload(R_host, (size_t)n * K, ctx->stream);
    gram(ctx, dR.p, n, K, dG.p, ctx->stream);
    MB_CUDA(cudaMemcpyAsync(G_host, dG.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int mb_gather_cells_dev(mb_ctx* ctx, const double* raster_dev, int64_t row_stride, int nrow, int ncol, const int32_t* row,
                        const int32_t* col, int n, double* out_host, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && raster_dev && row && col && out_host, "NULL argument");
    MB_REQUIRE(nrow > 0 && ncol > 0 && row_stride >= ncol, "raster shape / row stride");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    gather_cells(ctx, raster_dev, row_stride, nrow, ncol, row, col, n, out_host, st);
  });
}

// ---- mltps parts 2-5 in one call -----------------------------------------------------------------
// cov_host != NULL (host-buffer entry point): `cov` is an empty device buffer; the planes are uploaded in row
// blocks on the copy stream and every block's ensemble kernels start as soon as its rows have landed, so the
// PCIe transfer (1.6 GB at config 3) hides behind part 2 instead of preceding it.
static void mltps_predict(mb_ctx* ctx, const mb_grid& g, const mb_ensemble* e, float* cov, int C,
                          const double* knots_xy, const double* resid, int n, double lambda, int tile_px,
                          double* out, mb_spline** spline_out, cudaStream_t st, const float* cov_host = nullptr,
                          int bcast_root = -1) {
  // bcast_root >= 0 (mb_mltps_predict_shard*): g is this rank's row block of a raster that is sharded over the communicator.
  // The global spline is fitted once, on the root, while every rank - the root included - runs the per-cell ensemble kernels of
  // its own block; its descriptor (24 bytes per knot) then travels by ncclBroadcast and each rank evaluates it on its own cells.
  const bool sharded = bcast_root >= 0 && ctx->comm != nullptr;
  const bool fit_here = !sharded || ctx->comm_rank == bcast_root;
  const mb_window full{0, g.nrow, 0, g.ncol};
  const size_t ncell = (size_t)g.nrow * g.ncol;
  if (spline_out) *spline_out = nullptr;
  // this call forks onto the context's streams: the next arena.begin() has to drain all of them before it reuses the blocks
  ctx->arena.also_used(ctx->stream);
  ctx->arena.also_used(ctx->side);
  ctx->arena.also_used(ctx->copy);
  ctx->arena.also_used(st);
  if (e) {
    const mb_grid eg = ensemble_grid(e);
    MB_REQUIRE(eg.nrow == g.nrow && eg.ncol == g.ncol && eg.xmin == g.xmin && eg.xmax == g.xmax &&
                   eg.ymin == g.ymin && eg.ymax == g.ymax, "grid differs from the grid the ensemble was created for");
    MB_REQUIRE(C == ensemble_ncov(e), "number of covariate planes does not match the model descriptors (P = C + 2)");
    MB_REQUIRE(C == 0 || cov, "covariate planes are NULL");
  }
  // part 2 (trees + svm) runs on the side stream; it only needs the covariates.  With the two-stage GCV fit the launch is
  // deferred to the point where stage 1 of the tridiagonalisation (full-GPU FP64 products, ~1/3 of the fit) has been
  // enqueued, and waits for it on the device: the rest of the fit - bulge chasing, bisection, Cholesky - is latency-bound
  // on a few SMs and shares the GPU with the per-cell kernels, whereas the panel kernels of stage 1 would be starved by
  // their 262 144-CTA grids.  In the host-buffer entry point the covariate planes travel meanwhile.
  double* acc = nullptr;
  const bool heavy = e != nullptr;
  const bool tps = sharded ? n > 0 : (knots_xy && resid && n > 0);
  MB_REQUIRE(!sharded || !tps || tile_px <= 0, "the sharded entry point evaluates one global spline (tile_px <= 0)");
  MB_REQUIRE(!(sharded && tps && fit_here) || (knots_xy && resid), "the broadcast root needs the knots and residuals");
  const bool one_spline = tps && (tile_px <= 0 || ((g.nrow + tile_px - 1) / tile_px) * ((g.ncol + tile_px - 1) / tile_px) == 1);
  const bool defer = heavy && one_spline && fit_here && lambda < 0 && ctx->eigen_impl != 1 &&
                     (ctx->sytrd_mode == 0 || ctx->sytrd_mode == 3) && ctx->defer_ensemble;
  int nblk_copy = 0, rows_per = 0;
  if (heavy) {
    acc = ctx->arena.take_n<double>((size_t)(acc_stride(full) * acc_rows(full)));
    MB_CUDA(cudaEventRecord(ctx->ev_fork, st));
    MB_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    if (cov_host && C > 0) {
      MB_CUDA(cudaStreamWaitEvent(ctx->copy, ctx->ev_fork, 0));
      nblk_copy = std::max(1, std::min(8, g.nrow / 256));
      rows_per = ((g.nrow + nblk_copy - 1) / nblk_copy + 63) / 64 * 64;
      int b = 0;
      for (int r0 = 0; r0 < g.nrow; r0 += rows_per, ++b) {
        const int r1 = std::min(g.nrow, r0 + rows_per);
        for (int p = 0; p < C; ++p) {
          const size_t off = (size_t)p * ncell + (size_t)r0 * g.ncol;
          MB_CUDA(cudaMemcpyAsync(cov + off, cov_host + off, sizeof(float) * (size_t)(r1 - r0) * g.ncol,
                                  cudaMemcpyHostToDevice, ctx->copy));
        }
        if ((int)ctx->ev_blocks.size() <= b) {
          cudaEvent_t ev;
          MB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
          ctx->ev_blocks.push_back(ev);
        }
        MB_CUDA(cudaEventRecord(ctx->ev_blocks[b], ctx->copy));
      }
    }
  }
  // SM partitions (greenctx.cu): with a forest kernel in the chain and a fit large enough to have a stage 1 worth the name, the
  // forest kernel starts NOW on the ensemble partition while stage 1 runs on the fit partition; what follows the forest kernel
  // (ksvm + smooth models) is deferred behind stage 1 as before, on all SMs.
  const bool use_gc = defer && ctx->gc_split >= 0 && n >= 1500 && ensemble_has_forest_kernel(e) &&
                      greenctx_setup(ctx, ctx->gc_split > 0 ? ctx->gc_split : 64);   // 64: forest kernel (84 SMs) and stage 1 end together (profiles/r3g_*)
  // rows [lo, hi) of the raster, in the row blocks the covariates travel in (host-buffer entry point) or in one piece
  auto accumulate = [&](cudaStream_t s, int part, int lo, int hi) {
    if (hi <= lo) return;
    if (nblk_copy > 0) {
      int b = 0;
      for (int r0 = 0; r0 < g.nrow; r0 += rows_per, ++b) {
        const int r1 = std::min(g.nrow, r0 + rows_per);
        const int a0 = std::max(r0, lo), a1 = std::min(r1, hi);
        if (a1 <= a0) continue;
        MB_CUDA(cudaStreamWaitEvent(s, ctx->ev_blocks[b], 0));
        const mb_window wb{a0, a1, 0, g.ncol};
        ensemble_accumulate(ctx, e, cov, C, wb, acc + (int64_t)a0 * acc_stride(full), s, part);
      }
    } else {
      const mb_window wb{lo, hi, 0, g.ncol};
      ensemble_accumulate(ctx, e, cov, C, wb, acc + (int64_t)lo * acc_stride(full), s, part);
    }
  };
  // the forest kernel's rows that run on the ensemble partition beside stage 1 (the rest follows on all SMs): sized so that the
  // partition is done when stage 1 is ("gc_share" percent of the rows; multiples of the 8-row forest tiles)
  int rows_gc = 0;
  if (use_gc) {
    const int share = ctx->gc_share > 0 ? ctx->gc_share : 100;    // measured flat between 78 and 100 (profiles/r2u_*)
    rows_gc = share >= 100 ? g.nrow : std::min(g.nrow, (int)((int64_t)g.nrow * share / 100) / 64 * 64);
    ctx->arena.also_used(ctx->gc_ens_stream);
    MB_CUDA(cudaStreamWaitEvent(ctx->gc_ens_stream, ctx->ev_fork, 0));
    accumulate(ctx->gc_ens_stream, 1, 0, rows_gc);
    MB_CUDA(cudaEventRecord(ctx->gc_ev[2], ctx->gc_ens_stream));
  }
  bool launched = false;
  auto launch_ensemble = [&](bool after_stage1) {
    if (launched || !heavy) return;
    launched = true;
    if (after_stage1) {
      MB_CUDA(cudaEventRecord(ctx->ev_stage1, ctx->stream));
      MB_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_stage1, 0));
    }
    if (use_gc) {
      accumulate(ctx->side, 1, rows_gc, g.nrow);                      // the forest kernel's remaining rows, all SMs
      MB_CUDA(cudaStreamWaitEvent(ctx->side, ctx->gc_ev[2], 0));
      accumulate(ctx->side, 2, 0, g.nrow);
    } else {
      accumulate(ctx->side, 0, 0, g.nrow);
    }
    MB_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
  };
  if (!defer) launch_ensemble(false);
  // part 3: fields::Tps of the residuals on the context stream, beside the kernels above
  std::unique_ptr<mb_spline> sp;
  const double* surface = nullptr;
  if (tps) {
    const int nRx = tile_px > 0 ? (g.nrow + tile_px - 1) / tile_px : 1;
    const int nCx = tile_px > 0 ? (g.ncol + tile_px - 1) / tile_px : 1;
    if (nRx * nCx == 1 || sharded) {                                 // V73:748-753
      mb_spline* raw = nullptr;
      if (fit_here) {
        if (defer) ctx->after_stage1 = [&] { launch_ensemble(true); };
        ctx->fit_shares_gpu = heavy;     // the bulge chase picks its small-footprint variant (sbr.cu)
        ctx->gc_stage1 = use_gc;         // stage 1 of the tridiagonalisation on the fit partition
        try {
          tps_fit(ctx, knots_xy, resid, n, 1, lambda, &raw);
        } catch (...) {
          ctx->after_stage1 = nullptr;
          ctx->fit_shares_gpu = false;
          ctx->gc_stage1 = false;
          throw;
        }
        ctx->after_stage1 = nullptr;
        ctx->fit_shares_gpu = false;
        ctx->gc_stage1 = false;
        sp.reset(raw);
        launch_ensemble(false);        // not reached through the hook (e.g. a fit small enough to skip stage 1 entirely)
        if (sharded) spline_bcast(ctx, sp.get(), n, bcast_root);
      } else {
        launch_ensemble(false);        // this rank's cells first; the receive below then waits for the root's fit
        sp.reset(spline_bcast(ctx, nullptr, n, bcast_root));
      }
    } else {                                                         // V73:649-895
      double* surf = ctx->arena.take_n<double>(ncell);
      tiles_tps(ctx, g, knots_xy, resid, n, tile_px, 0.2, 0.025, 10, lambda, MB_EVAL_FAST, surf, ctx->stream);
      MB_CUDA(cudaStreamSynchronize(ctx->stream));
      surface = surf;
    }
  }
  // parts 3-5: TPS surface + smooth models + combine, one pass over the grid.  With a spline the join of the ensemble stream is
  // deferred to the point where the fast evaluator launches its grid-evaluation kernel (its far-field set-up needs the spline only)
  if (heavy && e && sp) ctx->leaf_wait = ctx->ev_join;
  else if (heavy) MB_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
  if (e) {
    try {
      ensemble_finish(ctx, e, sp.get(), surface, full, acc, out, st);
    } catch (...) {
      ctx->leaf_wait = nullptr;
      throw;
    }
    if (ctx->leaf_wait) {            // not consumed (cannot happen with a spline; defensive): wait here
      MB_CUDA(cudaStreamWaitEvent(st, ctx->leaf_wait, 0));
      ctx->leaf_wait = nullptr;
    }
  } else if (sp) {
    tps_eval_fast(ctx, sp.get(), g, full, out, g.ncol, st);
  } else if (surface) {
    MB_CUDA(cudaMemcpyAsync(out, surface, ncell * sizeof(double), cudaMemcpyDeviceToDevice, st));
  } else {
    throw Error(MB_E_ARG, "nothing to predict: no ensemble and no TPS input");
  }
  if (spline_out) *spline_out = sp.release();
  else if (sp) { MB_CUDA(cudaStreamSynchronize(st)); }
}

int mb_mltps_predict_dev(mb_ctx* ctx, const mb_grid* g, const mb_ensemble* e, const float* cov_dev, int C,
                         const double* knots_xy, const double* resid, int n, double lambda, int tile_px,
                         double* out_dev, mb_spline** spline_out, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && out_dev, "NULL argument");
    check_grid(g);
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    try {
      mltps_predict(ctx, *g, e, const_cast<float*>(cov_dev), C, knots_xy, resid, n, lambda, tile_px, out_dev, spline_out, st);
    } catch (...) {
      cudaDeviceSynchronize();   // side / copy / fit streams may still be running on arena blocks and caller buffers
      throw;
    }
  });
}

int mb_mltps_predict(mb_ctx* ctx, const mb_grid* g, const mb_ensemble* e, const float* cov_host, int C,
                     const double* knots_xy, const double* resid, int n, double lambda, int tile_px,
                     double* out_host, mb_spline** spline_out) {
  return guarded([&] {
    MB_REQUIRE(ctx && out_host, "NULL argument");
    check_grid(g);
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->side;   // keeps ctx->stream free for the fit; copies and kernels are ordered on one stream
    ctx->arena.begin(st);
    const size_t ncell = (size_t)g->nrow * g->ncol;
    float* d_cov = nullptr;
    if (e && C > 0) {
      MB_REQUIRE(cov_host, "covariate planes are NULL");
      d_cov = ctx->arena.take_n<float>((size_t)C * ncell);
    }
    double* d_out = ctx->arena.take_n<double>(ncell);
    // planes travel in row blocks on the copy stream, block b's ensemble kernels follow on `side`, the fit
    // runs on the context stream meanwhile
    try {
      mltps_predict(ctx, *g, e, d_cov, C, knots_xy, resid, n, lambda, tile_px, d_out, spline_out, st, cov_host);
      MB_CUDA(cudaMemcpyAsync(out_host, d_out, ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
      MB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
      cudaDeviceSynchronize();   // the copy stream may still be reading cov_host, the side stream the arena
      throw;
    }
  });
}

// ---- the same for one row block of a raster sharded over the communicator ------------------------------------------------
int mb_mltps_predict_shard_dev(mb_ctx* ctx, const mb_grid* g_block, const mb_ensemble* e, const float* cov_dev, int C,
                               const double* knots_xy, const double* resid, int n, double lambda, int root,
                               double* out_dev, mb_spline** spline_out, void* stream) {
  return guarded([&] {
    MB_REQUIRE(ctx && out_dev, "NULL argument");
    check_grid(g_block);
    MB_REQUIRE(root >= 0 && root < ctx->comm_size, "root out of range");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->arena.begin(st);
    try {
      mltps_predict(ctx, *g_block, e, const_cast<float*>(cov_dev), C, knots_xy, resid, n, lambda, 0, out_dev, spline_out, st,
                    nullptr, root);
    } catch (...) {
      cudaDeviceSynchronize();
      throw;
    }
  });
}

int mb_mltps_predict_shard(mb_ctx* ctx, const mb_grid* g_block, const mb_ensemble* e, const float* cov_host, int C,
                           const double* knots_xy, const double* resid, int n, double lambda, int root,
                           double* out_host, mb_spline** spline_out) {
  return guarded([&] {
    MB_REQUIRE(ctx && out_host, "NULL argument");
    check_grid(g_block);
    MB_REQUIRE(root >= 0 && root < ctx->comm_size, "root out of range");
    MB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->side;
    ctx->arena.begin(st);
    const size_t ncell = (size_t)g_block->nrow * g_block->ncol;
    float* d_cov = nullptr;
    if (e && C > 0) {
      MB_REQUIRE(cov_host, "covariate planes are NULL");
      d_cov = ctx->arena.take_n<float>((size_t)C * ncell);
    }
    double* d_out = ctx->arena.take_n<double>(ncell);
    try {
      mltps_predict(ctx, *g_block, e, d_cov, C, knots_xy, resid, n, lambda, 0, d_out, spline_out, st, cov_host, root);
      MB_CUDA(cudaMemcpyAsync(out_host, d_out, ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
      MB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
      cudaDeviceSynchronize();
      throw;
    }
  });
}

}  // extern "C"
