// Shared internals of libmachisplin_b200: context, handles, error plumbing, small device helpers.
// Nothing here is part of the C ABI (see include/machisplin_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <memory>
#include <functional>
#include <algorithm>

#include "../../include/machisplin_b200.h"

namespace mb {

// ---- error plumbing: exceptions never cross the ABI ------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& m);

#define MB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      throw mb::Error(MB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +  \
                                     __FILE__ + ":" + std::to_string(__LINE__) + ")");        \
  } while (0)

#define MB_REQUIRE(cond, msg)                                   \
  do {                                                          \
    if (!(cond)) throw mb::Error(MB_E_ARG, std::string(msg));   \
  } while (0)

template <class F>
int guarded(F&& f) {
  try {
    f();
    return MB_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::bad_alloc&) {
    set_last_error("host allocation failed");
    return MB_E_NOMEM;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return MB_E_NUMERIC;
  }
}

// ---- RAII device buffer -----------------------------------------------------------------
// Long-lived device arrays (model descriptors, spline coefficients) come from a process-wide cache of freed blocks (abi.cu):
// cudaMalloc / cudaFree cost 0.5 - 2 ms each on a 180 GB device and an ensemble handle owns ~30 arrays, which made creating
// and freeing one per response 80 ms of the host-buffer path.  give() synchronises the device before a block can be reused
// (what cudaFree did implicitly); MB_DEV_CACHE=0 in the environment restores plain cudaMalloc / cudaFree.
void* dev_cache_take(size_t bytes);          // throws Error(MB_E_NOMEM)
void dev_cache_give(void* p) noexcept;
void dev_cache_flush() noexcept;             // cudaFree of every cached block of the current device
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() { if (p) dev_cache_give(p); p = nullptr; n = 0; }
  void alloc(size_t n_) {
    release();
    if (n_ == 0) return;
    p = static_cast<T*>(dev_cache_take(n_ * sizeof(T)));
    n = n_;
  }
  void ensure(size_t n_) { if (n_ > n) alloc(n_); }
  void upload(const T* h, size_t cnt, cudaStream_t s) {
    ensure(cnt);
    if (cnt) MB_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
};

// ---- per-context scratch arena ------------------------------------------------------------
// Temporaries of one top-level call are bump-allocated from blocks that stay alive across calls, so
// the hot path performs no cudaMalloc / cudaFree (both synchronise the device).  begin(stream) starts a
// new call: if the previous call ran on a different stream it is drained first, because its kernels may
// still be reading the blocks that are about to be reused.  A call that forks work onto further streams (mb_mltps_predict*:
// fit / ensemble / copy streams) registers them with also_used(); the next begin() drains every one of them, whatever
// stream it runs on and however the forking call ended (normal return, asynchronous return, exception).
struct Arena {
  struct Block { char* p; size_t cap; };
  std::vector<Block> blocks;
  size_t cur = 0, off = 0;
  cudaStream_t last = nullptr;
  bool used = false;
  std::vector<cudaStream_t> forked;
  void also_used(cudaStream_t s) { if (s && std::find(forked.begin(), forked.end(), s) == forked.end()) forked.push_back(s); }
  void begin(cudaStream_t st) {
    if (!forked.empty()) {
      for (cudaStream_t s : forked) cudaStreamSynchronize(s);
      if (used) cudaStreamSynchronize(last);
      forked.clear();
    } else if (used && last != st) {
      cudaStreamSynchronize(last);
    }
    last = st; used = true; cur = 0; off = 0;
  }
  void* take(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    for (; cur < blocks.size(); ++cur, off = 0)
      if (off + bytes <= blocks[cur].cap) { void* r = blocks[cur].p + off; off += bytes; return r; }
    Block b{nullptr, std::max(bytes, size_t(8) << 20)};
    cudaError_t e = cudaMalloc(&b.p, b.cap);
    if (e != cudaSuccess) throw Error(MB_E_NOMEM, std::string("cudaMalloc (arena): ") + cudaGetErrorString(e));
    blocks.push_back(b);
    cur = blocks.size() - 1; off = bytes;
    return b.p;
  }
  template <class T> T* take_n(size_t n) { return static_cast<T*>(take(n * sizeof(T))); }
  template <class T> T* upload(const T* h, size_t n, cudaStream_t st) {
    T* d = take_n<T>(std::max<size_t>(n, 1));
    if (n) {
      cudaError_t e = cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) throw Error(MB_E_CUDA, std::string("cudaMemcpyAsync (arena): ") + cudaGetErrorString(e));
    }
    return d;
  }
  void release() { for (Block& b : blocks) cudaFree(b.p); blocks.clear(); cur = off = 0; used = false; }
};

// Arena-backed array with the DevBuf surface the kernels' call sites use (.p, upload).
template <class T>
struct ABuf {
  T* p = nullptr;
  size_t n = 0;
  Arena* ar = nullptr;
  explicit ABuf(Arena& a, size_t n_ = 0) : ar(&a) { if (n_) { p = a.take_n<T>(n_); n = n_; } }
  void ensure(size_t n_) { if (n_ > n) { p = ar->take_n<T>(n_); n = n_; } }
  void upload(const T* h, size_t cnt, cudaStream_t s) {
    ensure(cnt);
    if (cnt) {
      cudaError_t e = cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess) throw Error(MB_E_CUDA, std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e));
    }
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
};

}  // namespace mb

// ---- handles ----------------------------------------------------------------------------
struct mb_timed_launch {
  const char* name;
  cudaEvent_t start, stop;
};

// What the two-stage tridiagonalisation leaves behind for the coefficient solve at the selected lambda ("coef_impl" = 1):
// M = Q1 B Q1' with B banded, Q1 = H_1 ... H_K in compact-WY form.  All pointers are arena memory of the current call.
struct mb_band_form {
  bool valid = false;
  int m = 0, L = 0, npanel = 0, rmax = 0;
  double* Vall = nullptr;            // panel k: V at Vall + voff[k], ldv[k] x 32 column-major, rows r0[k] .. m-1 of the matrix
  double* Tall = nullptr;            // 32 x 32 row-major per panel
  double* band = nullptr;            // 64 x (m + 64) band storage after stage 1
  double* z1 = nullptr;              // Q1' z, m x L
  std::vector<size_t> voff;
  std::vector<int> r, ldv, r0;
};

struct mb_ctx {
  int device = 0;
  bool timing = false;
  std::vector<mb_timed_launch> timed;
  std::vector<cudaEvent_t> event_pool;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  // fast-evaluator tunables (0 = automatic)
  int cheb_p = 0, leaf_cols = 0, leaf_rows = 0;
  int ens_overlap = 0;        // per-cell ensemble: 0 = 2 = forest kernel, then the tensor-pipe ksvm kernel (default); 1 = side by side on two streams (A/B)
  int leaf_impl = 0;          // TPS-only grid evaluation: 0 = 1 = one warp per box (k_leaf_warp), 2 = one CTA per box (k_leaf_stream)
  int leaf_tma = 0;           // grid-evaluation kernel: 0 = 1 = accumulator tile by one 2-D tensor copy, 2 = one bulk copy per row
  int ens_order = 0;          // forests + tensor-pipe ksvm: 0 = 1 = forest kernel first, 2 = ksvm kernel first
  int ens_tma = 0;            // k_ens_svm_tma: 0 = 1 = covariate tiles by TMA tensor copies when the raster layout allows, 2 = plain loads
  int svm_ctas_per_sm = 0;    // persistent grid of k_ens_svm_tma beside the forest kernel (0 = 2 per SM)
  cudaStream_t ens_aux = nullptr;   // stream of the ksvm kernel in overlap mode + fork / join events
  cudaEvent_t ev_ens[2] = {nullptr, nullptr};
  int tree_levels = 0;        // forest kernel: 0 = 2 = CTA-level + warp-level interval pruning, 1 = CTA-level only (A/B measurements)
  int svm_impl = 0;           // ksvm kernel: 0 = automatic (3 when P <= 8, else 2), 1 = dot products on the tensor pipe as 3 x TF32 (P <= 8),
                              // 2 = packed FP32 (k_ens_svm), 3 = tensor pipe with FP16 split operands (P <= 8); read by mb_ensemble_create and at launch
  // 256-entry (1/m_k, -log(1/m_k)) table for the float64 table-driven log
  mb::DevBuf<double2> logtab;
  int eval_precision = 0;     // fast evaluator: 0 = automatic, 1 = float64 only, 2 = force mixed
  int eigen_impl = 0;         // GCV fit: 0 = in-house tridiagonalisation + bisection, 1 = cuSOLVER Dsyevd (validation)
  int sytrd_mode = 0;         // tridiagonalisation: 0 = default = 3 = two-stage (band reduction + bulge chasing, sbr.cu),
                              // 1 = one-stage persistent kernel with grid barrier, 2 = one-stage, one kernel per phase
  int sbr_debug = 0;          // two-stage path: keep the band matrix of stage 1 for mb_debug_values("sbr_band")
  std::vector<double> dbg_band;
  cudaStream_t sbr_aux = nullptr;   // two-stage path: stream of the look-ahead trailing updates + its events
  cudaEvent_t sbr_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int coef_impl = 0;          // coefficients at the selected lambda: 0 = band form of the two-stage reduction (block band Cholesky +
                              // back-transformation by the stored panel reflectors) when cond(M + lambda I) <= 1e8, else dense Cholesky;
                              // 1 = band form whenever it exists, 2 = always the dense Cholesky of M + lambda I
  mb_band_form band_form;
  int sbr_chase_ctas = 0;     // cap of the bulge chase's grid (0 = one CTA per sweep that can be in flight)
  int sbr_chase_sleep = 0;    // nanoseconds the spinning lanes of the bulge chase sleep between polls (-1 = none)
  int sbr_chase_impl = 0;     // bulge chase: 1 = three warps per sweep + watcher and publisher warps (flags + one fence per step), 2 = three
                              // warps per sweep (96 x 160 registers: the smallest footprint beside a per-cell kernel), 3 = tagged elements
                              // (LL protocol: no flags, no fences; 3 compute + 4 loader warps), 0 = 3 for a fit that has the GPU to
                              // itself, 2 when the ensemble kernels run beside it
  // SM partitions (greenctx.cu): stage 1 of the fit on one, the forest kernel on the other, inside mb_mltps_predict*
  bool gc_tried = false, gc_ok = false;
  void* gc_fit = nullptr;  void* gc_ens = nullptr;                 // CUgreenCtx
  cudaStream_t gc_fit_stream = nullptr, gc_fit_aux = nullptr, gc_ens_stream = nullptr;
  int gc_fit_sms = 0, gc_ens_sms = 0;
  cudaEvent_t gc_ev[4] = {nullptr, nullptr, nullptr, nullptr};     // 0 fork to the fit partition, 1 its join, 2 forest kernels done, 3 spare
  int gc_split = 0;           // "gc_split": SMs of the fit partition (0 = default 64, -1 = no partitions: deferred ensemble)
  int gc_share = 0;           // "gc_share": percent of the raster's rows whose forest kernel runs on the ensemble partition (0 = default)
  bool gc_stage1 = false;     // set by mb_mltps_predict* around its fit: stage 1 of the tridiagonalisation runs on the fit partition
  cudaEvent_t leaf_wait = nullptr;   // consumed by the fast evaluator right before its grid-evaluation kernel (tps_eval.cu)
  bool fit_shares_gpu = false;  // set by mb_mltps_predict* around its fit
  int sbr_qr_impl = 0;        // two-stage path, panel QR: 0 = by cluster size, 1 = panel rows in shared memory, 2 = in registers
  int sbr_qr_grid = 0;        // two-stage path: 1 = panel QR through the software grid barrier even where a cluster fits
  int sytrd_ctas_per_sm = 0;  // persistent grid size (0 = 2 per SM)
  double sytrd_prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per-phase time of the last k_sytrd launch (timing on)
  // scratch reused across calls
  mb::Arena arena;
  // second stream + events: the TPS fit runs beside the per-cell ensemble kernels (mb_mltps_predict*)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_stage1 = nullptr;
  // mb_mltps_predict*: called by the two-stage tridiagonalisation once its stage 1 is enqueued on `stream` (starts the
  // per-cell ensemble kernels behind it); "defer_ensemble" = 0 starts them before the fit instead
  std::function<void()> after_stage1;
  int defer_ensemble = 1;
  // copy stream + per-row-block events of the host-buffer path (H2D of block b+1 overlaps the kernels of block b)
  cudaStream_t copy = nullptr;
  std::vector<cudaEvent_t> ev_blocks;
  // worker lanes of the tile loop (mltps part 3): own stream + arena each, created on first use
  std::vector<std::unique_ptr<mb_ctx>> lanes;
  // multi-GPU (comm.cu): NCCL communicator of this context (ncclComm_t), one process per GPU
  void* comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  std::vector<double> comm_host;   // host staging of broadcast / gathered descriptors
};

struct mb_spline {
  mb_ctx* ctx = nullptr;             // NOT dereferenced by mb_spline_free: a handle may outlive its context
  int device = 0;                    // copied at creation for the free path
  int np = 0;
  std::vector<double> sx, sy;        // scaled knot coordinates
  std::vector<double> kx, ky;        // unscaled
  std::vector<double> c;             // np
  double d[3] = {0, 0, 0};
  double center[2] = {0, 0}, scale[2] = {1, 1};
  double lambda = -1, eff_df = -1, gcv = -1;
  std::vector<double> eta;           // eigenvalues of Q2'KQ2, decreasing (GCV fits only)
  std::vector<double> tri_diag, tri_off, zhat;   // T = Q'(Q2'KQ2)Q and Q'Q2'y of the in-house GCV fit
  double sum_abs_c = 0;              // sum |c_i|  (amplification estimate for the fast evaluator)
  double fscale = 0;                 // max |f(knot_i)|
  mb::DevBuf<double> d_sx, d_sy, d_c;
};

namespace mb {

constexpr double kRbfConst = 0.039788735772973836;  // 1/(8 pi): radbas.constant(m=2, d=2)
constexpr double kD2Clamp = 1e-20;                  // Fortran radfun clamp

struct GridAffine {
  // s_x(col) = (xmin + (col+0.5) rx - cx)/scx ; s_y(row) = (ymax - (row+0.5) ry - cy)/scy
  double xmin, ymax, rx, ry, cx, cy, scx, scy;
};
inline GridAffine make_affine(const mb_grid& g, const mb_spline& s) {
  GridAffine a;
  a.xmin = g.xmin; a.ymax = g.ymax;
  a.rx = (g.xmax - g.xmin) / g.ncol;
  a.ry = (g.ymax - g.ymin) / g.nrow;
  a.cx = s.center[0]; a.cy = s.center[1]; a.scx = s.scale[0]; a.scy = s.scale[1];
  return a;
}
inline void check_grid(const mb_grid* g) {
  MB_REQUIRE(g != nullptr, "grid is NULL");
  MB_REQUIRE(g->nrow > 0 && g->ncol > 0, "grid has no cells");
  MB_REQUIRE(g->xmax > g->xmin && g->ymax > g->ymin, "grid extent is empty");
}
inline void check_window(const mb_grid* g, const mb_window* w) {
  MB_REQUIRE(w != nullptr, "window is NULL");
  MB_REQUIRE(w->r0 >= 0 && w->c0 >= 0 && w->r1 <= g->nrow && w->c1 <= g->ncol && w->r1 > w->r0 && w->c1 > w->c0,
             "window is empty or outside the grid");
}

// Brackets a launch with events when timing is on:  { KernelTimer t(ctx, "k_name", stream); k<<<...>>>(...); }
struct KernelTimer {
  mb_ctx* ctx;
  cudaStream_t st;
  mb_timed_launch rec{};
  bool on;
  KernelTimer(mb_ctx* c, const char* name, cudaStream_t s) : ctx(c), st(s), on(c->timing) {
    c->launches++;
    if (!on) return;
    rec.name = name;
    rec.start = take();
    rec.stop = take();
    cudaEventRecord(rec.start, st);
  }
  ~KernelTimer() {
    if (!on) return;
    cudaEventRecord(rec.stop, st);
    ctx->timed.push_back(rec);
  }
  cudaEvent_t take() {
    if (!ctx->event_pool.empty()) { cudaEvent_t e = ctx->event_pool.back(); ctx->event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};
#define MB_LAUNCH(ctx, name, st) for (mb::KernelTimer _kt(ctx, name, st), *_once = &_kt; _once; _once = nullptr)

// implemented in tps_eval.cu
// uploads, computes sum|c| and fscale = max |f(knot)| (evaluated on the device unless the caller knows it)
void spline_finalize(mb_ctx* ctx, mb_spline* s, double fscale_known = -1.0);
void tps_eval_direct(mb_ctx* ctx, const mb_spline* s, const mb_grid& g, const mb_window& w, double* out,
                     int64_t stride, cudaStream_t st);
// Part 5 fused into the grid-evaluation kernel: final = acc * inv_w + TPS (V73:619, 906-907).  acc is the
// weighted ensemble sum of the window (every kept model; NaN = NA cell) in the padded accumulator layout:
// row stride acc_stride(w) doubles, acc_rows(w) rows, so that a 32-column x bh-row tile of any leaf lattice is
// 16-byte aligned and in bounds for the bulk copies of k_leaf_stream.
struct AccFuse {
  const double* acc;
  int64_t stride;
  double inv_w;
};
inline int64_t acc_stride(const mb_window& w) { return ((int64_t)(w.c1 - w.c0) + 31) / 32 * 32; }
inline int64_t acc_rows(const mb_window& w) { return (int64_t)(w.r1 - w.r0) + 128; }
// Temporaries come from ctx->arena: the caller must have called ctx->arena.begin(st) for this call.
void tps_eval_fast(mb_ctx* ctx, const mb_spline* s, const mb_grid& g, const mb_window& w, double* out,
                   int64_t stride, cudaStream_t st, const AccFuse* fuse = nullptr);
void tps_predict_points_dev(mb_ctx* ctx, const mb_spline* s, const double* x_dev, const double* y_dev, int n,
                            double* out_dev, cudaStream_t st);
void init_logtab(mb_ctx* ctx);

}  // namespace mb
