// Per-cell evaluation of the weighted model ensemble - replaces terra::predict(rast_stack, model_k)
// for the kept models and the accumulation / normalisation of V73:468-475, 497-499, 521-523, 543-545,
// 582-584, 604-606, 619, plus the NA-propagating sum with the TPS surface of part 5 (V73:906-907).
//
//   pred = sum_{k kept} round(w_k, 2) f_k(cov_1..cov_C, LONG, LAT) / sum_all w_k ;  final = pred + TPS
//
// Feature order is xnam (V73:194): the C covariate planes, then LONG, LAT (cell-centre coordinates,
// V73:127-133).  Kernels:
//   k_ens_trees   randomForest + gbm: trees re-laid out breadth-first into 8-byte nodes
//                 {float threshold | leaf value, feature/leaf/child}; children adjacent.  Thresholds are
//                 converted so that every comparison is bit-identical to the double comparison of the
//                 reference on float-valued rasters: x <= t  <=>  x <= fl_down(t);  x < t  <=>
//                 x <= pred(fl_up(t)); splits on LONG / LAT become splits on the integer column / row.
//   k_ens_svm     ksvm rbfdot: exp2(a + b_i + x . sv'_i), float32 dot product, float64 accumulation.
//   k_ens_final   gam + nnet + earth in float64, adds the tree/svm accumulator, divides by the total
//                 weight, adds the TPS surface, applies the NA rule.
#include "common.cuh"
#include "internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>

struct PackedForest {
  int ntrees = 0;
  double offset = 0.0;               // subtracted from leaf values before the float conversion
  double scale = 1.0;                // rf: 1/ntree ; gbm: 1
  mb::DevBuf<int2> nodes;            // .x = float bits (threshold or leaf value), .y = meta
  mb::DevBuf<int> root;              // root node index of every tree
};

struct mb_ensemble {
  mb_ctx* ctx = nullptr;
  mb_grid g{};
  int P = 0, C = 0;
  double w[6] = {0, 0, 0, 0, 0, 0};  // b g n m r v
  double w_total = 1.0;
  bool has[6] = {false, false, false, false, false, false};
  bool only_gbm = false;             // gbm alone returns numbers on NA cells (MissingNode)
  // smooth models (float64)
  mb::DevBuf<double> gam, nn;
  int nn_H = 0;
  double nn_max2 = 1, nn_min = 0;
  int mars_T = 0;
  mb::DevBuf<double> mars_coef, mars_cut;   // per factor
  mb::DevBuf<int> mars_off, mars_var, mars_dir;
  // svm
  int svm_S = 0;
  mb::DevBuf<float> svm_sv;          // [S][P] pre-scaled by 2 sigma log2e
  mb::DevBuf<float> svm_b;           // -sigma |sv|^2 log2e
  mb::DevBuf<float> svm_alpha;
  mb::DevBuf<double> svm_xc, svm_xis; // centre, 1/scale
  double svm_bias = 0, svm_sigma = 0, svm_yc = 0, svm_ys = 1;
  // trees
  PackedForest rf, gbm;
  // reference-layout copies for the point path (float64 thresholds, arbitrary coordinates)
  mb::DevBuf<int> rfp_left, rfp_right, rfp_var;
  mb::DevBuf<signed char> rfp_status;
  mb::DevBuf<double> rfp_split, rfp_pred;
  int rf_ntree = 0, rf_nrnodes = 0;
  mb::DevBuf<int> gbp_off, gbp_var, gbp_left, gbp_right, gbp_miss;
  mb::DevBuf<double> gbp_code;
  int gb_ntrees = 0;
  double gb_initF = 0;
  mb::DevBuf<double> svp_sv, svp_alpha;   // float64 copies for the point path
};

namespace mb {

enum { MB_B = 0, MB_G = 1, MB_N = 2, MB_M = 3, MB_R = 4, MB_V = 5 };
static int letter_index(char c) {
  switch (c) {
    case 'b': return MB_B; case 'g': return MB_G; case 'n': return MB_N;
    case 'm': return MB_M; case 'r': return MB_R; case 'v': return MB_V;
  }
  throw Error(MB_E_ARG, std::string("unknown model letter '") + c + "' (expected b g n m r v)");
}

// ---------------------------------------------------------------------------------------------
// host: threshold conversion and tree re-layout
// ---------------------------------------------------------------------------------------------
static float fl_down(double t) {     // largest float <= t
  float f = (float)t;
  if ((double)f > t) f = std::nextafterf(f, -INFINITY);
  return f;
}
static float fl_up(double t) {       // smallest float >= t
  float f = (float)t;
  if ((double)f < t) f = std::nextafterf(f, INFINITY);
  return f;
}
// largest index i in [0, n) with coord(i) <= t (strict: < t), coord monotone; -1 if none.
template <class F>
static int last_index_le(int n, F coord, double t, bool strict, bool increasing) {
  // returns count of indices satisfying the predicate from the "low-coordinate" end
  int lo = 0, hi = n;   // predicate true on a prefix (increasing) or suffix (decreasing)
  auto pred = [&](int i) { const double v = coord(i); return strict ? v < t : v <= t; };
  if (increasing) {
    while (lo < hi) { int mid = (lo + hi) / 2; if (pred(mid)) lo = mid + 1; else hi = mid; }
    return lo - 1;      // indices [0, lo) satisfy
  }
  while (lo < hi) { int mid = (lo + hi) / 2; if (pred(mid)) hi = mid; else lo = mid + 1; }
  return lo;            // indices [lo, n) satisfy
}

constexpr int kMetaLeaf = 1 << 4;
struct HostNode { float val; int meta; };

// Builds the "go to first child iff feature <= thr" node for a reference split.
//   strict = false: reference goes LEFT iff x <= t (randomForest);  strict = true: LEFT iff x < t (gbm).
// Returns the threshold and whether the children must be swapped (LAT decreases with the row index).
static void convert_split(const mb_grid& g, int C, int var, double t, bool strict, float* thr, bool* swap) {
  *swap = false;
  if (var < C) {
    *thr = strict ? std::nextafterf(fl_up(t), -INFINITY) : fl_down(t);
    return;
  }
  const double rx = (g.xmax - g.xmin) / g.ncol, ry = (g.ymax - g.ymin) / g.nrow;
  if (var == C) {       // LONG = xmin + (col + 0.5) rx, increasing in col
    const int c = last_index_le(g.ncol, [&](int i) { return g.xmin + (i + 0.5) * rx; }, t, strict, true);
    *thr = (float)c;    // col <= c
  } else {              // LAT = ymax - (row + 0.5) ry, decreasing in row: LEFT iff row >= r0
    const int r0 = last_index_le(g.nrow, [&](int i) { return g.ymax - (i + 0.5) * ry; }, t, strict, false);
    *thr = (float)(r0 - 1);   // row <= r0 - 1  ->  RIGHT
    *swap = true;
  }
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
struct EnsGeom {
  double xmin, ymax, rx, ry;
  int nrow, ncol;
};

// ---------------------------------------------------------------------------------------------
// k_ens_trees: one thread per cell, features staged in shared memory ([feature][thread] -> conflict
// free), trees walked from global memory (a tree is a few KB and stays in L1 while the CTA walks it).
// ---------------------------------------------------------------------------------------------
constexpr int kTreeThreads = 256;

template <bool kMissing>
__device__ __forceinline__ float walk_tree(const int2* __restrict__ nodes, int idx, const float* s_feat) {
  for (;;) {
    const int2 nd = __ldg(&nodes[idx]);
    if (nd.y & kMetaLeaf) return __int_as_float(nd.x);
    const float x = s_feat[(nd.y & 15) * kTreeThreads];
    int next = (nd.y >> 5) + (x <= __int_as_float(nd.x) ? 0 : 1);
    if (kMissing) { if (x != x) next = (nd.y >> 5) + 2; }
    idx = next;
  }
}

__global__ void __launch_bounds__(kTreeThreads) k_ens_trees(
    const float* __restrict__ cov, int C, int64_t plane, EnsGeom eg, mb_window w,
    const int2* __restrict__ rf_nodes, const int* __restrict__ rf_root, int rf_n, double rf_w, double rf_off,
    const int2* __restrict__ gb_nodes, const int* __restrict__ gb_root, int gb_n, double gb_w, double gb_init,
    int gbm_missing, double* __restrict__ acc) {
  __shared__ float s_feat[16 * kTreeThreads];
  const int wc = w.c1 - w.c0;
  const int64_t cell = (int64_t)blockIdx.x * kTreeThreads + threadIdx.x;
  const int64_t ncell = (int64_t)(w.r1 - w.r0) * wc;
  const bool live = cell < ncell;
  const int row = w.r0 + (int)(live ? cell / wc : 0);
  const int col = w.c0 + (int)(live ? cell % wc : 0);
  bool anynan = false;
  for (int f = 0; f < C; ++f) {
    const float v = live ? cov[f * plane + (int64_t)row * eg.ncol + col] : 0.f;
    anynan |= (v != v);
    s_feat[f * kTreeThreads + threadIdx.x] = v;
  }
  s_feat[C * kTreeThreads + threadIdx.x] = (float)col;
  s_feat[(C + 1) * kTreeThreads + threadIdx.x] = (float)row;
  if (!live) return;
  const float* sf = s_feat + threadIdx.x;
  double out = 0.0;
  if (rf_n > 0 && !anynan) {
    double s = 0.0;
    for (int t = 0; t < rf_n; ++t) s += (double)walk_tree<false>(rf_nodes, __ldg(&rf_root[t]), sf);
    out += rf_w * (s / rf_n + rf_off);
  }
  if (gb_n > 0 && (!anynan || gbm_missing)) {
    double s = 0.0;
    if (gbm_missing) for (int t = 0; t < gb_n; ++t) s += (double)walk_tree<true>(gb_nodes, __ldg(&gb_root[t]), sf);
    else for (int t = 0; t < gb_n; ++t) s += (double)walk_tree<false>(gb_nodes, __ldg(&gb_root[t]), sf);
    out += gb_w * (s + gb_init);
  }
  acc[cell] += out;
}

// ---------------------------------------------------------------------------------------------
// k_ens_svm: support vectors streamed through shared memory
// ---------------------------------------------------------------------------------------------
constexpr int kSvmThreads = 256;
constexpr int kSvmChunk = 128;

template <int PP>
__global__ void __launch_bounds__(kSvmThreads) k_ens_svm(
    const float* __restrict__ cov, int C, int64_t plane, EnsGeom eg, mb_window w,
    const float* __restrict__ sv, const float* __restrict__ svb, const float* __restrict__ alpha, int S,
    const double* __restrict__ xc, const double* __restrict__ xis, double sigma, double bias, double ys, double yc,
    double wv, double* __restrict__ acc) {
  __shared__ float s_sv[kSvmChunk * PP];
  __shared__ float s_b[kSvmChunk], s_a[kSvmChunk];
  const int wc = w.c1 - w.c0;
  const int64_t cell = (int64_t)blockIdx.x * kSvmThreads + threadIdx.x;
  const int64_t ncell = (int64_t)(w.r1 - w.r0) * wc;
  const bool live = cell < ncell;
  const int row = w.r0 + (int)(live ? cell / wc : 0);
  const int col = w.c0 + (int)(live ? cell % wc : 0);
  float x[PP];
  double n2 = 0.0;
  bool anynan = false;
#pragma unroll
  for (int f = 0; f < PP; ++f) {
    double v;
    if (f < C) v = live ? (double)cov[f * plane + (int64_t)row * eg.ncol + col] : 0.0;
    else if (f == C) v = eg.xmin + (col + 0.5) * eg.rx;
    else v = eg.ymax - (row + 0.5) * eg.ry;
    anynan |= (v != v);
    const double xs = (v - xc[f]) * xis[f];
    n2 += xs * xs;
    x[f] = (float)xs;
  }
  const float a0 = (float)(-sigma * n2 * 1.4426950408889634);
  double total = 0.0;
  for (int base = 0; base < S; base += kSvmChunk) {
    const int n = min(kSvmChunk, S - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n * PP; i += kSvmThreads) s_sv[i] = sv[(size_t)base * PP + i];
    for (int i = threadIdx.x; i < n; i += kSvmThreads) { s_b[i] = svb[base + i]; s_a[i] = alpha[base + i]; }
    __syncthreads();
    float part = 0.f;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
      float e = a0 + s_b[i];
#pragma unroll
      for (int f = 0; f < PP; ++f) e = fmaf(x[f], s_sv[i * PP + f], e);
      part = fmaf(s_a[i], exp2f(e), part);
    }
    total += (double)part;
  }
  if (live && !anynan) acc[cell] += wv * ((total - bias) * ys + yc);
}

// ---------------------------------------------------------------------------------------------
// k_ens_final: smooth models in float64 + combine
// ---------------------------------------------------------------------------------------------
struct SmoothParams {
  const double* gam; double w_g;
  const double* nn; int nn_H; double nn_max2, nn_min, w_n;
  int mars_T; const double* mars_coef; const int* mars_off; const int* mars_var; const int* mars_dir;
  const double* mars_cut; double w_m;
  double w_total;
  int only_gbm;
};

__global__ void __launch_bounds__(256) k_ens_final(const float* __restrict__ cov, int C, int64_t plane, EnsGeom eg,
                                                   mb_window w, SmoothParams sp, const double* __restrict__ acc,
                                                   const double* __restrict__ tps, double* __restrict__ out) {
  const int col = w.c0 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int row = w.r0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (col >= w.c1 || row >= w.r1) return;
  const int64_t o = (int64_t)(row - w.r0) * (w.c1 - w.c0) + (col - w.c0);
  double x[16];
  bool anynan = false;
  for (int f = 0; f < C; ++f) {
    const float v = cov[f * plane + (int64_t)row * eg.ncol + col];
    anynan |= (v != v);
    x[f] = (double)v;
  }
  x[C] = eg.xmin + (col + 0.5) * eg.rx;
  x[C + 1] = eg.ymax - (row + 0.5) * eg.ry;
  const int P = C + 2;
  double s = acc ? acc[o] : 0.0;
  if (!anynan) {
    if (sp.gam) {
      double v = sp.gam[0];
      for (int f = 0; f < P; ++f) v = fma(sp.gam[1 + f], x[f], v);
      s = fma(sp.w_g, v, s);
    }
    if (sp.nn) {
      const double* wo = sp.nn + (P + 1) * sp.nn_H;
      double v = wo[0];
      for (int h = 0; h < sp.nn_H; ++h) {
        const double* wh = sp.nn + h * (P + 1);
        double z = wh[0];
        for (int f = 0; f < P; ++f) z = fma(wh[1 + f], x[f], z);
        v = fma(wo[1 + h], 1.0 / (1.0 + exp(-z)), v);
      }
      s = fma(sp.w_n, v * sp.nn_max2 + sp.nn_min, s);
    }
    if (sp.mars_T > 0) {
      double v = 0.0;
      for (int t = 0; t < sp.mars_T; ++t) {
        double b = sp.mars_coef[t];
        for (int q = sp.mars_off[t]; q < sp.mars_off[t + 1]; ++q) {
          const double xv = x[sp.mars_var[q]];
          const int dir = sp.mars_dir[q];
          b *= (dir == 2) ? xv : fmax(0.0, dir * (xv - sp.mars_cut[q]));
        }
        v += b;
      }
      s = fma(sp.w_m, v, s);
    }
  }
  double r = s / sp.w_total;
  if (anynan && !sp.only_gbm) r = __longlong_as_double(0x7ff8000000000000LL);
  if (tps) r += tps[o];
  out[o] = r;
}

// ---------------------------------------------------------------------------------------------
// point path (residual side, V73:477-611): float64 features given explicitly, reference tree layout
// ---------------------------------------------------------------------------------------------
struct PointModels {
  SmoothParams sp;
  int P;
  // rf
  int rf_ntree, rf_nrnodes; const int* rf_left; const int* rf_right; const signed char* rf_status;
  const int* rf_var; const double* rf_split; const double* rf_pred; double w_r;
  // gbm
  int gb_ntrees; double gb_initF; const int* gb_off; const int* gb_var; const double* gb_code;
  const int* gb_left; const int* gb_right; const int* gb_miss; double w_b;
  // svm
  int S; const double* sv; const double* alpha; const double* xc; const double* xis;
  double sigma, bias, ys, yc, w_v;
};

__global__ void __launch_bounds__(128) k_ens_points(const double* __restrict__ X, int n, PointModels pm,
                                                    double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[16];
  bool anynan = false;
  for (int f = 0; f < pm.P; ++f) { x[f] = X[(size_t)f * n + i]; anynan |= (x[f] != x[f]); }
  const SmoothParams& sp = pm.sp;
  double s = 0.0;
  if (!anynan) {
    if (sp.gam) {
      double v = sp.gam[0];
      for (int f = 0; f < pm.P; ++f) v = fma(sp.gam[1 + f], x[f], v);
      s = fma(sp.w_g, v, s);
    }
    if (sp.nn) {
      const double* wo = sp.nn + (pm.P + 1) * sp.nn_H;
      double v = wo[0];
      for (int h = 0; h < sp.nn_H; ++h) {
        const double* wh = sp.nn + h * (pm.P + 1);
        double z = wh[0];
        for (int f = 0; f < pm.P; ++f) z = fma(wh[1 + f], x[f], z);
        v = fma(wo[1 + h], 1.0 / (1.0 + exp(-z)), v);
      }
      s = fma(sp.w_n, v * sp.nn_max2 + sp.nn_min, s);
    }
    if (sp.mars_T > 0) {
      double v = 0.0;
      for (int t = 0; t < sp.mars_T; ++t) {
        double b = sp.mars_coef[t];
        for (int q = sp.mars_off[t]; q < sp.mars_off[t + 1]; ++q) {
          const double xv = x[sp.mars_var[q]];
          const int dir = sp.mars_dir[q];
          b *= (dir == 2) ? xv : fmax(0.0, dir * (xv - sp.mars_cut[q]));
        }
        v += b;
      }
      s = fma(sp.w_m, v, s);
    }
    if (pm.rf_ntree > 0) {
      double a = 0.0;
      for (int t = 0; t < pm.rf_ntree; ++t) {
        const size_t o = (size_t)t * pm.rf_nrnodes;
        int k = 0;
        while (pm.rf_status[o + k] != -1)
          k = (x[pm.rf_var[o + k] - 1] <= pm.rf_split[o + k] ? pm.rf_left[o + k] : pm.rf_right[o + k]) - 1;
        a += pm.rf_pred[o + k];
      }
      s = fma(pm.w_r, a / pm.rf_ntree, s);
    }
    if (pm.S > 0) {
      double a = 0.0;
      for (int q = 0; q < pm.S; ++q) {
        double d2 = 0.0;
        for (int f = 0; f < pm.P; ++f) {
          const double d = (x[f] - pm.xc[f]) * pm.xis[f] - pm.sv[(size_t)q * pm.P + f];
          d2 = fma(d, d, d2);
        }
        a = fma(pm.alpha[q], exp(-pm.sigma * d2), a);
      }
      s = fma(pm.w_v, (a - pm.bias) * pm.ys + pm.yc, s);
    }
  }
  if (pm.gb_ntrees > 0 && (!anynan || sp.only_gbm)) {
    double a = pm.gb_initF;
    for (int t = 0; t < pm.gb_ntrees; ++t) {
      const int o = pm.gb_off[t];
      int k = 0;
      while (pm.gb_var[o + k] != -1) {
        const double xv = x[pm.gb_var[o + k]];
        k = (xv != xv) ? pm.gb_miss[o + k] : (xv < pm.gb_code[o + k] ? pm.gb_left[o + k] : pm.gb_right[o + k]);
      }
      a += pm.gb_code[o + k];
    }
    s = fma(pm.w_b, a, s);
  }
  double r = s / sp.w_total;
  if (anynan && !sp.only_gbm) r = __longlong_as_double(0x7ff8000000000000LL);
  out[i] = r;
}

// ---------------------------------------------------------------------------------------------
// host: create / evaluate
// ---------------------------------------------------------------------------------------------
mb_ensemble* ensemble_create(mb_ctx* ctx, const mb_grid& g, const mb_models& m, const char* kept, const double* w,
                             double w_total) {
  cudaStream_t st = ctx->stream;
  auto e = std::make_unique<mb_ensemble>();
  e->ctx = ctx;
  e->g = g;
  e->P = m.P;
  e->C = m.P - 2;
  MB_REQUIRE(m.P >= 2 && m.P <= 16, "P = C + 2 must be in [2, 16]");
  MB_REQUIRE(w_total > 0, "w_total must be positive");
  const int nk = (int)std::strlen(kept);
  for (int i = 0; i < nk; ++i) {
    const int k = letter_index(kept[i]);
    MB_REQUIRE(!e->has[k], "model letter repeated in `kept`");
    e->has[k] = true;
    e->w[k] = w[i];
  }
  e->w_total = w_total;
  e->only_gbm = e->has[MB_B] && nk == 1;
  const int P = m.P, C = e->C;

  if (e->has[MB_G]) {
    MB_REQUIRE(m.gam_coef, "gam kept but gam_coef is NULL");
    e->gam.upload(m.gam_coef, P + 1, st);
  }
  if (e->has[MB_N]) {
    MB_REQUIRE(m.nn_wts && m.nn_H >= 1, "nnet kept but descriptor is empty");
    e->nn_H = m.nn_H; e->nn_max2 = m.nn_max2; e->nn_min = m.nn_min;
    e->nn.upload(m.nn_wts, (size_t)(P + 1) * m.nn_H + m.nn_H + 1, st);
  }
  if (e->has[MB_M]) {
    MB_REQUIRE(m.mars_T >= 1 && m.mars_dirs && m.mars_cuts && m.mars_coef, "earth kept but descriptor is empty");
    std::vector<int> off(1, 0), var, dir;
    std::vector<double> cut;
    for (int t = 0; t < m.mars_T; ++t) {
      for (int v = 0; v < P; ++v) {
        const int dd = m.mars_dirs[(size_t)t * P + v];
        if (dd != 0) {
          MB_REQUIRE(dd == 1 || dd == -1 || dd == 2, "earth dirs must be 0, +-1 or 2");
          var.push_back(v); dir.push_back(dd); cut.push_back(m.mars_cuts[(size_t)t * P + v]);
        }
      }
      off.push_back((int)var.size());
    }
    e->mars_T = m.mars_T;
    e->mars_coef.upload(m.mars_coef, m.mars_T, st);
    e->mars_off.upload(off, st);
    if (var.empty()) { var.push_back(0); dir.push_back(1); cut.push_back(0.0); }
    e->mars_var.upload(var, st); e->mars_dir.upload(dir, st); e->mars_cut.upload(cut, st);
  }
  if (e->has[MB_V]) {
    MB_REQUIRE(m.svm_S >= 1 && m.svm_sv && m.svm_alpha && m.svm_x_center && m.svm_x_scale, "ksvm kept but descriptor is empty");
    const int S = m.svm_S;
    const double l2e = 1.4426950408889634;
    std::vector<float> sv((size_t)S * P), b(S), al(S);
    std::vector<double> xis(P);
    for (int f = 0; f < P; ++f) {
      MB_REQUIRE(m.svm_x_scale[f] != 0, "ksvm x.scale has a zero entry");
      xis[f] = 1.0 / m.svm_x_scale[f];
    }
    for (int i = 0; i < S; ++i) {
      double n2 = 0;
      for (int f = 0; f < P; ++f) {
        const double v = m.svm_sv[(size_t)i * P + f];
        n2 += v * v;
        sv[(size_t)i * P + f] = (float)(2.0 * m.svm_sigma * l2e * v);
      }
      b[i] = (float)(-m.svm_sigma * n2 * l2e);
      al[i] = (float)m.svm_alpha[i];
    }
    e->svm_S = S;
    e->svm_sv.upload(sv, st); e->svm_b.upload(b, st); e->svm_alpha.upload(al, st);
    e->svm_xc.upload(m.svm_x_center, P, st);
    e->svm_xis.upload(xis, st);
    e->svm_bias = m.svm_b; e->svm_sigma = m.svm_sigma; e->svm_yc = m.svm_y_center; e->svm_ys = m.svm_y_scale;
    e->svp_sv.upload(m.svm_sv, (size_t)S * P, st);
    e->svp_alpha.upload(m.svm_alpha, S, st);
  }
  if (e->has[MB_R]) {
    MB_REQUIRE(m.rf_ntree >= 1 && m.rf_nrnodes >= 1 && m.rf_left && m.rf_right && m.rf_status && m.rf_bestvar &&
                   m.rf_split && m.rf_nodepred, "randomForest kept but descriptor is empty");
    const int nt = m.rf_ntree, nn = m.rf_nrnodes;
    // leaf offset: mean of the root-reachable leaf values of the first tree is a good centre
    double off = 0; int cnt = 0;
    for (int k = 0; k < nn; ++k) if (m.rf_status[k] == -1 && (k == 0 || true)) { off += m.rf_nodepred[k]; ++cnt; }
    off = cnt ? off / cnt : 0.0;
    std::vector<int2> nodes;
    std::vector<int> roots(nt);
    nodes.reserve((size_t)nt * 64);
    for (int t = 0; t < nt; ++t) {
      const size_t o = (size_t)t * nn;
      // breadth-first re-layout with adjacent children
      std::queue<std::pair<int, size_t>> q;   // (reference node, packed position)
      roots[t] = (int)nodes.size();
      nodes.push_back(make_int2(0, 0));
      q.push({0, (size_t)roots[t]});
      while (!q.empty()) {
        auto [k, pos] = q.front();
        q.pop();
        MB_REQUIRE(k >= 0 && k < nn, "randomForest: daughter index out of range");
        if (m.rf_status[o + k] == -1) {
          const float v = (float)(m.rf_nodepred[o + k] - off);
          nodes[pos] = make_int2(__builtin_bit_cast(int, v), kMetaLeaf);
          continue;
        }
        const int var = m.rf_bestvar[o + k] - 1;
        MB_REQUIRE(var >= 0 && var < P, "randomForest: bestvar out of range");
        float thr; bool swap;
        convert_split(g, C, var, m.rf_split[o + k], false, &thr, &swap);
        const size_t child = nodes.size();
        MB_REQUIRE(child < ((size_t)1 << 26), "forest too large for the packed node format");
        nodes.push_back(make_int2(0, 0));
        nodes.push_back(make_int2(0, 0));
        nodes[pos] = make_int2(__builtin_bit_cast(int, thr), (int)(child << 5) | var);
        const int l = m.rf_left[o + k] - 1, r = m.rf_right[o + k] - 1;
        q.push({swap ? r : l, child});
        q.push({swap ? l : r, child + 1});
      }
    }
    e->rf.ntrees = nt; e->rf.offset = off;
    e->rf.nodes.upload(nodes, st); e->rf.root.upload(roots, st);
    e->rf_ntree = nt; e->rf_nrnodes = nn;
    const size_t tot = (size_t)nt * nn;
    e->rfp_left.upload(m.rf_left, tot, st); e->rfp_right.upload(m.rf_right, tot, st);
    e->rfp_status.upload((const signed char*)m.rf_status, tot, st); e->rfp_var.upload(m.rf_bestvar, tot, st);
    e->rfp_split.upload(m.rf_split, tot, st); e->rfp_pred.upload(m.rf_nodepred, tot, st);
  }
  if (e->has[MB_B]) {
    MB_REQUIRE(m.gbm_ntrees >= 1 && m.gbm_tree_off && m.gbm_splitvar && m.gbm_splitcode && m.gbm_left &&
                   m.gbm_right && m.gbm_missing, "gbm kept but descriptor is empty");
    const int nt = m.gbm_ntrees;
    std::vector<int2> nodes;
    std::vector<int> roots(nt);
    for (int t = 0; t < nt; ++t) {
      const int o = m.gbm_tree_off[t], cntn = m.gbm_tree_off[t + 1] - o;
      std::queue<std::pair<int, size_t>> q;
      roots[t] = (int)nodes.size();
      nodes.push_back(make_int2(0, 0));
      q.push({0, (size_t)roots[t]});
      while (!q.empty()) {
        auto [k, pos] = q.front();
        q.pop();
        MB_REQUIRE(k >= 0 && k < cntn, "gbm: child index out of range");
        const int var = m.gbm_splitvar[o + k];
        if (var == -1) {
          const float v = (float)m.gbm_splitcode[o + k];
          nodes[pos] = make_int2(__builtin_bit_cast(int, v), kMetaLeaf);
          continue;
        }
        MB_REQUIRE(var >= 0 && var < P, "gbm: SplitVar out of range");
        float thr; bool swap;
        convert_split(g, C, var, m.gbm_splitcode[o + k], true, &thr, &swap);
        const size_t child = nodes.size();
        MB_REQUIRE(child < ((size_t)1 << 26), "boosted model too large for the packed node format");
        for (int c3 = 0; c3 < 3; ++c3) nodes.push_back(make_int2(0, 0));
        nodes[pos] = make_int2(__builtin_bit_cast(int, thr), (int)(child << 5) | var);
        const int l = m.gbm_left[o + k], r = m.gbm_right[o + k];
        q.push({swap ? r : l, child});
        q.push({swap ? l : r, child + 1});
        q.push({m.gbm_missing[o + k], child + 2});
      }
    }
    e->gbm.ntrees = nt;
    e->gbm.nodes.upload(nodes, st); e->gbm.root.upload(roots, st);
    e->gb_ntrees = nt; e->gb_initF = m.gbm_initF;
    const size_t tot = m.gbm_tree_off[nt];
    e->gbp_off.upload(m.gbm_tree_off, nt + 1, st); e->gbp_var.upload(m.gbm_splitvar, tot, st);
    e->gbp_code.upload(m.gbm_splitcode, tot, st); e->gbp_left.upload(m.gbm_left, tot, st);
    e->gbp_right.upload(m.gbm_right, tot, st); e->gbp_miss.upload(m.gbm_missing, tot, st);
  }
  MB_CUDA(cudaStreamSynchronize(st));
  return e.release();
}

void ensemble_free(mb_ensemble* e) {
  if (!e) return;
  if (e->ctx) cudaSetDevice(e->ctx->device);
  delete e;
}

mb_grid ensemble_grid(const mb_ensemble* e) { return e->g; }

static SmoothParams smooth_params(const mb_ensemble* e) {
  SmoothParams sp{};
  sp.gam = e->has[MB_G] ? e->gam.p : nullptr; sp.w_g = e->w[MB_G];
  sp.nn = e->has[MB_N] ? e->nn.p : nullptr; sp.nn_H = e->nn_H; sp.nn_max2 = e->nn_max2; sp.nn_min = e->nn_min;
  sp.w_n = e->w[MB_N];
  sp.mars_T = e->has[MB_M] ? e->mars_T : 0; sp.mars_coef = e->mars_coef.p; sp.mars_off = e->mars_off.p;
  sp.mars_var = e->mars_var.p; sp.mars_dir = e->mars_dir.p; sp.mars_cut = e->mars_cut.p; sp.w_m = e->w[MB_M];
  sp.w_total = e->w_total;
  sp.only_gbm = e->only_gbm ? 1 : 0;
  return sp;
}

template <int PP>
static void launch_svm(const mb_ensemble* e, const float* cov, int64_t plane, const EnsGeom& eg, const mb_window& w,
                       double* acc, int64_t ncell, cudaStream_t st) {
  mb_ctx* ctx = e->ctx;
  MB_LAUNCH(ctx, "k_ens_svm", st) k_ens_svm<PP><<<(unsigned)((ncell + kSvmThreads - 1) / kSvmThreads), kSvmThreads, 0, st>>>(
      cov, e->C, plane, eg, w, e->svm_sv.p, e->svm_b.p, e->svm_alpha.p, e->svm_S, e->svm_xc.p, e->svm_xis.p,
      e->svm_sigma, e->svm_bias, e->svm_ys, e->svm_yc, e->w[MB_V], acc);
}

void ensemble_eval(mb_ctx* ctx, const mb_ensemble* e, const float* cov, int C, const mb_spline* spline,
                   const double* tps_surface, const mb_window* wp, double* out, cudaStream_t st) {
  const mb_grid& g = e->g;
  check_window(&g, wp);
  const mb_window w = *wp;
  MB_REQUIRE(C == e->C, "number of covariate planes does not match the model descriptors (P = C + 2)");
  const int64_t plane = (int64_t)g.nrow * g.ncol;
  const int64_t ncell = (int64_t)(w.r1 - w.r0) * (w.c1 - w.c0);
  EnsGeom eg{g.xmin, g.ymax, (g.xmax - g.xmin) / g.ncol, (g.ymax - g.ymin) / g.nrow, g.nrow, g.ncol};
  const bool heavy = e->has[MB_R] || e->has[MB_B] || e->has[MB_V];
  DevBuf<double> acc;
  if (heavy) {
    acc.alloc((size_t)ncell);
    MB_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double) * ncell, st));
    if (e->has[MB_R] || e->has[MB_B]) {
      MB_LAUNCH(ctx, "k_ens_trees", st) k_ens_trees<<<(unsigned)((ncell + kTreeThreads - 1) / kTreeThreads), kTreeThreads, 0, st>>>(
          cov, C, plane, eg, w, e->rf.nodes.p, e->rf.root.p, e->has[MB_R] ? e->rf.ntrees : 0, e->w[MB_R],
          e->rf.offset, e->gbm.nodes.p, e->gbm.root.p, e->has[MB_B] ? e->gbm.ntrees : 0, e->w[MB_B], e->gb_initF,
          e->only_gbm ? 1 : 0, acc.p);
    }
    if (e->has[MB_V]) {
      switch (e->P) {
#define MB_SVM_CASE(n) case n: launch_svm<n>(e, cov, plane, eg, w, acc.p, ncell, st); break;
        MB_SVM_CASE(2) MB_SVM_CASE(3) MB_SVM_CASE(4) MB_SVM_CASE(5) MB_SVM_CASE(6) MB_SVM_CASE(7) MB_SVM_CASE(8)
        MB_SVM_CASE(9) MB_SVM_CASE(10) MB_SVM_CASE(11) MB_SVM_CASE(12) MB_SVM_CASE(13) MB_SVM_CASE(14)
        MB_SVM_CASE(15) MB_SVM_CASE(16)
#undef MB_SVM_CASE
      }
    }
  }
  const double* tps = tps_surface;
  if (spline) {
    MB_REQUIRE(!tps_surface, "pass either a spline or a precomputed TPS surface, not both");
    tps_eval_fast(ctx, spline, g, w, out, w.c1 - w.c0, st);   // TPS lands in `out`, combined in place below
    tps = out;
  }
  dim3 grid((w.c1 - w.c0 + 31) / 32, (w.r1 - w.r0 + 7) / 8);
  MB_LAUNCH(ctx, "k_ens_final", st) k_ens_final<<<grid, 256, 0, st>>>(cov, C, plane, eg, w, smooth_params(e), heavy ? acc.p : nullptr, tps, out);
  MB_CUDA(cudaGetLastError());
  if (heavy) MB_CUDA(cudaStreamSynchronize(st));   // acc is a stream-ordered temporary
}

void ensemble_predict_points(mb_ctx* ctx, const mb_ensemble* e, const double* X, int n, double* out_host) {
  if (n <= 0) return;
  cudaStream_t st = ctx->stream;
  DevBuf<double> dX((size_t)n * e->P), dO(n);
  dX.upload(X, (size_t)n * e->P, st);
  PointModels pm{};
  pm.sp = smooth_params(e);
  pm.P = e->P;
  pm.rf_ntree = e->has[MB_R] ? e->rf_ntree : 0; pm.rf_nrnodes = e->rf_nrnodes;
  pm.rf_left = e->rfp_left.p; pm.rf_right = e->rfp_right.p; pm.rf_status = e->rfp_status.p;
  pm.rf_var = e->rfp_var.p; pm.rf_split = e->rfp_split.p; pm.rf_pred = e->rfp_pred.p; pm.w_r = e->w[MB_R];
  pm.gb_ntrees = e->has[MB_B] ? e->gb_ntrees : 0; pm.gb_initF = e->gb_initF; pm.gb_off = e->gbp_off.p;
  pm.gb_var = e->gbp_var.p; pm.gb_code = e->gbp_code.p; pm.gb_left = e->gbp_left.p; pm.gb_right = e->gbp_right.p;
  pm.gb_miss = e->gbp_miss.p; pm.w_b = e->w[MB_B];
  pm.S = e->has[MB_V] ? e->svm_S : 0; pm.sv = e->svp_sv.p; pm.alpha = e->svp_alpha.p; pm.xc = e->svm_xc.p;
  pm.xis = e->svm_xis.p; pm.sigma = e->svm_sigma; pm.bias = e->svm_bias; pm.ys = e->svm_ys; pm.yc = e->svm_yc;
  pm.w_v = e->w[MB_V];
  MB_LAUNCH(ctx, "k_ens_points", st) k_ens_points<<<(n + 127) / 128, 128, 0, st>>>(dX.p, n, pm, dO.p);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaMemcpyAsync(out_host, dO.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace mb
